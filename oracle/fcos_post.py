"""oracle/fcos_post.py -- TEST INFRASTRUCTURE ONLY.

numpy (fp32, one rounding per operation) restatement of the reference's anchor-free (FCOS) inference post-processing:
  head output transform   nerf_rpn/model/fcos/fcos.py:116-126   (per-level Scale, ReLU on the 6 distances, x stride)
  locations               nerf_rpn/model/fcos/fcos.py:221-250   (idx*stride + stride//2)
  per-level selection     nerf_rpn/model/fcos/inference.py:48-129 (sigmoid, candidates > pre_nms_thresh, x centerness,
                                                                  top-k, decode, clip, min-size, sqrt score)
  OBB decode              nerf_rpn/model/fcos/utils.py:12-61    (decode_fcos_obb)
  cross-level NMS + cap   nerf_rpn/model/fcos/inference.py:164-195 (one NMS over all levels, kthvalue cut)
Tie conventions as in oracle/rpn_post.py: equal scores -> lower index first.  Pinned by tests/golden/fcos_small_*.npz.
"""
import numpy as np

from . import box as obox
from .rpn_post import sigmoid

F = np.float32


def locations(grid, stride):
    gx, gy, gz = grid
    X, Y, Z = np.meshgrid(np.arange(gx, dtype=F) * F(stride), np.arange(gy, dtype=F) * F(stride),
                          np.arange(gz, dtype=F) * F(stride), indexing="ij")
    return (np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1) + F(stride // 2)).astype(F)


def head_transform(raw_reg, scale, stride):
    """bbox_pred * scale, ReLU on the first 6 channels, those 6 multiplied by the FPN stride (eval mode)."""
    r = (np.asarray(raw_reg, dtype=F) * F(scale)).astype(F)
    r[:, :6] = np.maximum(r[:, :6], F(0)) * F(stride)
    return r.astype(F)


def _norm2(x, y):
    return np.sqrt((x * x + y * y).astype(F)).astype(F)


def decode_obb(loc, reg):
    loc = np.asarray(loc, dtype=F); r = np.asarray(reg, dtype=F)
    with np.errstate(all="ignore"):
        x0 = loc[:, 0] - r[:, 0]; y0 = loc[:, 1] - r[:, 1]; z0 = loc[:, 2] - r[:, 2]
        x1 = loc[:, 0] + r[:, 3]; y1 = loc[:, 1] + r[:, 4]; z1 = loc[:, 2] + r[:, 5]
        vx = (x1 + x0) / F(2) + r[:, 6] * (x1 - x0)
        vy = (y1 + y0) / F(2) + r[:, 7] * (y1 - y0)
        vx = np.minimum(np.maximum(vx, x0), x1).astype(F); vy = np.minimum(np.maximum(vy, y0), y1).astype(F)
        cx = (x0 + x1) / F(2); cy = (y0 + y1) / F(2); cz = (z0 + z1) / F(2)
        v0x = vx - cx; v0y = y1 - cy; v1x = x1 - cx; v1y = vy - cy
        d0 = _norm2(v0x, v0y); d1 = _norm2(v1x, v1y)
        dmax = np.maximum(d0, d1)
        v0x = v0x / (d0 + F(1e-7)) * dmax + cx; v0y = v0y / (d0 + F(1e-7)) * dmax + cy
        v1x = v1x / (d1 + F(1e-7)) * dmax + cx; v1y = v1y / (d1 + F(1e-7)) * dmax + cy
        ln = _norm2(v0x - v1x, v0y - v1y)
        mx = (v0x + v1x) / F(2) - cx; my = (v0y + v1y) / F(2) - cy
        w = _norm2(mx, my) * F(2)
        h = z1 - z0
        mx = np.where((mx == 0) & (my == 0), F(1e-7), mx).astype(F)
        theta = np.arctan2(my.astype(np.float64), mx.astype(np.float64)).astype(F)
    return np.stack([cx, cy, cz, w, ln, h, theta], 1).astype(F)


def fcos_proposals(raw_cls, raw_reg, raw_ctr, scales, grids, strides, grid_size, use_obb, pre_nms_thresh=0.0,
                   pre_nms_top_n=2500, nms_thresh=0.3, post_top_n=2500, min_size=0.0, padded=False, reg_is_raw=True):
    """One scene. raw_*[l]: conv outputs of level l in voxel order ((x*gy+y)*gz+z): (V,), (V,6|8), (V,).
    grid_size: the scene's own (unpadded) extent (inference.py:118-119).  Returns boxes (K, 1+6|7) with the level id in
    column 0, scores (K,).  reg_is_raw=False: raw_reg already is the head's output (scaled / rectified / x stride)."""
    det, sc = [], []
    for l in range(len(raw_cls)):
        loc = locations(grids[l], strides[l])
        cls = sigmoid(np.asarray(raw_cls[l], dtype=F))
        ctr = sigmoid(np.asarray(raw_ctr[l], dtype=F))
        reg = head_transform(raw_reg[l], scales[l], strides[l]) if reg_is_raw else np.asarray(raw_reg[l], dtype=F)
        if padded:
            m = (loc[:, 0] < grid_size[0]) & (loc[:, 1] < grid_size[1]) & (loc[:, 2] < grid_size[2])
            cls = np.where(m, cls, F(-1e5)).astype(F)
        cand = np.nonzero(cls > F(pre_nms_thresh))[0]
        k = min(cand.shape[0], pre_nms_top_n)
        score = (cls * ctr).astype(F)[cand]
        if cand.shape[0] > k:
            order = np.lexsort((cand, -score.astype(np.float64)))[:k]
            cand, score = cand[order], score[order]
        lo, rg = loc[cand], reg[cand]
        if not use_obb:
            b = np.stack([lo[:, 0] - rg[:, 0], lo[:, 1] - rg[:, 1], lo[:, 2] - rg[:, 2],
                          lo[:, 0] + rg[:, 3], lo[:, 1] + rg[:, 4], lo[:, 2] + rg[:, 5]], 1).astype(F)
            for a in range(3):
                b[:, a] = np.clip(b[:, a], F(0), F(grid_size[a])); b[:, 3 + a] = np.clip(b[:, 3 + a], F(0), F(grid_size[a]))
            keep = ((b[:, 3] - b[:, 0]) >= F(min_size)) & ((b[:, 4] - b[:, 1]) >= F(min_size)) & ((b[:, 5] - b[:, 2]) >= F(min_size))
        else:
            b = decode_obb(lo, rg)
            keep = (b[:, 3] >= F(min_size)) & (b[:, 4] >= F(min_size)) & (b[:, 5] >= F(min_size))
        b, score = b[keep], score[keep]
        with np.errstate(all="ignore"):
            s = np.sqrt(score).astype(F)
        det.append(np.concatenate([np.full((b.shape[0], 1), l, dtype=F), b], 1)); sc.append(s)
    boxes = np.concatenate(det); scores = np.concatenate(sc)
    keep = obox.nms(boxes[:, 1:], scores, nms_thresh)
    bk, sk = boxes[keep], scores[keep]
    n = keep.shape[0]
    if n > post_top_n > 0:
        thr = np.sort(sk)[n - post_top_n]           # kthvalue(n - top_n + 1): keeps ties
        m = sk >= thr
        bk, sk = bk[m], sk[m]
    return bk, sk
