"""oracle/rpn_post.py -- TEST INFRASTRUCTURE ONLY.

numpy restatement (fp32, one rounding per operation) of the reference's RPN post-processing:
  anchors            nerf_rpn/model/anchor.py:51-122,154-174
  flat anchor index  nerf_rpn/model/rpn.py:20-27,105-130
  AABB decode        nerf_rpn/model/coder/AABB_coder.py:86-137
  OBB decode         nerf_rpn/model/coder/midpoint_offset_coder.py:160-223, coder/misc.py:5-43
  filter_proposals   nerf_rpn/model/rpn.py:292-370 (incl. the OBB clip/score misalignment, utils.py:359-367)
  clip / small boxes nerf_rpn/model/utils.py:268-289,344-367
Tie conventions (the reference leaves them to torch.topk / argsort): equal logits -> lower flat index
first; equal scores -> lower candidate position first.
Pinned by tests/golden/decode.npz and tests/golden/rpn_small_*.npz (reference outputs generated here).
"""
import itertools

import numpy as np

from . import box as obox

F = np.float32
PI = 3.141592  # coder/misc.py:3


def _exp(x):
    with np.errstate(all="ignore"):
        return np.exp(x.astype(np.float64)).astype(F)


def sigmoid(x):
    x = np.asarray(x, dtype=F)
    with np.errstate(all="ignore"):
        return (F(1.0) / (F(1.0) + _exp(-x))).astype(F)


def cell_anchors(sizes, aspect_ratios):
    """anchor.py:51-82. Returns (A, 6) fp32 [-w,-h,-d,w,h,d]/2 rounded, in the reference's order."""
    scales = np.asarray(sizes, dtype=F)
    ws, hs, ds = [], [], []
    for ratio in aspect_ratios:
        perms = np.asarray(list(set(itertools.permutations(ratio))), dtype=F)   # same set-iteration order
        ws.append(perms[:, 0]); hs.append(perms[:, 1]); ds.append(perms[:, 2])
    w = (np.concatenate(ws)[:, None] * scales[None, :]).reshape(-1)
    h = (np.concatenate(hs)[:, None] * scales[None, :]).reshape(-1)
    d = (np.concatenate(ds)[:, None] * scales[None, :]).reshape(-1)
    base = np.stack([-w, -h, -d, w, h, d], axis=1).astype(F) / F(2)
    return np.round(base).astype(F)     # torch.round == rint (half to even)


def grid_anchors(cell, grid, stride):
    """anchor.py:98-122: anchors of one level, (gx*gy*gz*A, 6), voxel-major / anchor-minor."""
    gx, gy, gz = grid
    sx = np.arange(gx, dtype=F) * F(stride[0])
    sy = np.arange(gy, dtype=F) * F(stride[1])
    sz = np.arange(gz, dtype=F) * F(stride[2])
    X, Y, Z = np.meshgrid(sx, sy, sz, indexing="ij")
    shifts = np.stack([X.ravel(), Y.ravel(), Z.ravel()] * 2, axis=1)
    return (shifts[:, None, :] + cell[None, :, :]).reshape(-1, 6).astype(F)


def decode_aabb(deltas, anchors):
    d = np.asarray(deltas, dtype=F); an = np.asarray(anchors, dtype=F)
    clip = F(np.log(2000.0))
    out = np.empty((d.shape[0], 6), dtype=F)
    with np.errstate(all="ignore"):
        for k in range(3):
            size = an[:, 3 + k] - an[:, k]
            ctr = an[:, k] + F(0.5) * size
            dw = np.where(d[:, 3 + k] > clip, clip, d[:, 3 + k]).astype(F)
            pc = d[:, k] * size + ctr
            ps = _exp(dw) * size
            half = F(0.5) * ps
            out[:, k] = pc - half
            out[:, 3 + k] = pc + half
    return out


def _clamp(v, lo, hi):
    v = np.where(v < lo, F(lo), v)
    return np.where(v > hi, F(hi), v).astype(F)


def decode_obb(deltas, anchors):
    d = np.asarray(deltas, dtype=F); an = np.asarray(anchors, dtype=F)
    mr = F(np.abs(np.log(16 / 1000)))
    with np.errstate(all="ignore"):
        dw, dh, dd = (_clamp(d[:, i], -mr, mr) for i in (3, 4, 5))
        px = (an[:, 0] + an[:, 3]) * F(0.5); py = (an[:, 1] + an[:, 4]) * F(0.5); pz = (an[:, 2] + an[:, 5]) * F(0.5)
        pw = an[:, 3] - an[:, 0]; ph = an[:, 4] - an[:, 1]; pd = an[:, 5] - an[:, 2]
        gw = pw * _exp(dw); gh = ph * _exp(dh); gd = pd * _exp(dd)
        gx = px + pw * d[:, 0]; gy = py + ph * d[:, 1]; gz = pz + pd * d[:, 2]
        hw = gw * F(0.5); hh = gh * F(0.5)
        x1 = gx - hw; y1 = gy - hh; x2 = gx + hw; y2 = gy + hh
        da = _clamp(d[:, 6], -0.5, 0.5); db = _clamp(d[:, 7], -0.5, 0.5)
        ga = gx + da * gw; ga_ = gx - da * gw; gb = gy + db * gh; gb_ = gy - db * gh
        qx = np.stack([ga, x2, ga_, x1], 1); qy = np.stack([y1, gb, y2, gb_], 1)
        cx = qx - gx[:, None]; cy = qy - gy[:, None]
        dl = np.sqrt(cx * cx + cy * cy).astype(F)
        mx = dl.max(axis=1, keepdims=True)
        sc = (mx / dl).astype(F)
        qx = (cx * sc + gx[:, None]).astype(F); qy = (cy * sc + gy[:, None]).astype(F)
        ty = -(qy[:, 1] - qy[:, 0]); tx = (qx[:, 1] - qx[:, 0]) + F(1e-7)
        theta = np.arctan2(ty.astype(np.float64), tx.astype(np.float64)).astype(F)
        Cos = np.cos(theta.astype(np.float64)).astype(F); Sin = np.sin(theta.astype(np.float64)).astype(F)
        xm = (((qx[:, 0] + qx[:, 1]) + qx[:, 2]) + qx[:, 3]) / F(4)
        ym = (((qy[:, 0] + qy[:, 1]) + qy[:, 2]) + qy[:, 3]) / F(4)
        ux = qx - xm[:, None]; uy = qy - ym[:, None]
        rx = ux * Cos[:, None] + uy * (-Sin)[:, None]
        ry = ux * Sin[:, None] + uy * Cos[:, None]
        w = rx.max(1) - rx.min(1); h = ry.max(1) - ry.min(1)
        wh = w > h
        wr = np.where(wh, w, h); hr = np.where(wh, h, w)
        th = np.where(wh, theta, theta + F(PI / 2)).astype(F)
        start = F(-PI / 2)
        th = th - start
        md = np.fmod(th, F(PI)).astype(F)
        md = np.where((md != 0) & (md < 0), md + F(PI), md).astype(F)
        th = md + start
    return np.stack([xm, ym, gz, wr, hr, gd, th], 1).astype(F)


def topk_stable(values, k):
    """indices of the k largest values, ordered by (value desc, index asc)."""
    v = np.asarray(values, dtype=F)
    k = min(k, v.shape[0])
    order = np.lexsort((np.arange(v.shape[0]), -v.astype(np.float64)))
    return order[:k]


def rpn_proposals(logits, deltas, grids, strides, cells, mesh, rotated, pre_nms_top_n=2500, post_nms_top_n=2500,
                  nms_thresh=0.3, score_thresh=0.0, min_size=1e-3, valid=None):
    """rpn.py:303-370 for one scene.

    logits[l]: (V_l*A,) fp32 in flat anchor order ((x*gy + y)*gz + z)*A + a;  deltas[l]: (V_l*A, code).
    Returns boxes (K, 6|7), scores (K,), levels (K,) float.
    """
    code = 8 if rotated else 6
    cand_box, cand_logit, cand_lvl = [], [], []
    for l, (lg, dl) in enumerate(zip(logits, deltas)):
        lg = np.asarray(lg, dtype=F).copy()
        gx, gy, gz = grids[l]
        A = cells[l].shape[0]
        if valid is not None:   # padded batch: anchors in padded voxels get -inf (anchor.py:124-152, rpn.py:321-322)
            lim = [int(np.ceil(valid[i] / strides[l][i])) for i in range(3)]
            m = np.zeros((gx, gy, gz, A), dtype=bool)
            m[:lim[0], :lim[1], :lim[2]] = True
            lg[~m.reshape(-1)] = -np.inf
        idx = topk_stable(lg, pre_nms_top_n)
        vox, a = idx // A, idx % A
        ix, iy, iz = vox // (gy * gz), (vox // gz) % gy, vox % gz
        shift = np.stack([ix * strides[l][0], iy * strides[l][1], iz * strides[l][2]], 1).astype(F)
        anchors = (np.concatenate([shift, shift], 1) + cells[l][a]).astype(F)
        d = np.asarray(dl, dtype=F).reshape(-1, code)[idx]
        cand_box.append(decode_obb(d, anchors) if rotated else decode_aabb(d, anchors))
        cand_logit.append(lg[idx]); cand_lvl.append(np.full(idx.shape[0], l, dtype=np.int32))
    boxes = np.concatenate(cand_box); lvl = np.concatenate(cand_lvl)
    scores = sigmoid(np.concatenate(cand_logit))
    size = [F(mesh[0]), F(mesh[1]), F(mesh[2])]
    if not rotated:
        for k in range(3):
            boxes[:, k] = np.clip(boxes[:, k], F(0), size[k]); boxes[:, 3 + k] = np.clip(boxes[:, 3 + k], F(0), size[k])
        s_al, l_al = scores, lvl
    else:
        ok = np.ones(boxes.shape[0], dtype=bool)
        for k in range(3):
            ok &= (boxes[:, k] >= 0) & (boxes[:, k] <= size[k])
        boxes = boxes[ok]                       # rows dropped from boxes ONLY (utils.py:359-367) ...
        s_al, l_al = scores, lvl                # ... scores / levels stay index-aligned with the old positions
    m = boxes.shape[0]
    if rotated:
        keep = (boxes[:, 3] >= F(min_size)) & (boxes[:, 4] >= F(min_size)) & (boxes[:, 5] >= F(min_size))
    else:
        keep = ((boxes[:, 3] - boxes[:, 0]) >= F(min_size)) & ((boxes[:, 4] - boxes[:, 1]) >= F(min_size)) & \
               ((boxes[:, 5] - boxes[:, 2]) >= F(min_size))
    pos = np.nonzero(keep)[0]
    b2, s2, l2 = boxes[pos], s_al[pos], l_al[pos]     # positions index the (unshortened) score / level arrays
    k2 = np.nonzero(s2 >= F(score_thresh))[0]
    b2, s2, l2 = b2[k2], s2[k2], l2[k2]
    kept = obox.batched_nms(b2, s2, l2, nms_thresh)[:post_nms_top_n]
    return b2[kept], s2[kept], l2[kept].astype(F)
