"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

numpy fp32 restatement of the reference's anchor/ground-truth assignment:
  obb2hbb_3d                       coder/misc.py:85-93
  box_iou_3d (AABB branch)         utils.py:418-458  (boxes1 = ground truth, boxes2 = anchors)
  Matcher + low-quality matches    utils.py:142-212
  label mapping / padding mask     rpn.py:260-288
Pinned by tests/golden/targets_small.npz (outputs of the reference's own functions, tools/make_golden.py:gen_targets)."""
import numpy as np

f32 = np.float32


def obb2hbb_3d(b):
    b = np.asarray(b, dtype=f32)
    x, y, z, w, h, d, th = (b[:, i] for i in range(7))
    co, si = np.cos(th.astype(np.float64)).astype(f32), np.sin(th.astype(np.float64)).astype(f32)
    hw, hh = (w / f32(2)).astype(f32), (h / f32(2)).astype(f32)
    xb = (np.abs(hw * co) + np.abs(hh * si)).astype(f32)
    yb = (np.abs(hw * si) + np.abs(hh * co)).astype(f32)
    zb = (d / f32(2)).astype(f32)
    return np.stack([x - xb, y - yb, z - zb, x + xb, y + yb, z + zb], 1).astype(f32)


def aabb_iou(gt, anchors):
    gt, anchors = np.asarray(gt, f32), np.asarray(anchors, f32)
    v1 = ((gt[:, 3] - gt[:, 0]) * (gt[:, 4] - gt[:, 1]) * (gt[:, 5] - gt[:, 2])).astype(f32)
    v2 = ((anchors[:, 3] - anchors[:, 0]) * (anchors[:, 4] - anchors[:, 1]) * (anchors[:, 5] - anchors[:, 2])).astype(f32)
    lt = np.maximum(gt[:, None, :3], anchors[None, :, :3])
    rb = np.minimum(gt[:, None, 3:], anchors[None, :, 3:])
    whd = np.clip((rb - lt).astype(f32), 0, None)
    inter = (whd[..., 0] * whd[..., 1]).astype(f32) * whd[..., 2]
    union = ((v1[:, None] + v2[None, :]).astype(f32) - inter).astype(f32)
    return (inter / union).astype(f32)


def assign(anchors, gt, valid, high, low, allow_low=True):
    gq = obb2hbb_3d(gt) if gt.shape[1] == 7 else np.asarray(gt, f32)
    m = aabb_iou(gq, anchors)
    if valid is not None:
        m[:, ~valid] = f32(-1.0)
    vals, idx = m.max(axis=0), m.argmax(axis=0).astype(np.int64)
    allm = idx.copy()
    idx[vals < f32(low)] = -1
    idx[(vals >= f32(low)) & (vals < f32(high))] = -2
    if allow_low:
        best = m.max(axis=1)
        upd = np.where(m == best[:, None])[1]
        idx[upd] = allm[upd]
    labels = (idx >= 0).astype(f32)
    labels[idx == -1] = 0.0
    labels[idx == -2] = -1.0
    if valid is not None:
        labels[~valid] = -1.0
    return labels, idx
