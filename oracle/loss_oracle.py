"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

numpy fp32 restatement of the loss side of the reference's RPN training step -- the checker for the device
kernels of the next round (the product has target assignment, dgrad, wgrad, bias-grad and ReLU-backward; sampler / losses are
not built yet):
  BalancedPositiveNegativeSampler   utils.py:35-96      (the two torch.randperm draws are inputs: they are the reference's RNG)
  encode_boxes_3d / AABBCoder       coder/AABB_coder.py:14-56
  bbox2delta_sp (midpoint offset)   coder/midpoint_offset_coder.py:106-158, coder/misc.py:46-58,76-83
  smooth-L1 (beta 1/9, sum) / BCE-with-logits (mean) as combined in compute_loss, rpn.py:394-417
Pinned by tests/golden/loss_small.npz (outputs of the reference's own functions, tools/make_golden.py:gen_losses)."""
import numpy as np

f32 = np.float32


def sample(labels, batch_size, positive_fraction, perm_pos, perm_neg):
    """labels: (N,) in {1, 0, -1}; perm_pos / perm_neg: the permutations torch.randperm returned for the positive / negative sets.
    Returns (pos_mask, neg_mask) like the reference's uint8 masks."""
    positive = np.nonzero(labels >= 1)[0]
    negative = np.nonzero(labels == 0)[0]
    num_pos = min(positive.shape[0], int(batch_size * positive_fraction))
    num_neg = min(negative.shape[0], batch_size - num_pos)
    pos = np.zeros(labels.shape[0], dtype=bool)
    neg = np.zeros(labels.shape[0], dtype=bool)
    pos[positive[np.asarray(perm_pos[:num_pos], dtype=np.int64)]] = True
    neg[negative[np.asarray(perm_neg[:num_neg], dtype=np.int64)]] = True
    return pos, neg


def encode_aabb(gt, anchors):
    """encode_boxes_3d(reference_boxes = gt, proposals = anchors): (dx,dy,dz,dw,dh,dd)."""
    g, a = np.asarray(gt, f32), np.asarray(anchors, f32)
    ew, eh, ed = a[:, 3] - a[:, 0], a[:, 4] - a[:, 1], a[:, 5] - a[:, 2]
    ex, ey, ez = a[:, 0] + f32(0.5) * ew, a[:, 1] + f32(0.5) * eh, a[:, 2] + f32(0.5) * ed
    gw, gh, gd = g[:, 3] - g[:, 0], g[:, 4] - g[:, 1], g[:, 5] - g[:, 2]
    gx, gy, gz = g[:, 0] + f32(0.5) * gw, g[:, 1] + f32(0.5) * gh, g[:, 2] + f32(0.5) * gd
    return np.stack([(gx - ex) / ew, (gy - ey) / eh, (gz - ez) / ed, np.log(gw / ew), np.log(gh / eh), np.log(gd / ed)], 1).astype(f32)


def _obb2poly(b):
    x, y, w, h, th = (b[:, i] for i in range(5))
    co, si = np.cos(th).astype(f32), np.sin(th).astype(f32)
    v1x, v1y = (w / f32(2)) * co, -(w / f32(2)) * si
    v2x, v2y = -(h / f32(2)) * si, -(h / f32(2)) * co
    p = [x + v1x + v2x, y + v1y + v2y, x + v1x - v2x, y + v1y - v2y, x - v1x - v2x, y - v1y - v2y, x - v1x + v2x, y - v1y + v2y]
    return np.stack(p, 1).astype(f32)


def _obb2hbb(b):
    x, y, w, h, th = (b[:, i] for i in range(5))
    co, si = np.cos(th).astype(f32), np.sin(th).astype(f32)
    xb = np.abs((w / f32(2)) * co) + np.abs((h / f32(2)) * si)
    yb = np.abs((w / f32(2)) * si) + np.abs((h / f32(2)) * co)
    return np.stack([x - xb, y - yb, x + xb, y + yb], 1).astype(f32)


def encode_obb_midpoint(anchors, gt):
    """bbox2delta_sp(proposals = anchors (N,6), gt (N,7)): (dx,dy,dz,dw,dh,dd,da,db)."""
    p, g = np.asarray(anchors, f32), np.asarray(gt, f32)
    px, py, pz = (p[:, 0] + p[:, 3]) * f32(0.5), (p[:, 1] + p[:, 4]) * f32(0.5), (p[:, 2] + p[:, 5]) * f32(0.5)
    pw, ph, pd = p[:, 3] - p[:, 0], p[:, 4] - p[:, 1], p[:, 5] - p[:, 2]
    gz, gd = g[:, 2], g[:, 5]
    g2 = np.stack([g[:, 0], g[:, 1], g[:, 3], g[:, 4], g[:, 6]], 1).astype(f32)
    hbb, poly = _obb2hbb(g2), _obb2poly(g2)
    gx, gy = (hbb[:, 0] + hbb[:, 2]) * f32(0.5), (hbb[:, 1] + hbb[:, 3]) * f32(0.5)
    gw, gh = hbb[:, 2] - hbb[:, 0], hbb[:, 3] - hbb[:, 1]
    xc, yc = poly[:, 0::2], poly[:, 1::2]
    ymin, xmax = yc.min(axis=1, keepdims=True), xc.max(axis=1, keepdims=True)
    _x = xc.copy(); _x[np.abs(yc - ymin) > f32(0.1)] = f32(-1000)
    ga = _x.max(axis=1)
    _y = yc.copy(); _y[np.abs(xc - xmax) > f32(0.1)] = f32(-1000)
    gb = _y.max(axis=1)
    return np.stack([(gx - px) / pw, (gy - py) / ph, (gz - pz) / pd, np.log(gw / pw), np.log(gh / ph), np.log(gd / pd),
                     (ga - gx) / gw, (gb - gy) / gh], 1).astype(f32)


def smooth_l1_sum(pred, target, beta):
    d = np.abs(np.asarray(pred, f32) - np.asarray(target, f32))
    b = f32(beta)
    return np.where(d < b, f32(0.5) * d * d / b, d - f32(0.5) * b).astype(f32).sum(dtype=np.float64)


def bce_with_logits_mean(logits, labels):
    x, y = np.asarray(logits, np.float64), np.asarray(labels, np.float64)
    return float(np.mean(np.maximum(x, 0) - x * y + np.log1p(np.exp(-np.abs(x)))))


def rpn_losses_3d(objectness, pred_deltas, labels, regression_targets, pos_mask, neg_mask):
    """loss_objectness and loss_rpn_box_reg (smooth-L1 branch) of compute_loss (rpn.py:394-417)."""
    pos, neg = np.nonzero(pos_mask)[0], np.nonzero(neg_mask)[0]
    sampled = np.concatenate([pos, neg])
    box = smooth_l1_sum(pred_deltas[pos], regression_targets[pos], 1.0 / 9.0) / float(sampled.shape[0])
    obj = bce_with_logits_mean(np.asarray(objectness).reshape(-1)[sampled], labels[sampled])
    return obj, box
