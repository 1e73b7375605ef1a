"""CPU restatement of the FCOS training loss -- TEST INFRASTRUCTURE (the checker of csrc/fcos_loss.cu; never imported by nerf_rpn_b200/).

Follows nerf_rpn/model/fcos/loss.py of the reference:
  gt_prepare                 compute_targets_for_locations_obb :331-340 + encode_fcos_obb (fcos/utils.py:64-108) + box2corners_th
                             (rotated_iou/oriented_iou_loss.py:6-36): per ground-truth box, the location-independent part
  targets                    prepare_targets :262-316, compute_targets_for_locations[_obb] :318-441, get_sample_region :213-260
  centerness_targets         compute_centerness_targets :443-450
  aabb_iou_losses            IOULoss.forward :78-131 (per box, before the weighted sum)
  projection_loss_2d         compute_2d_projection_loss :452-485 (+ decode_fcos_obb, get_w2cs, project, obb2points_3d of fcos/utils.py)
  loss                       FCOSLossComputation.__call__ :487-591 (single rank: world_size 1), AABB head or OBB head with smooth-L1;
                             the rotated-IoU term of the OBB head (RotatedIOULoss :134-181) needs the reference's CUDA-only vertex sort and
                             is checked against the staged reference itself on the GPU box (tests/test_gpu_fcos_loss.py).
Pinned by tests/golden/fcos_loss.npz = outputs of the unmodified reference (tools/make_golden.py gen_fcos_loss), tests/test_fcos_loss_cpu.py.
numpy for the integer / comparison work, torch (CPU, fp32, autograd) for the differentiable sums.
"""
import numpy as np
import torch

INF = np.float32(100000000.0)
SIZES_OF_INTEREST = ((-1.0, 16.0), (16.0, 32.0), (32.0, 64.0), (64.0, 100000000.0))          # loss.py:263-268


def compute_locations(grids, strides):
    """fcos.py:221-250: per level (w*l*h, 3) fp32, z fastest, idx * stride + stride // 2."""
    out = []
    for (w, l, h), s in zip(grids, strides):
        ax = [np.arange(0, n * s, s, dtype=np.float32) for n in (w, l, h)]
        g = np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 3)
        out.append((g + np.float32(s // 2)).astype(np.float32))
    return out


def gt_prepare(gt):
    """(G, 6|7) -> lo (G,3), hi (G,3), alpha (G), beta (G), volume (G), all fp32."""
    gt = np.asarray(gt, np.float32)
    if gt.shape[1] == 6:
        lo, hi = gt[:, :3].copy(), gt[:, 3:].copy()
        alpha = np.zeros(len(gt), np.float32); beta = np.zeros(len(gt), np.float32)
    else:
        x, y, z, w, h, d, th = (gt[:, k] for k in range(7))
        co, si = np.cos(th.astype(np.float64)).astype(np.float32), np.sin(th.astype(np.float64)).astype(np.float32)
        sx = np.array([0.5, -0.5, -0.5, 0.5], np.float32); sy = np.array([0.5, 0.5, -0.5, -0.5], np.float32)
        x4, y4 = sx[None] * w[:, None], sy[None] * h[:, None]
        xs = (x4 * co[:, None] + y4 * (-si[:, None])).astype(np.float32) + x[:, None]
        ys = (x4 * si[:, None] + y4 * co[:, None]).astype(np.float32) + y[:, None]
        xmax, xmin, ymax, ymin = xs.max(1), xs.min(1), ys.max(1), ys.min(1)
        xt = np.where(ymax[:, None] - ys > np.float32(0.1), np.float32(-1e6), xs)
        yt = np.where(xmax[:, None] - xs > np.float32(0.1), np.float32(1e6), ys)
        vx, vy = xt.max(1), yt.min(1)
        ids = np.isclose(vx, xmax, rtol=1e-5, atol=1e-8) & np.isclose(vy, ymin, rtol=1e-5, atol=1e-8)
        vx = np.where(ids, xmax, vx); vy = np.where(ids, ymin, vy)
        alpha = ((vx - x) / (xmax - xmin)).astype(np.float32)
        beta = ((vy - y) / (ymax - ymin)).astype(np.float32)
        lo = np.stack([xmin, ymin, z - d / np.float32(2)], 1).astype(np.float32)
        hi = np.stack([xmax, ymax, z + d / np.float32(2)], 1).astype(np.float32)
    vol = ((hi[:, 0] - lo[:, 0]) * (hi[:, 1] - lo[:, 1]) * (hi[:, 2] - lo[:, 2])).astype(np.float32)
    return lo, hi, alpha, beta, vol


def targets(locations, strides, gt, center_sampling_radius, norm_reg_targets):
    """One scene: locations = list of (P_l, 3) per level, gt (G, 6|7) -> labels (P) f32, reg_targets (P, 6|8) f32, levels concatenated."""
    pts = np.concatenate(locations, 0).astype(np.float32)
    n_per = [len(p) for p in locations]
    gt = np.asarray(gt, np.float32)
    D = 8 if gt.shape[1] == 7 else 6
    if gt.shape[0] == 0:
        return np.zeros(len(pts), np.float32), np.zeros((len(pts), D), np.float32)
    lo, hi, alpha, beta, vol = gt_prepare(gt)
    reg = np.concatenate([pts[:, None, :] - lo[None], hi[None] - pts[:, None, :]], axis=2).astype(np.float32)          # (P, G, 6)
    lvl = np.repeat(np.arange(len(n_per)), n_per)
    if center_sampling_radius > 0:
        srad = np.array([np.float32(s * center_sampling_radius) for s in strides], np.float32)[lvl][:, None, None]      # (P,1,1)
        c = ((lo + hi) / np.float32(2)).astype(np.float32)[None]                                                         # (1,G,3)
        cmin, cmax = c - srad, c + srad
        clo = np.where(cmin > lo[None], cmin, lo[None]); chi = np.where(cmax > hi[None], hi[None], cmax)
        dist = np.concatenate([pts[:, None, :] - clo, chi - pts[:, None, :]], axis=2)
        inside = dist.min(2) > 0
    else:
        inside = reg.min(2) > 0
    mx = reg.max(2)
    soi = np.array(SIZES_OF_INTEREST, np.float32)[lvl]
    cared = (mx >= soi[:, :1]) & (mx <= soi[:, 1:])
    area = np.where(inside & cared, vol[None], INF).astype(np.float32)
    idx = area.argmin(1)                                                   # first minimum
    labels = (area[np.arange(len(pts)), idx] != INF).astype(np.float32)
    out = np.concatenate([reg[np.arange(len(pts)), idx], alpha[idx][:, None], beta[idx][:, None]], 1)[:, :D].astype(np.float32)
    if norm_reg_targets:
        out[:, :6] = out[:, :6] / np.array(strides, np.float32)[lvl][:, None]
    return labels, out


def centerness_targets(rt: torch.Tensor) -> torch.Tensor:
    lr, tb, fb = rt[:, [0, 3]], rt[:, [1, 4]], rt[:, [2, 5]]
    return torch.sqrt((lr.min(1)[0] / lr.max(1)[0]) * (tb.min(1)[0] / tb.max(1)[0]) * (fb.min(1)[0] / fb.max(1)[0]))


def aabb_iou_losses(pred: torch.Tensor, tgt: torch.Tensor, loss_type: str) -> torch.Tensor:
    pv = (pred[:, 0] + pred[:, 3]) * (pred[:, 1] + pred[:, 4]) * (pred[:, 2] + pred[:, 5])
    tv = (tgt[:, 0] + tgt[:, 3]) * (tgt[:, 1] + tgt[:, 4]) * (tgt[:, 2] + tgt[:, 5])
    inter = [torch.min(pred[:, k], tgt[:, k]) + torch.min(pred[:, k + 3], tgt[:, k + 3]) for k in range(3)]
    outer = [torch.max(pred[:, k], tgt[:, k]) + torch.max(pred[:, k + 3], tgt[:, k + 3]) for k in range(3)]
    ac = outer[0] * outer[1] * outer[2] + 1e-7
    vi = inter[0] * inter[1] * inter[2]
    vu = tv + pv - vi
    iou = (vi + 1.0) / (vu + 1.0)
    if loss_type == "iou":
        return -torch.log(iou)
    if loss_type == "linear_iou":
        return 1 - iou
    if loss_type == "giou":
        return 1 - (iou - (ac - vu) / ac)
    raise NotImplementedError(loss_type)


def focal_sum(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """torchvision.ops.sigmoid_focal_loss(alpha 0.25, gamma 2, reduction sum)."""
    p = torch.sigmoid(logits)
    ce = torch.nn.functional.binary_cross_entropy_with_logits(logits, labels, reduction="none")
    pt = p * labels + (1 - p) * (1 - labels)
    return ((0.25 * labels + 0.75 * (1 - labels)) * (ce * (1 - pt) ** 2)).sum()


def flatten_level_first(per_level, channels):
    """list of (N, C, w, l, h) -> (sum_l N * P_l, C): loss.py:511-518 (level by level, scenes inside a level)."""
    return torch.cat([t.permute(0, 2, 3, 4, 1).reshape(-1, channels) for t in per_level], 0)


def decode_obb(reg: torch.Tensor) -> torch.Tensor:
    """decode_fcos_obb (fcos/utils.py:12-61) at location 0: (K, 8) -> (K, 7)."""
    x0, y0, z0, x1, y1, z1 = -reg[:, 0], -reg[:, 1], -reg[:, 2], reg[:, 3], reg[:, 4], reg[:, 5]
    vx = torch.clamp((x1 + x0) / 2 + reg[:, 6] * (x1 - x0), min=x0, max=x1)
    vy = torch.clamp((y1 + y0) / 2 + reg[:, 7] * (y1 - y0), min=y0, max=y1)
    c = torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, (z0 + z1) / 2], 1)
    v0, v1 = torch.stack([vx, y1], 1) - c[:, :2], torch.stack([x1, vy], 1) - c[:, :2]
    d0, d1 = v0.norm(dim=1), v1.norm(dim=1)
    dm = torch.max(d0, d1)
    v0 = v0 / (d0[:, None] + 1e-7) * dm[:, None] + c[:, :2]
    v1 = v1 / (d1[:, None] + 1e-7) * dm[:, None] + c[:, :2]
    mid = (v0 + v1) / 2 - c[:, :2]
    mx = torch.where((mid[:, 0] == 0) & (mid[:, 1] == 0), torch.full_like(mid[:, 0], 1e-7), mid[:, 0])
    return torch.stack([c[:, 0], c[:, 1], c[:, 2], mid.norm(dim=1) * 2, (v0 - v1).norm(dim=1), z1 - z0, torch.atan2(mid[:, 1], mx)], 1)


def projection_2d(points: torch.Tensor, res=160.0) -> torch.Tensor:
    """get_w2cs + project (fcos/utils.py:300-377): (M, 3) -> (4 M, 2) pixels in the four corner cameras of a `res` scene."""
    K = torch.tensor([[600.0, 0, 320.0], [0, 600.0, 240.0], [0, 0, 1.0]])
    ctr = np.full(3, res / 2)
    out = []
    for dx, dy in ((res, res), (res, -res), (-res, res), (-res, -res)):
        cam = ctr + np.array([dx, dy, res])
        zax = (cam - ctr) / np.linalg.norm(cam - ctr)
        xax = np.cross([0.0, 0.0, 1.0], zax); xax /= np.linalg.norm(xax)
        yax = np.cross(zax, xax); yax /= np.linalg.norm(yax)
        c2w = np.eye(4); c2w[:3, :3] = np.stack([xax, yax, zax], 1); c2w[:3, 3] = cam
        w2c = torch.tensor(np.linalg.inv(c2w), dtype=torch.float32)
        camc = w2c @ torch.cat([points, torch.ones(len(points), 1)], 1).t()
        pic = K @ camc[:3]
        out.append((pic[:2] / pic[2]).t())
    return torch.cat(out, 0)


def projection_loss_2d(reg: torch.Tensor, rt: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """compute_2d_projection_loss (fcos/loss.py:452-485)."""
    def pts(b):
        v = torch.stack([b[:, 3] / 2 * torch.cos(b[:, 6]) - b[:, 4] / 2 * torch.sin(b[:, 6]), b[:, 3] / 2 * torch.sin(b[:, 6]) + b[:, 4] / 2 * torch.cos(b[:, 6]),
                         b[:, 5] / 2], 1)
        return torch.cat([b[:, :3] - v, b[:, :3] + v], 0)
    l2 = torch.nn.functional.smooth_l1_loss(projection_2d(pts(decode_obb(reg))), projection_2d(pts(decode_obb(rt))), beta=1 / 9, reduction="none") / 160
    factor = l2.shape[0] // weights.shape[0]
    return (l2 * weights[:, None].repeat(factor, 1)).sum() / (factor * l2.shape[1])


def loss(box_cls, box_regression, centerness, labels, reg_targets, masks, iou_loss_type, use_obb=False, use_additional_l1_loss=False,
         proj2d_loss_weight=0.0):
    """box_cls / box_regression / centerness: lists of (N, C, w, l, h) torch tensors (may require grad); labels / reg_targets: per scene
    (P,) / (P, D) numpy from `targets`; masks: None or list per level of (N, P_l) bool.  -> (loss_cls, loss_reg, loss_centerness) and the
    raw sums dict.  OBB with an IoU-type loss returns loss_reg WITHOUT the rotated-IoU term (see the module docstring)."""
    D = 8 if use_obb else 6
    n_per = [int(np.prod(t.shape[2:])) for t in box_cls]
    N = box_cls[0].shape[0]
    cls = flatten_level_first(box_cls, 1).reshape(-1)
    reg = flatten_level_first(box_regression, D)
    ctr = flatten_level_first(centerness, 1).reshape(-1)
    lab, rt = [], []
    off = 0
    for pl in n_per:
        lab.append(np.concatenate([labels[n][off:off + pl] for n in range(N)]))
        rt.append(np.concatenate([reg_targets[n][off:off + pl] for n in range(N)]))
        off += pl
    lab = torch.from_numpy(np.concatenate(lab)); rt = torch.from_numpy(np.concatenate(rt))
    if masks is not None:
        m = torch.cat([torch.as_tensor(mm).reshape(-1) for mm in masks])
        cls, reg, ctr, lab, rt = cls[m], reg[m], ctr[m], lab[m], rt[m]
    pos = torch.nonzero(lab > 0).squeeze(1)
    n_pos = max(float(pos.numel()), 1.0)
    sums = {"focal": focal_sum(cls, lab), "n_pos": pos.numel()}
    loss_cls = sums["focal"] / n_pos
    if pos.numel() == 0:
        z = reg.sum() * 0
        return loss_cls, z, z, sums
    reg, rt, ctr = reg[pos], rt[pos], ctr[pos]
    ct = centerness_targets(rt)
    sums["sum_ct"] = ct.sum()
    if iou_loss_type == "smooth_l1":
        reg_raw = (torch.nn.functional.smooth_l1_loss(reg, rt, reduction="none") * ct[:, None]).sum()
    elif not use_obb:
        reg_raw = (aabb_iou_losses(reg, rt, iou_loss_type) * ct).sum()
    else:
        reg_raw = reg.sum() * 0
    sums["reg"] = reg_raw
    loss_reg = reg_raw / ct.sum()
    if use_obb and use_additional_l1_loss and iou_loss_type != "smooth_l1":
        add = (torch.nn.functional.smooth_l1_loss(reg[:, 6:], rt[:, 6:], reduction="none") * ct[:, None]).sum()
        sums["add_l1"] = add
        loss_reg = loss_reg + add / ct.sum()
    if use_obb and proj2d_loss_weight > 0:
        loss_reg = loss_reg + projection_loss_2d(reg, rt, ct) / ct.sum() * proj2d_loss_weight
    sums["bce"] = torch.nn.functional.binary_cross_entropy_with_logits(ctr, ct, reduction="sum")
    return loss_cls, loss_reg, sums["bce"] / n_pos, sums
