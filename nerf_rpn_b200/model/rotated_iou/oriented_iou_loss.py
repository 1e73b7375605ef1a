"""cal_iou_3d. Mirrors nerf_rpn/model/rotated_iou/oriented_iou_loss.py:82-107: same signature, `verbose=True` returns
(iou, corners1, corners2, z_range, u3d), and the result is differentiable w.r.t. both boxes (the reference uses it as a loss:
rpn.py:133-165 RotatedIOULoss, fcos/loss.py).  Forward values come from the fused exact kernel (bit-identical to the reference's chain
on the same GPU); the backward pass is nrpn_iou3d_pairs_backward (intersection volume differentiated in fp64).
cal_giou_3d / cal_diou_3d add the enclosing rectangle (aligned / pca / smallest, min_enclosing_box.py) as differentiable torch ops on top."""
import ctypes

import torch

from ... import ops
from ..._lib import check, lib


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _IoU3D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, box1, box2):
        shape = box1.shape[:-1]
        a = box1.detach().reshape(-1, 7).to(device="cuda", dtype=torch.float32).contiguous()
        b = box2.detach().reshape(-1, 7).to(device="cuda", dtype=torch.float32).contiguous()
        n = a.shape[0]
        iou = torch.empty(n, device="cuda"); zr = torch.empty(n, device="cuda"); u3 = torch.empty(n, device="cuda")
        c1 = torch.empty((n, 4, 2), device="cuda"); c2 = torch.empty((n, 4, 2), device="cuda")
        check(lib().nrpn_iou3d_pairs_verbose(_p(a), _p(b), n, _p(iou), _p(c1), _p(c2), _p(zr), _p(u3), _stream()), "iou3d_pairs_verbose")
        ctx.save_for_backward(a, b)
        ctx.shape, ctx.dev = tuple(box1.shape), (box1.device, box2.device)
        ctx.mark_non_differentiable(c1, c2, zr)
        dev = box1.device
        return (iou.reshape(shape).to(dev), c1.reshape(*shape, 4, 2).to(dev), c2.reshape(*shape, 4, 2).to(dev), zr.reshape(shape).to(dev),
                u3.reshape(shape).to(dev))

    @staticmethod
    def backward(ctx, g_iou, g_c1, g_c2, g_zr, g_u):
        a, b = ctx.saved_tensors
        n = a.shape[0]
        gi = None if g_iou is None else g_iou.reshape(-1).to(device="cuda", dtype=torch.float32).contiguous()
        gu = None if g_u is None else g_u.reshape(-1).to(device="cuda", dtype=torch.float32).contiguous()
        if gi is None and gu is None:
            return None, None
        ga = torch.empty((n, 7), device="cuda"); gb = torch.empty((n, 7), device="cuda")
        check(lib().nrpn_iou3d_pairs_backward(_p(a), _p(b), n, _p(gi), _p(gu), _p(ga), _p(gb), _stream()), "iou3d_pairs_backward")
        return ga.reshape(ctx.shape).to(ctx.dev[0]), gb.reshape(ctx.shape).to(ctx.dev[1])


def cal_iou_3d(box3d1: torch.Tensor, box3d2: torch.Tensor, verbose=False):
    """(B, N, 7) x (B, N, 7) (x, y, z, w, h, l, alpha) -> (B, N) IoU, or the 5-tuple of the reference with verbose=True."""
    if not verbose and not (box3d1.requires_grad or box3d2.requires_grad):
        shape = box3d1.shape[:-1]
        a = box3d1.detach().reshape(-1, 7).to(device="cuda", dtype=torch.float32).contiguous()
        b = box3d2.detach().reshape(-1, 7).to(device="cuda", dtype=torch.float32).contiguous()
        return ops.iou3d_pairs(a, b).reshape(shape).to(box3d1.device)
    iou, c1, c2, zr, u3 = _IoU3D.apply(box3d1, box3d2)
    return (iou, c1, c2, zr, u3) if verbose else iou


# ---------------------------------------------------------------------------------------------- GIoU / DIoU (oriented_iou_loss.py:109-247)
# IoU and union come from the fused kernel (with its backward); what the two losses add -- the boxes' corners, the common z range and the enclosing
# rectangle of the eight corners -- are a few dozen differentiable torch ops on (B, N) tensors, written here from the reference's formulas.
def _corners(box3d):
    """(.., 7) -> (.., 4, 2) footprint corners in the reference's order (box2corners_th, :6-36)."""
    x, y, w, h, alpha = box3d[..., 0:1], box3d[..., 1:2], box3d[..., 3:4], box3d[..., 4:5], box3d[..., 6:7]
    sx = torch.tensor([0.5, -0.5, -0.5, 0.5], dtype=box3d.dtype, device=box3d.device)
    sy = torch.tensor([0.5, 0.5, -0.5, -0.5], dtype=box3d.dtype, device=box3d.device)
    x4, y4 = sx * w, sy * h
    c, s = torch.cos(alpha), torch.sin(alpha)
    return torch.stack([x4 * c - y4 * s + x, x4 * s + y4 * c + y], dim=-1)


def _z_range(b1, b2):
    zmax = torch.max(b1[..., 2] + b1[..., 5] * 0.5, b2[..., 2] + b2[..., 5] * 0.5)
    zmin = torch.min(b1[..., 2] - b1[..., 5] * 0.5, b2[..., 2] - b2[..., 5] * 0.5)
    return (zmax - zmin).clamp_min(0.0)


_HULL_EDGES = [(i, j) for i in range(8) for j in range(i + 1, 8) if (i, j) not in ((0, 2), (1, 3), (4, 6), (5, 7))]      # 24 candidates


def _enclosing_smallest(pts):
    """Smallest enclosing rectangle of 8 points by brute force (min_enclosing_box.py:118-167): a side of the minimum-area rectangle is collinear
    with an edge of the hull, so every point pair that can be a hull edge is tried (the 4 box diagonals cannot)."""
    dev = pts.device
    ii = torch.tensor([e[0] for e in _HULL_EDGES], device=dev)
    jj = torch.tensor([e[1] for e in _HULL_EDGES], device=dev)
    rest = torch.tensor([[k for k in range(8) if k not in e] for e in _HULL_EDGES], device=dev)            # (24, 6)
    p1, p2 = pts[..., ii, :], pts[..., jj, :]                                    # (.., 24, 2)
    x1, y1, x2, y2 = p1[..., 0:1], p1[..., 1:2], p2[..., 0:1], p2[..., 1:2]      # (.., 24, 1)
    others = pts[..., rest, :]                                                   # (.., 24, 6, 2)
    # extent ALONG the candidate edge: projection of all 8 points on (1, k), k = slope with the reference's 1e-8 guard
    k = (y2 - y1) / (x2 - x1 + 1e-8)
    vec = torch.cat([torch.ones_like(k), k], dim=-1)                             # (.., 24, 2)
    allp = torch.cat([p1.unsqueeze(-2), p2.unsqueeze(-2), others], dim=-2)       # (.., 24, 8, 2)
    proj = (allp * vec.unsqueeze(-2)).sum(-1) / torch.norm(vec, dim=-1, keepdim=True)
    along = proj.max(-1)[0] - proj.min(-1)[0]
    # extent ACROSS it: signed point-line distances of the other six points
    den = (y2 - y1) * others[..., 0] - (x2 - x1) * others[..., 1] + x2 * y1 - y2 * x1
    d = den / torch.sqrt((y2 - y1).square() + (x2 - x1).square() + 1e-14)
    across = torch.max(d.max(-1)[0] - d.min(-1)[0], d.abs().max(-1)[0])
    area = along * across
    area = area + (area == 0).to(area.dtype) * 1e8                               # coincident end points: not a candidate
    idx = area.min(dim=-1, keepdim=True)[1]
    return along.gather(-1, idx).squeeze(-1), across.gather(-1, idx).squeeze(-1)


def enclosing_box(corners1, corners2, enclosing_type="smallest"):
    """(w, h) of the rectangle enclosing both footprints (oriented_iou_loss.py:156-246): "aligned" (axis-aligned), "pca" (axes = principal
    components of the 8 corners, closed-form 2x2 eigenvectors in fp64) or "smallest" (minimum area over the hull-edge directions)."""
    pts = torch.cat([corners1, corners2], dim=-2)
    if enclosing_type == "aligned":
        return pts[..., 0].max(-1)[0] - pts[..., 0].min(-1)[0], pts[..., 1].max(-1)[0] - pts[..., 1].min(-1)[0]
    if enclosing_type == "pca":
        c = pts - pts.mean(dim=-2, keepdim=True)
        m = c.transpose(-1, -2) @ c
        a, cc, b = m[..., 0, 0].double(), m[..., 0, 1].double(), m[..., 1, 1].double()
        delta = torch.sqrt(a * a + 4 * cc * cc - 2 * a * b + b * b)
        vs = []
        for sign in (-1.0, 1.0):
            v = torch.stack([(a - b + sign * delta) / 2.0 / cc, torch.ones_like(a)], dim=-1)
            vs.append((v / v.norm(dim=-1, keepdim=True)).to(pts.dtype))
        p1, p2 = (c * vs[0].unsqueeze(-2)).sum(-1), (c * vs[1].unsqueeze(-2)).sum(-1)
        return p1.max(-1)[0] - p1.min(-1)[0], p2.max(-1)[0] - p2.min(-1)[0]
    if enclosing_type == "smallest":
        return _enclosing_smallest(pts)
    raise ValueError("Unknown type enclosing. Supported: aligned, pca, smallest")


def _iou_union(box3d1, box3d2):
    iou, _, _, _, u3d = _IoU3D.apply(box3d1, box3d2)
    return iou, u3d


def cal_giou_3d(box3d1: torch.Tensor, box3d2: torch.Tensor, enclosing_type: str = "smallest"):
    """-> (giou_loss, giou, iou3d), each (B, N) (oriented_iou_loss.py:109-127)."""
    iou3d, u3d = _iou_union(box3d1, box3d2)
    w, h = enclosing_box(_corners(box3d1), _corners(box3d2), enclosing_type)
    v_c = _z_range(box3d1, box3d2) * w * h
    giou_loss = 1.0 - iou3d + (v_c - u3d) / v_c
    return giou_loss, 1 - giou_loss, iou3d


def cal_diou_3d(box3d1: torch.Tensor, box3d2: torch.Tensor, enclosing_type: str = "smallest"):
    """-> (diou_loss, iou3d), each (B, N) (oriented_iou_loss.py:129-150)."""
    iou3d, _ = _iou_union(box3d1, box3d2)
    w, h = enclosing_box(_corners(box3d1), _corners(box3d2), enclosing_type)
    z_range = _z_range(box3d1, box3d2)
    off = box3d1[..., :3] - box3d2[..., :3]
    d2 = (off * off).sum(-1)
    c2 = w * w + h * h + z_range * z_range
    return 1.0 - iou3d + d2 / c2, iou3d
