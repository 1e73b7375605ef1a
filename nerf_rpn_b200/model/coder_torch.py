"""Differentiable box decoding for the IoU-type regression losses (training only): MidpointOffsetCoder.decode of the reference
(coder/midpoint_offset_coder.py:34-46, 163-215 `delta_sp2bbox`; coder/misc.py:5-43 `regular_theta`, `regular_obb`, `rectpoly2obb`) written with
torch ops so that autograd reaches the head's deltas.  Inference decodes on the device inside nrpn_rpn_proposals (csrc/rpn_decode.cuh: same formulas,
the reference's rounding order); this module is only evaluated on the <= 128 sampled positive anchors of a training step."""
import math

import torch

_MAX_RATIO = abs(math.log(16 / 1000))
_PI = 3.141592                      # the reference's literal (misc.py:8)


_XFORM_CLIP = math.log(2000.0)      # AABBCoder.bbox_xform_clip (coder/AABB_coder.py:66)


def decode_aabb(anchors: torch.Tensor, deltas: torch.Tensor) -> torch.Tensor:
    """AABBCoder.decode_single (coder/AABB_coder.py:86-137), differentiable: anchors (P,6), deltas (P,6) (dx,dy,dz,dw,dh,dd) -> (P,6) corners.
    Used by the 2-D projection loss of the axis-aligned head (model/proj2d.py); inference decodes in csrc/rpn_decode.cuh."""
    size = anchors[:, 3:] - anchors[:, :3]
    ctr = anchors[:, :3] + 0.5 * size
    centre = deltas[:, :3] * size + ctr
    half = 0.5 * (torch.exp(deltas[:, 3:].clamp(max=_XFORM_CLIP)) * size)
    return torch.cat([centre - half, centre + half], 1)


def decode_obb(anchors: torch.Tensor, deltas: torch.Tensor) -> torch.Tensor:
    """anchors (P,6) [x1,y1,z1,x2,y2,z2], deltas (P,8) (dx,dy,dz,dw,dh,dd,da,db) -> (P,7) (x,y,z,w,h,d,theta)."""
    dx, dy, dz, dw, dh, dd, da, db = deltas.unbind(1)
    dw, dh, dd = (t.clamp(-_MAX_RATIO, _MAX_RATIO) for t in (dw, dh, dd))
    px, py, pz = (anchors[:, 0] + anchors[:, 3]) * 0.5, (anchors[:, 1] + anchors[:, 4]) * 0.5, (anchors[:, 2] + anchors[:, 5]) * 0.5
    pw, ph, pd = anchors[:, 3] - anchors[:, 0], anchors[:, 4] - anchors[:, 1], anchors[:, 5] - anchors[:, 2]
    gw, gh, gd = pw * dw.exp(), ph * dh.exp(), pd * dd.exp()
    gx, gy, gz = px + pw * dx, py + ph * dy, pz + pd * dz
    x1, y1, x2, y2 = gx - gw * 0.5, gy - gh * 0.5, gx + gw * 0.5, gy + gh * 0.5
    da, db = da.clamp(-0.5, 0.5), db.clamp(-0.5, 0.5)
    ga, ga_, gb, gb_ = gx + da * gw, gx - da * gw, gy + db * gh, gy - db * gh
    qx = torch.stack([ga, x2, ga_, x1], 1)                      # the four midpoint-offset vertices
    qy = torch.stack([y1, gb, y2, gb_], 1)
    cx, cy = qx - gx[:, None], qy - gy[:, None]                 # rectangularise: every centred vertex scaled to the longest diagonal
    dl = (cx * cx + cy * cy).sqrt()
    mx = dl.max(dim=1, keepdim=True)[0]
    qx, qy = cx * (mx / dl) + gx[:, None], cy * (mx / dl) + gy[:, None]
    theta = torch.atan2(-(qy[:, 1] - qy[:, 0]), qx[:, 1] - qx[:, 0] + 1e-7)          # rectpoly2obb
    cos, sin = theta.cos(), theta.sin()
    xm, ym = qx.sum(1) / 4.0, qy.sum(1) / 4.0
    ux, uy = qx - xm[:, None], qy - ym[:, None]
    rx, ry = ux * cos[:, None] - uy * sin[:, None], ux * sin[:, None] + uy * cos[:, None]
    w, h = rx.max(1)[0] - rx.min(1)[0], ry.max(1)[0] - ry.min(1)[0]
    swap = w > h                                                # regular_obb: the longer side first, theta into [-pi/2, pi/2)
    wr, hr = torch.where(swap, w, h), torch.where(swap, h, w)
    th = torch.where(swap, theta, theta + _PI / 2)
    th = torch.remainder(th + _PI / 2, _PI) - _PI / 2
    return torch.stack([xm, ym, gz, wr, hr, gd, th], 1)


def decode_fcos_obb(locations: torch.Tensor, reg: torch.Tensor) -> torch.Tensor:
    """FCOS midpoint-offset decode (fcos/utils.py:12-61), differentiable: locations (P,3), reg (P,8) = distances to the six faces of the OBB's
    AABB + (alpha, beta) -> (P,7) (x,y,z,w,l,h,theta).  Inference runs the same formulas inside nrpn_fcos_proposals (csrc/fcos_post.cu); here they
    are torch ops because the rotated-IoU loss of the FCOS head back-propagates through them on the positive locations of a training step."""
    lo = locations - reg[:, 0:3]
    hi = locations + reg[:, 3:6]
    ext = hi[:, :2] - lo[:, :2]
    c = (lo + hi) / 2
    v = (hi[:, :2] + lo[:, :2]) / 2 + reg[:, 6:8] * ext                     # the vertex on the top edge (x) / on the right edge (y)
    v = torch.maximum(torch.minimum(v, hi[:, :2]), lo[:, :2])               # clamp(min=lo, max=hi)
    v0 = torch.stack([v[:, 0], hi[:, 1]], 1) - c[:, :2]
    v1 = torch.stack([hi[:, 0], v[:, 1]], 1) - c[:, :2]
    d0, d1 = torch.norm(v0, dim=1), torch.norm(v1, dim=1)
    dmax = torch.max(d0, d1)
    c2 = c[:, :2]
    v0 = v0 / (d0[:, None] + 1e-7) * dmax[:, None] + c2                     # rectangularise: both half diagonals as long as the longer one
    v1 = v1 / (d1[:, None] + 1e-7) * dmax[:, None] + c2                     # (centre added back first, as the reference does: a degenerate
    length = torch.norm(v0 - v1, dim=1)                                     #  box -- opposite corners -- then has mid == 0 exactly)
    mid = (v0 + v1) / 2 - c2
    width = torch.norm(mid, dim=1) * 2
    degenerate = (mid[:, 0] == 0) & (mid[:, 1] == 0)
    mx = torch.where(degenerate, torch.full_like(mid[:, 0], 1e-7), mid[:, 0])
    theta = torch.atan2(mid[:, 1], mx)
    return torch.stack([c[:, 0], c[:, 1], c[:, 2], width, length, hi[:, 2] - lo[:, 2], theta], 1)
