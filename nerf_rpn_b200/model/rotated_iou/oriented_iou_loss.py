"""cal_iou_3d. Mirrors nerf_rpn/model/rotated_iou/oriented_iou_loss.py:82-107: same signature, `verbose=True` returns
(iou, corners1, corners2, z_range, u3d), and the result is differentiable w.r.t. both boxes (the reference uses it as a loss:
rpn.py:133-165 RotatedIOULoss, fcos/loss.py).  Forward values come from the fused exact kernel (bit-identical to the reference's chain
on the same GPU); the backward pass is nrpn_iou3d_pairs_backward (intersection volume differentiated in fp64).
cal_giou_3d / cal_diou_3d need the smallest enclosing rotated box (min_enclosing_box.py) and are not built."""
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


def cal_giou_3d(box3d1, box3d2, enclosing_type="smallest"):
    raise NotImplementedError("nerf_rpn_b200: cal_giou_3d (smallest enclosing rotated box, min_enclosing_box.py) is not implemented; "
                              "reg_loss_type 'iou' / 'linear_iou' use cal_iou_3d(verbose=True), which is")


def cal_diou_3d(box3d1, box3d2, enclosing_type="smallest"):
    raise NotImplementedError("nerf_rpn_b200: cal_diou_3d (smallest enclosing rotated box, min_enclosing_box.py) is not implemented; "
                              "reg_loss_type 'iou' / 'linear_iou' use cal_iou_3d(verbose=True), which is")
