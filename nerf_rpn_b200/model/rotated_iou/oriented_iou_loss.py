"""cal_iou_3d. Mirrors nerf_rpn/model/rotated_iou/oriented_iou_loss.py:82-107 (forward values only: the fused
kernel has no autograd yet, so using it as a training loss -- rpn.py:133-165 -- is a later-round item)."""
import torch

from ... import ops


def cal_iou_3d(box3d1: torch.Tensor, box3d2: torch.Tensor, verbose=False):
    if verbose:
        raise NotImplementedError("nerf_rpn_b200: cal_iou_3d(verbose=True) (corners / union for the IoU loss) is not implemented")
    if box3d1.requires_grad or box3d2.requires_grad:
        raise NotImplementedError("nerf_rpn_b200: cal_iou_3d is forward-only in this round (no autograd)")
    shape = box3d1.shape[:-1]
    a = box3d1.detach().reshape(-1, 7).to(device="cuda", dtype=torch.float32).contiguous()
    b = box3d2.detach().reshape(-1, 7).to(device="cuda", dtype=torch.float32).contiguous()
    return ops.iou3d_pairs(a, b).reshape(shape)
