"""Box utilities. Mirrors the hot-path functions of nerf_rpn/model/utils.py (215-265, 268-289, 344-458) with
the same signatures and return conventions, on top of the fused CUDA kernels.  Inputs may live on CPU or GPU
(the reference's OBB branch moves them to the GPU itself, utils.py:412); results are returned on CPU where the
reference returns CPU tensors (nms -> CPU LongTensor, utils.py:230; OBB box_iou_3d -> CPU fp32, utils.py:412)."""
from typing import Tuple

import torch
from torch import Tensor

from .. import ops


def _cuda(t: Tensor) -> Tensor:
    if not torch.cuda.is_available():
        raise RuntimeError("nerf_rpn_b200 requires a CUDA device (no CPU fallback)")
    return t.detach().to(device="cuda", dtype=torch.float32).contiguous()


def box_iou_3d(boxes1: Tensor, boxes2: Tensor) -> Tensor:
    if boxes1.size(1) == boxes2.size(1) == 6:
        return ops.iou3d_matrix(_cuda(boxes1), _cuda(boxes2)).to(boxes1.device)
    elif boxes1.size(1) == boxes2.size(1) == 7:
        return ops.iou3d_matrix(_cuda(boxes1), _cuda(boxes2)).cpu().type(torch.float32)
    raise ValueError("The second dimension of boxes1 and boxes2 should be the same, both 6 or 7. But get {} and {}."
                     .format(boxes1.size(1), boxes2.size(1)))


@torch.no_grad()
def batched_box_iou(boxes1: Tensor, boxes2: Tensor, batch_size=16) -> Tensor:
    return box_iou_3d(boxes1, boxes2)      # the fused kernel never materialises the (n,m,7) tiles: no batching needed


def nms(boxes: Tensor, scores: Tensor, iou_threshold: float) -> Tensor:
    keep, n = ops.nms_device(_cuda(boxes), _cuda(scores), None, iou_threshold)
    return keep[: int(n.item())].cpu()


def batched_nms(boxes: Tensor, scores: Tensor, idxs: Tensor, iou_threshold: float) -> Tensor:
    uniq, inv = torch.unique(idxs, return_inverse=True)
    if uniq.numel() > 255:
        raise ValueError("batched_nms supports at most 255 distinct categories")
    keep, n = ops.nms_device(_cuda(boxes), _cuda(scores), inv.to(device="cuda", dtype=torch.int32).contiguous(), iou_threshold)
    return keep[: int(n.item())].to(boxes.device)


def remove_small_boxes(boxes: Tensor, min_size: float) -> Tensor:
    if boxes.size(1) == 6:
        ws, hs, ds = boxes[:, 3] - boxes[:, 0], boxes[:, 4] - boxes[:, 1], boxes[:, 5] - boxes[:, 2]
    else:
        ws, hs, ds = boxes[:, 3], boxes[:, 4], boxes[:, 5]
    return torch.where((ws >= min_size) & (hs >= min_size) & (ds >= min_size))[0]


def clip_boxes_to_mesh(boxes: Tensor, size: Tuple[int, int, int]) -> Tensor:
    if boxes.size(1) == 6:
        out = boxes.clone()
        for k in range(3):
            out[..., k] = boxes[..., k].clamp(min=0, max=size[k])
            out[..., 3 + k] = boxes[..., 3 + k].clamp(min=0, max=size[k])
        return out
    valid = (boxes[..., 0] >= 0) & (boxes[..., 0] <= size[0]) & (boxes[..., 1] >= 0) & (boxes[..., 1] <= size[1]) & \
            (boxes[..., 2] >= 0) & (boxes[..., 2] <= size[2])
    return boxes[valid]


def print_shape(obj):
    def pt(o):
        return [pt(i) for i in o] if isinstance(o, (list, tuple)) else o.shape
    print(pt(obj))
