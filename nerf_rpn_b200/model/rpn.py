"""RegionProposalNetwork. Mirrors nerf_rpn/model/rpn.py:167-229,458-536 (constructor, attributes, eval forward).
The eval branch (head -> decode -> per-level top-k -> clip/filter -> per-level NMS -> top post_nms_top_n,
rpn.py:485-512) is executed by the fused engine owned by NeRFRegionProposalNetwork; this class carries the
hyper-parameters and the sub-modules under the reference's attribute names so checkpoints and DDP wrap alike.
Training (rpn.py:514-534): the anchor <-> ground-truth assignment (`assign_targets_to_anchors`, rpn.py:240-290) runs on the
device (`nrpn_assign_targets`); sampling, the losses, the backward pass and the optimiser step are not implemented in this round.
"""
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import ops


class RegionProposalNetwork(nn.Module):
    def __init__(self, anchor_generator, head, fg_iou_thresh: float, bg_iou_thresh: float, batch_size_per_mesh: int,
                 positive_fraction: float, pre_nms_top_n: Dict[str, int], post_nms_top_n: Dict[str, int],
                 nms_thresh: float, score_thresh: float = 0.0, iou_batch_size: int = 16, rotated_bbox: bool = False,
                 reg_loss_type: str = "smooth_l1"):
        super().__init__()
        self.anchor_generator = anchor_generator
        self.head = head
        self.rotate = rotated_bbox
        self.num_bbox_digits = 6 if not rotated_bbox else 7
        self.num_delta_digits = 6 if not rotated_bbox else 8
        self.fg_iou_thresh, self.bg_iou_thresh = fg_iou_thresh, bg_iou_thresh
        self.batch_size_per_mesh, self.positive_fraction = batch_size_per_mesh, positive_fraction
        self.iou_batch_size = iou_batch_size
        self._pre_nms_top_n = pre_nms_top_n
        self._post_nms_top_n = post_nms_top_n
        self.nms_thresh = nms_thresh
        self.score_thresh = score_thresh
        self.reg_loss_type = reg_loss_type
        self.min_size = 1e-3

    def pre_nms_top_n(self) -> int:
        return self._pre_nms_top_n["training" if self.training else "testing"]

    def post_nms_top_n(self) -> int:
        return self._post_nms_top_n["training" if self.training else "testing"]

    @torch.no_grad()
    def assign_targets_to_anchors(self, anchors: List[Tensor], targets: List[Tensor],
                                  padding_masks: Optional[List[Tensor]] = None) -> Tuple[List[Tensor], List[Tensor]]:
        """rpn.py:240-290 with the same arguments and results (labels in {1, 0, -1} and the matched ground-truth box of every
        anchor), computed by one fused device pass per mesh: no (G, num_anchors) IoU matrix, no chunking over ground truth."""
        labels, matched_gt_boxes = [], []
        for i, (anchors_per_mesh, gt_boxes) in enumerate(zip(anchors, targets)):
            dev = anchors_per_mesh.device
            a = anchors_per_mesh.to(device="cuda", dtype=torch.float32).contiguous()
            if gt_boxes.numel() == 0:                       # background mesh (rpn.py:246-250)
                matched_gt_boxes.append(torch.zeros(anchors_per_mesh.shape, dtype=torch.float32, device=dev))
                lab = torch.zeros((anchors_per_mesh.shape[0],), dtype=torch.float32, device=dev)
                if padding_masks is not None:
                    lab[~padding_masks[i].to(dev)] = -1.0
                labels.append(lab)
                continue
            gt = gt_boxes.to(device="cuda", dtype=torch.float32).contiguous()
            valid = None if padding_masks is None else padding_masks[i].to("cuda")
            lab, idx = ops.assign_targets(a, gt, valid, self.fg_iou_thresh, self.bg_iou_thresh, True)
            labels.append(lab.to(dev))
            matched_gt_boxes.append(gt[idx.clamp(min=0)].to(dev))
        return labels, matched_gt_boxes

    def forward(self, meshes, features, original_mesh_sizes, targets=None, objectness_output_paths=None):
        raise RuntimeError("nerf_rpn_b200.RegionProposalNetwork runs inside NeRFRegionProposalNetwork.forward "
                           "(backbone, head and post-processing are one captured launch sequence)")
