"""Proposal recall on the device: the mirror of `evaluate_box_proposals_recall` (eval.py:14-81, SURVEY.md 8(a) a19 /
8(f) rank 1).

Same arguments and return dictionary as the reference. Per scene the proposals are ordered by score (descending), truncated
to `limit`, the (P, G) IoU matrix comes from the fused OBB/AABB kernel (`nrpn_iou3d_matrix`) and the greedy
"best-covered ground truth first" matching runs in one CTA (`nrpn_recall_match`); the only device->host traffic is the
min(P, G) matched IoUs per scene.  The reference moves the whole matrix to the CPU and loops in Python.
"""
from typing import List, Optional

import torch

from . import ops


def _cuda(t: torch.Tensor) -> torch.Tensor:
    return t.to(device="cuda", dtype=torch.float32).contiguous()


@torch.no_grad()
def evaluate_box_proposals_recall(proposals_list: List[torch.Tensor], proposal_scores_list: List[torch.Tensor],
                                  gt_boxes_list: List[torch.Tensor], thresholds: Optional[torch.Tensor] = None,
                                  limit: Optional[int] = None):
    gt_overlaps = []
    num_pos = 0
    for proposals, scores, gt_boxes in zip(proposals_list, proposal_scores_list, gt_boxes_list):
        if proposals.shape[0] == 0 or gt_boxes.shape[0] == 0:
            continue
        ids = torch.argsort(scores, descending=True)
        proposals = proposals[ids]
        num_pos += gt_boxes.shape[0]
        if limit is not None and len(proposals) > limit:
            proposals = proposals[:limit]
        overlaps = ops.iou3d_matrix(_cuda(proposals), _cuda(gt_boxes))
        matched = ops.recall_match(overlaps)
        _gt = torch.zeros(gt_boxes.shape[0])
        _gt[: matched.numel()] = matched.cpu()
        gt_overlaps.append(_gt)
    gt_overlaps = torch.cat(gt_overlaps, dim=0) if len(gt_overlaps) else torch.zeros(0, dtype=torch.float32)
    gt_overlaps, _ = torch.sort(gt_overlaps)
    if thresholds is None:
        thresholds = torch.arange(0.5, 0.95 + 1e-5, 0.05, dtype=torch.float32)
    recalls = torch.zeros_like(thresholds)
    for i, t in enumerate(thresholds):
        recalls[i] = (gt_overlaps >= t).float().sum() / float(num_pos)
    return {"ar": recalls.mean(), "recalls": recalls, "thresholds": thresholds, "gt_overlaps": gt_overlaps, "num_pos": num_pos}
