"""Proposal recall and VOC-style AP on the device: mirrors of `evaluate_box_proposals_recall` (eval.py:14-81) and
`evaluate_box_proposals_ap` (eval.py:319-395) -- SURVEY.md 8(a) a19 / 8(f) rank 1.

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


@torch.no_grad()
def evaluate_box_proposals_ap(proposals_list: List[torch.Tensor], proposal_scores_list: List[torch.Tensor],
                              gt_boxes_list: List[torch.Tensor], iou_thresh: float = 0.25, top_k: Optional[int] = None):
    """eval.py:319-395. The reference calls box_iou_3d once per detection (one kernel chain + host sync each); here each scene's
    detections meet its ground truth in ONE IoU-matrix launch followed by a row-max launch, and only (max IoU, arg-max) per
    detection return to the host for the order-dependent true/false-positive bookkeeping."""
    num_gt = 0
    scene_ids, all_scores, all_max, all_arg = [], [], [], []
    for i, (proposals, scores, gt_boxes) in enumerate(zip(proposals_list, proposal_scores_list, gt_boxes_list)):
        if top_k is not None and len(proposals) > top_k:
            ids = torch.argsort(scores, descending=True)[:top_k]
            proposals, scores = proposals[ids], scores[ids]
        num_gt += gt_boxes.shape[0]
        n = len(proposals)
        scene_ids.append(torch.full((n,), i, dtype=torch.int64))
        all_scores.append(scores.detach().float().cpu())
        if n == 0:
            continue
        if gt_boxes.shape[0] == 0:
            all_max.append(torch.full((n,), -1.0)); all_arg.append(torch.zeros((n,), dtype=torch.int64))
            continue
        mv, am = ops.rowmax(ops.iou3d_matrix(_cuda(proposals), _cuda(gt_boxes)))
        all_max.append(mv.cpu()); all_arg.append(am.cpu().long())
    scene_ids = torch.cat(scene_ids) if scene_ids else torch.zeros(0, dtype=torch.int64)
    all_scores = torch.cat(all_scores) if all_scores else torch.zeros(0)
    all_max = torch.cat(all_max) if all_max else torch.zeros(0)
    all_arg = torch.cat(all_arg) if all_arg else torch.zeros(0, dtype=torch.int64)
    order = torch.argsort(all_scores, descending=True)
    scene_ids, all_max, all_arg = scene_ids[order].tolist(), all_max[order].tolist(), all_arg[order].tolist()
    gt_used = [[False] * len(g) for g in gt_boxes_list]
    n_det = len(scene_ids)
    tp = torch.zeros(n_det, dtype=torch.bool)
    fp = torch.zeros(n_det, dtype=torch.bool)
    for k in range(n_det):
        s, g = scene_ids[k], all_arg[k]
        if all_max[k] > iou_thresh and not gt_used[s][g]:
            tp[k] = True
            gt_used[s][g] = True
        else:
            fp[k] = True
    tp = torch.cumsum(tp, dim=0)
    fp = torch.cumsum(fp, dim=0)
    recalls = tp / num_gt
    precisions = tp / (tp + fp)
    mrec = torch.cat((torch.tensor([0.0]), recalls, torch.tensor([1.0])))
    mpre = torch.cat((torch.tensor([0.0]), precisions, torch.tensor([0.0])))
    for i in range(mpre.size(0) - 1, 0, -1):
        mpre[i - 1] = torch.max(mpre[i - 1], mpre[i])
    idx = torch.where(mrec[1:] != mrec[:-1])[0]
    ap = torch.sum((mrec[idx + 1] - mrec[idx]) * mpre[idx + 1])
    return {"ap": ap, "precisions": precisions, "recalls": recalls, "thresholds": iou_thresh, "num_det": tp + fp}
