"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the greedy matching inside the reference's recall metric, eval.py:33-52
(`evaluate_box_proposals_recall`): repeatedly take, per ground-truth box, the maximum IoU over proposals; pick the
best-covered ground truth; record that IoU; invalidate the proposal's row and the ground truth's column with -1.
numpy argmax returns the first maximal index, like torch.max on CPU.  Pinned by tests/golden/recall_small_obb.npz
(`match_m*` / `match_o*`: outputs of the reference function on tie-heavy random matrices)."""
import numpy as np


def greedy_match(overlaps: np.ndarray) -> np.ndarray:
    ov = np.array(overlaps, dtype=np.float32, copy=True)
    p, g = ov.shape
    out = np.zeros(g, dtype=np.float32)
    for j in range(min(p, g)):
        max_overlaps = ov.max(axis=0)
        argmax_overlaps = ov.argmax(axis=0)
        gt_ind = int(max_overlaps.argmax())
        box_ind = int(argmax_overlaps[gt_ind])
        out[j] = ov[box_ind, gt_ind]
        ov[box_ind, :] = -1
        ov[:, gt_ind] = -1
    return out


def recall(gt_overlaps_per_scene, num_pos, thresholds):
    """eval.py:57-72: sorted concatenation, recall_t = #(gt_overlaps >= t) / num_pos."""
    allv = np.sort(np.concatenate(gt_overlaps_per_scene)) if gt_overlaps_per_scene else np.zeros(0, np.float32)
    return np.array([(allv >= np.float32(t)).sum() / float(num_pos) for t in thresholds], dtype=np.float32), allv
