"""CPU: the loss-side oracle (sampler, encoders, smooth-L1 / BCE) against the reference's own outputs (tests/golden/loss_small.npz)."""
import os

import numpy as np

from oracle import loss_oracle as LO
from oracle import targets_oracle as TO
from tests.test_targets_cpu import small_anchors


def test_sampler_encoders_and_losses_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss_small.npz"))
    t = np.load(os.path.join(golden_dir, "targets_small.npz"))
    anchors = small_anchors().numpy()
    for kind in ("aabb", "obb"):
        gt = t[f"gt_a_{kind}_0"]
        labels = t[f"labels_a_{kind}_0"].astype(np.float32)
        matched = t[f"matched_a_{kind}_0"].astype(np.int64)
        pos_m, neg_m = LO.sample(labels, 256, 0.5, g[f"{kind}_perm_pos"], g[f"{kind}_perm_neg"])
        np.testing.assert_array_equal(pos_m, g[f"{kind}_pos_mask"])
        np.testing.assert_array_equal(neg_m, g[f"{kind}_neg_mask"])
        pos = np.nonzero(pos_m)[0]
        np.testing.assert_array_equal(pos, g[f"{kind}_pos_idx"])
        mgt = gt[np.clip(matched, 0, None)][pos]
        tg = LO.encode_aabb(mgt, anchors[pos]) if kind == "aabb" else LO.encode_obb_midpoint(anchors[pos], mgt)
        ref = g[f"{kind}_targets_pos"]
        assert np.abs(tg - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), kind       # log / cos / sin: libm vs numpy, last-ulp
        neg = g[f"{kind}_neg_idx"]
        sampled_labels = np.concatenate([labels[pos], labels[neg]])
        box = LO.smooth_l1_sum(g[f"{kind}_deltas_pos"], ref, 1.0 / 9.0) / float(pos.shape[0] + neg.shape[0])
        obj = LO.bce_with_logits_mean(g[f"{kind}_objectness_sampled"], sampled_labels)
        assert abs(box - float(g[f"{kind}_loss_box"])) <= 1e-5 * max(1.0, abs(float(g[f"{kind}_loss_box"])))
        assert abs(obj - float(g[f"{kind}_loss_obj"])) <= 1e-6
