"""CPU: the target-assignment oracle against the reference's own outputs (tests/golden/targets_small.npz)."""
import os

import numpy as np
import torch

from oracle import targets_oracle as T
from tests import recipes


def small_anchors():
    from nerf_rpn_b200.model.anchor import AnchorGenerator3D
    dims = (32, 48, 40)
    ag = AnchorGenerator3D(recipes.ANCHOR_SIZES, recipes.ASPECT)
    feats = [torch.zeros(1, 1, *[(d + s - 1) // s for d in dims]) for s in (4, 8, 16, 32)]
    return ag(torch.zeros(1, 4, *dims), feats)[0][0]


def test_assignment_oracle_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "targets_small.npz"))
    anchors = small_anchors()
    assert anchors.shape[0] == int(g["n_anchors"]) and abs(anchors.double().sum().item() - float(g["anchors_sum"])) < 1e-6
    a = anchors.numpy()
    mask = g["mask"]
    for tag in ("a", "b"):
        for kind in ("obb", "aabb"):
            for use_mask in (0, 1):
                key = f"{tag}_{kind}_{use_mask}"
                labels, idx = T.assign(a, g["gt_" + key], mask if use_mask else None, 0.35, 0.2)
                bad = int((labels.astype(np.int8) != g["labels_" + key]).sum() + (idx != g["matched_" + key]).sum())
                # AABB ground truth: exact.  OBB: obb2hbb_3d goes through cos/sin (fp64-rounded here, libm fp32 in torch): allow a
                # handful of 1-ulp flips, none were observed
                assert bad == 0, f"{key}: {bad} mismatches"
