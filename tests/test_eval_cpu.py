"""CPU: the oracle restatement of the recall matching loop against the reference's own outputs (golden)."""
import os

import numpy as np

from oracle import eval_oracle


def test_greedy_match_oracle_matches_reference_outputs(golden_dir):
    g = np.load(os.path.join(golden_dir, "recall_small_obb.npz"))
    k = 0
    while f"match_m{k}" in g.files:
        m, o = g[f"match_m{k}"], g[f"match_o{k}"]
        got = np.sort(eval_oracle.greedy_match(m))            # the reference returns the sorted concatenation
        np.testing.assert_array_equal(got, o)
        k += 1
    assert k == 4


def test_recall_definition(golden_dir):
    g = np.load(os.path.join(golden_dir, "recall_small_obb.npz"))
    for limit in (300, 1000, 2500):
        r, _ = eval_oracle.recall([g[f"gt_overlaps_{limit}"]], 12 * 24, [0.25, 0.5])
        np.testing.assert_allclose(r, g[f"recalls_{limit}"], rtol=0, atol=1e-7)
