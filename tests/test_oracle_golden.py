"""CPU: the oracle (oracle/) against the golden vectors generated from the unmodified reference
(tools/make_golden.py).  Tolerances: IoU 2e-6 absolute (the reference's own CPU/GPU builds differ by more:
ATen reduction trees, libm), NMS keep sets exact, decoded boxes 1e-5 relative."""
import os

import numpy as np
import pytest

from oracle import box as obox
from oracle import rpn_post as rp

SIZES = ((8,), (16,), (32,), (64,))
ASPECT = ((1., 1., 1.), (1., 1., 2.), (1., 2., 2.), (1., 1., 3.), (1., 3., 3.))


@pytest.fixture(scope="module")
def G(golden_dir):
    return {n: np.load(os.path.join(golden_dir, n + ".npz")) for n in
            ("iou", "nms", "decode", "sort_vertices", "rpn_small_aabb", "rpn_small_obb")}


def test_iou_known_answers(G):
    g = G["iou"]
    got = obox.iou_pairs(g["kat_a"], g["kat_b"])
    np.testing.assert_array_equal(got, g["kat_iou"])          # incl. the IoU(C,C)=1/3 quirk vector
    assert abs(got[0] - 0.11379701644182205) < 1e-7 and got[1] == 1.0 and abs(got[5] - 1 / 3) < 1e-6


def test_iou_random_pairs_and_matrices(G):
    g = G["iou"]
    np.testing.assert_allclose(obox.iou_pairs(g["rand_a"], g["rand_b"]), g["rand_iou"], rtol=0, atol=2e-6)
    m = obox.iou_matrix(g["mat_boxes"], g["mat_boxes"])
    off = ~np.eye(m.shape[0], dtype=bool)
    np.testing.assert_allclose(m[off], g["mat_iou"][off], rtol=0, atol=2e-6)
    np.testing.assert_allclose(np.diag(m), np.diag(g["mat_iou"]), rtol=0, atol=2e-5)   # self-IoU sits on every tolerance edge
    np.testing.assert_array_equal(obox.iou_matrix(g["aabb_boxes"], g["aabb_boxes"]), g["aabb_iou"])   # AABB is exact


def test_sort_vertices_restatement(G):
    g = G["sort_vertices"]
    idx = obox.sort_vertices(g["vertices"], g["mask"], g["num_valid"])
    np.testing.assert_array_equal(idx, g["idx"])


@pytest.mark.parametrize("name,thr", [("s64", 0.3), ("o700", 0.3), ("a1500", 0.3)])
def test_nms_keep_sets(G, name, thr):
    g = G["nms"]
    keep = obox.nms(g[f"{name}_boxes"], g[f"{name}_scores"], thr)
    np.testing.assert_array_equal(keep, g[f"{name}_keep"])
    bkeep = obox.batched_nms(g[f"{name}_boxes"], g[f"{name}_scores"], g[f"{name}_levels"], thr)
    np.testing.assert_array_equal(bkeep, g[f"{name}_bkeep"])


def test_nms_other_threshold_and_edge_cases(G):
    g = G["nms"]
    np.testing.assert_array_equal(obox.nms(g["o700_boxes"], g["o700_scores"], 0.5), g["o700_keep_t5"])
    assert obox.nms(np.zeros((0, 7), np.float32), np.zeros((0,), np.float32), 0.3).shape == (0,)
    one = np.array([[1, 1, 1, 2, 2, 2, 0.1]], np.float32)
    np.testing.assert_array_equal(obox.nms(one, np.array([0.5], np.float32), 0.3), [0])
    first10 = [4, 23, 51, 50, 56, 46, 7, 2, 37, 53]
    assert list(obox.nms(g["s64_boxes"], g["s64_scores"], 0.3)[:10]) == first10


def test_anchors_and_decode(G):
    g = G["decode"]
    cells = [rp.cell_anchors(s, ASPECT) for s in SIZES]
    for i in range(4):
        np.testing.assert_array_equal(cells[i], g[f"cell{i}"])
    grids = [(8, 12, 10), (4, 6, 5), (2, 3, 3), (1, 2, 2)]
    mesh = (32, 48, 40)
    strides = [tuple(mesh[i] // gr[i] for i in range(3)) for gr in grids]
    anc = np.concatenate([rp.grid_anchors(cells[l], grids[l], strides[l]) for l in range(4)])
    np.testing.assert_array_equal(anc, g["anchors"])
    np.testing.assert_allclose(rp.decode_aabb(g["d6"], g["anchors"]), g["dec6"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(rp.decode_obb(g["d8"], g["anchors"]), g["dec7"], rtol=1e-5, atol=1e-3)   # h of near-degenerate rectangles is a cancellation


@pytest.mark.parametrize("name,rot", [("rpn_small_aabb", False), ("rpn_small_obb", True)])
def test_rpn_postprocessing_end_to_end(G, name, rot):
    r = G[name]
    grids = [(8, 12, 10), (4, 6, 5), (2, 3, 3), (1, 2, 2)]
    mesh = (32, 48, 40)
    strides = [tuple(mesh[i] // gr[i] for i in range(3)) for gr in grids]
    cells = [rp.cell_anchors(s, ASPECT) for s in SIZES]
    code = 8 if rot else 6
    logits, deltas = [], []
    for l in range(4):
        lg, dl = r[f"logits{l}"], r[f"deltas{l}"]
        logits.append(np.transpose(lg, (1, 2, 3, 0)).reshape(-1))
        deltas.append(np.transpose(dl.reshape(13, code, *dl.shape[1:]), (2, 3, 4, 0, 1)).reshape(-1, code))
    b, s, lv = rp.rpn_proposals(logits, deltas, grids, strides, cells, mesh, rot)
    assert b.shape == r["proposals"].shape                    # identical proposal count ...
    np.testing.assert_array_equal(lv, r["level_index"])       # ... same ordering (levels line up row by row)
    np.testing.assert_allclose(b, r["proposals"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(s, r["scores"], rtol=0, atol=2e-7)
