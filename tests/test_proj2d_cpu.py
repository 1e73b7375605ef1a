"""CPU: the 2-D projection regression loss of the RPN (rpn.py:421-453, `--reg_loss_weight_2d`) as built in nerf_rpn_b200/model/proj2d.py, and the
differentiable torch decoders under it (model/coder_torch.py), against tests/golden/proj2d.npz = the reference's own RegionProposalNetwork.compute_loss
and coders (tools/make_golden.py gen_proj2d).  These are torch ops in the product too (evaluated on the <= 128 sampled positives of a training step), so the
CPU run exercises the code the GPU runs; tests/test_gpu_train.py checks the engine plumbing (gather of the deltas, gradient into d(pred))."""
import os

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "proj2d.npz"))


@pytest.mark.parametrize("kind", ["aabb", "obb"])
def test_projection_loss_value_and_gradient(g, kind):
    from nerf_rpn_b200.model.proj2d import rpn_projection_loss
    n_pos, res = int(g[f"{kind}_n_pos"]), int(g[f"{kind}_res"])
    pred = torch.tensor(g[f"{kind}_decoded"][:n_pos], requires_grad=True)          # the golden's positives are the first n_pos rows
    tgt = torch.tensor(g[f"{kind}_target"][:n_pos])
    loss = rpn_projection_loss(pred, tgt, n_pos, res)
    np.testing.assert_allclose(loss.item(), float(g[f"{kind}_loss_2d"]), rtol=2e-5)
    loss.backward()
    want = g[f"{kind}_dpred"]
    assert (want[n_pos:] == 0).all()
    np.testing.assert_allclose(pred.grad.numpy(), want[:n_pos], rtol=2e-3, atol=1e-6 * np.abs(want).max())
    # a mesh split (the engine loops over the meshes of a batch, dividing by the batch's positives) adds up to the same loss
    half = n_pos // 2
    parts = rpn_projection_loss(pred[:half], tgt[:half], n_pos, res) + rpn_projection_loss(pred[half:], tgt[half:], n_pos, res)
    np.testing.assert_allclose(parts.item(), loss.item(), rtol=1e-6)


@pytest.mark.parametrize("kind", ["aabb", "obb"])
def test_torch_decoders_match_reference_coders(g, kind):
    from nerf_rpn_b200.model.coder_torch import decode_aabb, decode_obb
    an, d = torch.tensor(g[f"{kind}_anchors"]), torch.tensor(g[f"{kind}_deltas"], requires_grad=True)
    got = (decode_aabb if kind == "aabb" else decode_obb)(an, d)
    want = g[f"{kind}_decoded"]
    err = np.abs(got.detach().numpy() - want)
    if kind == "obb":
        err[:, 6] = np.minimum(err[:, 6], np.abs(err[:, 6] - 3.141592))              # theta modulo pi at the wrap
    assert err.max() < 2e-4 * max(1.0, np.abs(want).max()), err.max(0)
    got.sum().backward()
    assert torch.isfinite(d.grad).all() and d.grad.abs().sum() > 0


def test_cameras_look_at_the_centroid():
    from nerf_rpn_b200.model.proj2d import IMG_H, IMG_W, project_all, w2c_matrices
    res = 160
    m = w2c_matrices(res)
    assert m.shape == (4, 4, 4) and m.dtype == np.float32
    px = project_all(torch.full((1, 3), res / 2.0), res)                                # the centroid projects onto the principal point of all four cameras
    np.testing.assert_allclose(px.numpy(), np.tile([[IMG_W / 2, IMG_H / 2]], (4, 1)), atol=1e-3)
