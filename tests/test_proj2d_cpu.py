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


@pytest.mark.parametrize("rotated", [True, False])
def test_engine_plumbing_of_the_projection_loss_on_cpu(rotated):
    """_TrainPlan._proj2d_loss (nerf_rpn_b200/train.py) on hand-made predictor tensors: the deltas of the sampled positives are gathered from the
    (voxel, 128-channel) rows of the right level, and the weighted gradient lands on exactly those entries of d(pred), added to what is there."""
    from nerf_rpn_b200 import train as T
    from nerf_rpn_b200.model.coder_torch import decode_aabb, decode_obb
    from nerf_rpn_b200.model.proj2d import rpn_projection_loss
    A, code = 13, 8 if rotated else 6
    g = torch.Generator().manual_seed(5 + int(rotated))
    plan = T._TrainPlan.__new__(T._TrainPlan)
    plan.feat_dims = [(4, 6, 5), (2, 3, 3), (1, 2, 2)]
    plan.dims = (16, 24, 20)
    plan.eng = type("E", (), dict(A=A, code=code, rotated=rotated, device=torch.device("cpu"), loss_scale=4.0,
                                  loss_2d=torch.zeros(())))()
    n_anchor = sum(d[0] * d[1] * d[2] for d in plan.feat_dims) * A
    c, half = torch.rand(n_anchor, 3, generator=g) * 20, torch.rand(n_anchor, 3, generator=g) * 4 + 1
    plan.anchors = torch.cat([c - half, c + half], 1)
    plan.pred_levels = [torch.randn(2, *d, 128, generator=g) * 0.3 for d in plan.feat_dims]
    plan.dpred_levels = [torch.full((2, *d, 128), 0.25, dtype=torch.bfloat16) for d in plan.feat_dims]
    samples, want_total = [], 0.0
    n_pos = 40 + 25
    grads = []
    for i, k in enumerate((40, 25)):
        pos = torch.randperm(n_anchor, generator=g)[:k]
        if rotated:
            gtp = torch.cat([torch.rand(k, 3, generator=g) * 20, torch.rand(k, 3, generator=g) * 6 + 2, (torch.rand(k, 1, generator=g) - 0.5) * 3], 1)
        else:
            lo = torch.rand(k, 3, generator=g) * 16
            gtp = torch.cat([lo, lo + torch.rand(k, 3, generator=g) * 6 + 1], 1)
        samples.append((pos, torch.zeros(0, dtype=torch.long), gtp))
        # the same thing on the flattened (anchor, code) view of the predictor output
        flat = torch.cat([p[i].reshape(-1, 128)[:, A:A + A * code].reshape(-1, code) for p in plan.pred_levels])
        d = flat[pos].clone().requires_grad_(True)
        boxes = (decode_obb if rotated else decode_aabb)(plan.anchors[pos], d)
        loss = rpn_projection_loss(boxes, gtp, n_pos, max(plan.dims))
        (gd,) = torch.autograd.grad(loss, d)
        want_total += loss.item()
        grads.append((pos, gd))
    plan.last_samples = samples
    plan._proj2d_loss(0.5)
    np.testing.assert_allclose(plan.eng.loss_2d.item(), want_total, rtol=1e-5)
    for i, (pos, gd) in enumerate(grads):
        got = torch.cat([p[i].reshape(-1, 128)[:, A:A + A * code].reshape(-1, code) for p in plan.dpred_levels]).float()
        want = torch.full_like(got, 0.25)
        want[pos] += gd * (0.5 * 4.0)
        assert torch.allclose(got, want.bfloat16().float(), rtol=2e-2, atol=2e-3)
        touched = torch.zeros(got.shape[0], dtype=torch.bool); touched[pos] = True
        assert (got[~touched] == 0.25).all()
        obj = torch.cat([p[i].reshape(-1, 128)[:, :A].reshape(-1) for p in plan.dpred_levels]).float()
        assert (obj == 0.25).all()                                  # the objectness channels are not the projection loss' business
    plan.eng.loss_2d.fill_(7.0)
    plan._proj2d_loss(0.0)                                          # value only (the drop-in path's report): d(pred) untouched
    np.testing.assert_allclose(plan.eng.loss_2d.item(), want_total, rtol=1e-5)
