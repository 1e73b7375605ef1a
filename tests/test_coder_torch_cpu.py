"""CPU: the differentiable torch decode used by the IoU-type regression losses (nerf_rpn_b200/model/coder_torch.py) against the device decode of
csrc/rpn_decode.cuh compiled for the host (tests/host_shim) -- which the goldens pin to the reference's MidpointOffsetCoder."""
import ctypes
import os
import subprocess

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_torch_decode_matches_device_decode(tmp_path):
    from nerf_rpn_b200.model.coder_torch import decode_obb
    out = str(tmp_path / "libbox_shim.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "host_shim", "box_iou_host.cpp")])
    S = ctypes.CDLL(out)
    fp = ctypes.POINTER(ctypes.c_float)
    S.shim_decode.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, fp]
    rng = np.random.default_rng(0)
    n = 20000
    c, half = rng.random((n, 3)) * 100, rng.random((n, 3)) * 20 + 2
    an = np.concatenate([c - half, c + half], 1).astype(np.float32)
    d = (rng.standard_normal((n, 8)) * np.array([0.3, 0.3, 0.3, 0.5, 0.5, 0.5, 0.4, 0.4])).astype(np.float32)
    want = np.empty((n, 7), np.float32)
    S.shim_decode(an.ctypes.data_as(fp), d.ctypes.data_as(fp), n, 1, want.ctypes.data_as(fp))
    dt = torch.from_numpy(d).requires_grad_(True)
    got = decode_obb(torch.from_numpy(an), dt)
    err = (got.detach().numpy() - want)
    err[:, 6] = np.minimum(np.abs(err[:, 6]), np.abs(np.abs(err[:, 6]) - 3.141592))          # theta modulo pi at the wrap
    assert np.abs(err).max() < 1e-3, np.abs(err).max(0)
    got.sum().backward()
    assert torch.isfinite(dt.grad).all() and dt.grad.abs().sum() > 0


def test_split_anchor_index_matches_the_flat_anchor_order():
    """index = level offset + voxel * A + a (rpn.py:20-27): the helper the IoU-type loss uses to find a sampled anchor's deltas."""
    from nerf_rpn_b200 import train as T
    plan = T._TrainPlan.__new__(T._TrainPlan)
    plan.feat_dims = [(4, 6, 5), (2, 3, 3), (1, 2, 2)]
    plan.eng = type("E", (), {"A": 13})()
    A = 13
    want = []
    for l, d in enumerate(plan.feat_dims):
        for v in range(d[0] * d[1] * d[2]):
            for a in range(A):
                want.append((l, v, a))
    idx = torch.arange(len(want))
    level, vox, a = plan._split_anchor_index(idx)
    got = list(zip(level.tolist(), vox.tolist(), a.tolist()))
    assert got == want
