"""CPU: the device functions of nerf_rpn_b200/csrc/box_iou.cuh, compiled for the host (tests/host_shim), against oracle/box_oracle.c
in every arithmetic mode (NRPN_IOU_MODE / orc_set_mode): the fused kernel's control flow -- polygon clipping, masked mean, angular
selection sort, shoelace terms and the three summation orders -- must agree with the oracle bit for bit.  Mode 3 on a GPU is what
tests/test_gpu_reference.py holds against the unmodified reference running on that GPU."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import box as obox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "libbox_shim.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "host_shim", "box_iou_host.cpp")])
    S = ctypes.CDLL(out)
    fp = ctypes.POINTER(ctypes.c_float)
    S.shim_iou_pairs.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
    return S


def _boxes(rng, n, ext):
    return np.concatenate([rng.random((n, 3)) * ext, rng.random((n, 3)) * 10 + 1, (rng.random((n, 1)) - 0.5) * math.pi], 1).astype(np.float32)


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_fused_iou_logic_matches_oracle_in_every_mode(shim, mode, golden_dir):
    fp = ctypes.POINTER(ctypes.c_float)
    rng = np.random.default_rng(mode)
    n = 60000
    a, b = _boxes(rng, n, 12.0), _boxes(rng, n, 12.0)
    a[:1500] = b[:1500]                                            # identical boxes: the nv == 8 special case of K1
    g = np.load(os.path.join(golden_dir, "iou.npz"))
    a = np.concatenate([a, g["kat_a"]]).astype(np.float32); b = np.concatenate([b, g["kat_b"]]).astype(np.float32)
    O = obox.lib()
    O.orc_set_mode.argtypes = [ctypes.c_int]
    try:
        shim.shim_set_iou_mode(mode); O.orc_set_mode(mode)
        got = np.empty(a.shape[0], np.float32)
        for cull in (0, 1):
            shim.shim_iou_pairs(a.ctypes.data_as(fp), b.ctypes.data_as(fp), a.shape[0], 7, cull, got.ctypes.data_as(fp))
            want = obox.iou_pairs(a, b)
            assert (want > 0).sum() > 15000
            np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    finally:
        O.orc_set_mode(0); shim.shim_set_iou_mode(0)
