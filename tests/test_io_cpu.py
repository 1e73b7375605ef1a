"""CPU: nerf_rpn_b200.io.read_rgbsigma (scene file -> host buffer in file order, no transpose / conversion) on .npy and
.npz (stored and deflated) files, float32 and uint8, into a caller-provided buffer (pinned allocation needs a CUDA runtime: tests/test_gpu_conv.py)."""
import numpy as np
import pytest
import torch

from nerf_rpn_b200 import io


@pytest.mark.parametrize("dtype", [np.float32, np.uint8])
@pytest.mark.parametrize("kind", ["npy", "npz", "npz_compressed"])
def test_read_rgbsigma_into_buffer(tmp_path, dtype, kind):
    rng = np.random.default_rng(0)
    a = (rng.random((9, 7, 5, 4)) * 255).astype(dtype)
    if kind == "npy":
        p = str(tmp_path / "s.npy"); np.save(p, a)
    elif kind == "npz":
        p = str(tmp_path / "s.npz"); np.savez(p, rgbsigma=a, resolution=np.array([9, 7, 5]))
    else:
        p = str(tmp_path / "s.npz"); np.savez_compressed(p, rgbsigma=a)
    out = torch.empty(a.shape, dtype=torch.float32 if dtype == np.float32 else torch.uint8)
    got = io.read_rgbsigma(p, out=out)
    assert got.shape == (4, 9, 7, 5) and got.data_ptr() == out.data_ptr()            # the (4, W, L, H) VIEW of datasets.py:55-56, no copy
    np.testing.assert_array_equal(got.permute(1, 2, 3, 0).numpy(), a)


def test_read_rgbsigma_rejects_what_the_stem_cannot_consume(tmp_path):
    p = str(tmp_path / "bad.npy")
    np.save(p, np.zeros((4, 4, 4, 3), np.float32))
    with pytest.raises(ValueError, match="expected a C-ordered"):
        io.read_rgbsigma(p, out=torch.empty(4, 4, 4, 3))
    np.save(p, np.zeros((4, 4, 4, 4), np.float64))
    with pytest.raises(ValueError, match="expected a C-ordered"):
        io.read_rgbsigma(p, out=torch.empty(4, 4, 4, 4))
    np.save(p, np.zeros((4, 4, 4, 4), np.float32))
    with pytest.raises(ValueError, match="out must be"):
        io.read_rgbsigma(p, out=torch.empty(4, 4, 4, 4, dtype=torch.uint8))
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-100])
    with pytest.raises(IOError, match="truncated"):
        io.read_rgbsigma(p, out=torch.empty(4, 4, 4, 4))
