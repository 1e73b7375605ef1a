"""Input pipeline pieces of SURVEY.md 8(f) rank 2: scene files -> pinned host memory in their on-disk (W, L, H, 4) order, ready for a
plain-memcpy H2D copy and the channels-last stem packing (no transpose, no float conversion of uint8 grids, no density_to_alpha on the
host: the model applies it on the device when built with density_to_alpha_on_device=True).

The reference's datasets.py:39-63 does np.load -> (optional) density_to_alpha -> np.transpose -> torch.from_numpy -> (uint8) .float() / 255
on a DataLoader worker and .cuda() on the main process (run_rpn.py:473).  `read_rgbsigma` replaces the part before `.cuda()`.
"""
import zipfile

import numpy as np
import torch


def read_rgbsigma(path: str, out: torch.Tensor = None, key: str = "rgbsigma") -> torch.Tensor:
    """Reads array `key` of an .npz (stored or deflated) or a plain .npy straight INTO a pinned host tensor (allocated when `out` is None)
    and returns its (4, W, L, H) view -- the view datasets.py:55-56 produces, in the memory order of the file.  dtype = the file's
    (float32, or uint8 which the stem packing normalises on the device)."""
    def fill(fh):
        version = np.lib.format.read_magic(fh)
        shape, fortran, dtype = np.lib.format.read_array_header_1_0(fh) if version == (1, 0) else np.lib.format.read_array_header_2_0(fh)
        if fortran or len(shape) != 4 or shape[-1] != 4 or dtype not in (np.dtype("float32"), np.dtype("uint8")):
            raise ValueError(f"{path}: expected a C-ordered (W, L, H, 4) float32 / uint8 array, got shape {shape} dtype {dtype}")
        tdt = torch.float32 if dtype == np.dtype("float32") else torch.uint8
        buf = out
        if buf is None:
            buf = torch.empty(shape, dtype=tdt).pin_memory()
        if tuple(buf.shape) != tuple(shape) or buf.dtype != tdt or not buf.is_contiguous():
            raise ValueError(f"out must be a contiguous {tdt} tensor of shape {shape}")
        mv = memoryview(buf.numpy()).cast("B")
        got = 0
        while got < len(mv):
            k = fh.readinto(mv[got:])
            if not k:
                raise IOError(f"{path}: truncated array data")
            got += k
        return buf
    if path.endswith(".npy"):
        with open(path, "rb") as fh:
            buf = fill(fh)
    else:
        with zipfile.ZipFile(path) as zf, zf.open(key + ".npy") as fh:
            buf = fill(fh)
    return buf.permute(3, 0, 1, 2)
