"""Training-time scene augmentation on the device (the dataset side of the training step).

Mirrors BaseDataset.augment_rpn_inputs (datasets.py:109-163) and rotate_and_scale_scene (datasets.py:290-329) for z-up scenes:
  rot90 about z (transpose x/y + flip x)  ->  flips along x and y  ->  small rotation about z + isotropic scale (OBB targets only).
The reference does this per sample on a DataLoader worker with torch CPU ops (transpose / flip copies of the 168 MB grid and an
F.grid_sample over 10.5 M points); here the three steps are ONE kernel over the grid already resident in HBM (nrpn_augment_scene:
one 128-bit load per source voxel, eight for the trilinear case, one 128-bit store per output voxel).  The box transforms are a few
scalars per box and are the same torch expressions the reference evaluates.

The random decisions are drawn from Python's `random` in the reference's order (rotate, flip x, flip y, [rot-scale, angle, scale]), so
a run seeded like the reference's makes the same decisions.
"""
import ctypes
import math
import random
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from ._lib import check, lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


@dataclass
class Augmentation:
    rot90: bool = False
    flip_x: bool = False
    flip_y: bool = False
    angle: Optional[float] = None            # rotate_and_scale_scene's (angle, scale); None = step skipped
    scale: Optional[float] = None

    @property
    def identity(self) -> bool:
        return not (self.rot90 or self.flip_x or self.flip_y or self.angle is not None)


def draw_augmentation(flip_prob: float, rotate_prob: float, rot_scale_prob: float, boxes_are_obb: bool, rng=random) -> Augmentation:
    """The reference's sequence of random draws (datasets.py:123,146,158-160)."""
    for name, p in (("flip_prob", flip_prob), ("rotate_prob", rotate_prob), ("rotate_and_scale_prob", rot_scale_prob)):
        if p < 0 or p > 1:
            raise ValueError(f"{name} must be between 0 and 1, but got {p}")
    a = Augmentation()
    a.rot90 = rng.random() < rotate_prob
    a.flip_x = rng.random() < flip_prob
    a.flip_y = rng.random() < flip_prob
    if boxes_are_obb and rng.random() < rot_scale_prob:
        a.angle = rng.uniform(-math.pi / 18, math.pi / 18)
        a.scale = rng.uniform(0.9, 1.1)
    return a


def augment_boxes(boxes: Optional[torch.Tensor], aug: Augmentation, in_dims: Tuple[int, int, int]) -> Optional[torch.Tensor]:
    """Ground-truth boxes of the augmented scene: (n,6) [xmin,ymin,zmin,xmax,ymax,zmax] or (n,7) [x,y,z,w,l,h,theta]."""
    if boxes is None:
        return None
    obb = boxes.shape[1] == 7
    assert obb or boxes.shape[1] == 6
    b = boxes.clone()
    ext = [int(in_dims[0]), int(in_dims[1]), int(in_dims[2])]
    if aug.rot90:                                                   # datasets.py:131-142
        ext[0], ext[1] = ext[1], ext[0]
        b[:, [0, 1, 3, 4]] = b[:, [1, 0, 4, 3]]
        if obb:
            b[:, 0] = ext[0] - b[:, 0]
        else:
            b[:, [0, 3]] = ext[0] - b[:, [3, 0]]
    for axis, on in ((0, aug.flip_x), (1, aug.flip_y)):              # datasets.py:144-156
        if not on:
            continue
        if obb:
            b[:, axis] = ext[axis] - b[:, axis]
            b[:, -1] = -b[:, -1]
        else:
            b[:, [axis, axis + 3]] = ext[axis] - b[:, [axis + 3, axis]]
    if aug.angle is not None:                                        # datasets.py:318-327
        assert obb
        c, s = math.cos(aug.angle), math.sin(aug.angle)
        xform = torch.tensor([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=torch.float) * aug.scale
        b[:, 6] = b[:, 6] - aug.angle
        b[:, 3:6] = b[:, 3:6] / aug.scale
        centre = torch.tensor(ext).unsqueeze(0) / 2
        off = (b[:, :3] - centre.to(b.device)) @ (xform.to(b.dtype) / (aug.scale * aug.scale)).to(b.device)
        b[:, :3] = off + centre.to(b.device)
    return b


def augment_scene(rgbsigma: torch.Tensor, aug: Augmentation, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rgbsigma: CUDA fp32 (4, X, Y, Z) -- the dataset's view of an (X, Y, Z, 4) array (datasets.py:55-56) or a contiguous NCDHW tensor
    (re-laid out once).  Returns the augmented scene as a (4, Xo, Yo, Z) view of a new (Xo, Yo, Z, 4) array."""
    if not (isinstance(rgbsigma, torch.Tensor) and rgbsigma.is_cuda and rgbsigma.dtype == torch.float32 and rgbsigma.dim() == 4 and rgbsigma.shape[0] == 4):
        raise RuntimeError("nerf_rpn_b200.augment: needs a CUDA fp32 (4, X, Y, Z) grid (no CPU path)")
    src = rgbsigma.permute(1, 2, 3, 0)
    if not src.is_contiguous():
        src = src.contiguous()
    if aug.identity:
        return src.permute(3, 0, 1, 2)
    X, Y, Z = src.shape[:3]
    xo, yo = (Y, X) if aug.rot90 else (X, Y)
    if out is None:
        out = torch.empty((xo, yo, Z, 4), dtype=torch.float32, device=src.device)
    elif tuple(out.shape) != (xo, yo, Z, 4) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != src.device:
        raise ValueError(f"out must be a contiguous fp32 ({xo}, {yo}, {Z}, 4) tensor on {src.device}")
    resample = aug.angle is not None
    check(lib().nrpn_augment_scene(_p(src), X, Y, Z, _p(out), int(aug.rot90), int(aug.flip_x), int(aug.flip_y), int(resample), float(aug.angle or 0.0),
                                   float(aug.scale or 1.0), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "nrpn_augment_scene")
    return out.permute(3, 0, 1, 2)


def augment_rpn_inputs(rgbsigma: torch.Tensor, boxes: Optional[torch.Tensor], flip_prob: float, rotate_prob: float, rot_scale_prob: float,
                       z_up: bool = True, rng=random):
    """Drop-in for BaseDataset.augment_rpn_inputs (datasets.py:109-163) with the scene on the device."""
    if not z_up:
        raise NotImplementedError("nerf_rpn_b200.augment: only z-up scenes (every dataset class of the reference passes z_up=True)")
    if boxes is not None:
        assert boxes.shape[1] in (6, 7)
    aug = draw_augmentation(flip_prob, rotate_prob, rot_scale_prob, boxes is not None and boxes.shape[1] == 7, rng)
    dims = tuple(rgbsigma.shape[1:])
    return augment_scene(rgbsigma, aug), augment_boxes(boxes, aug, dims)
