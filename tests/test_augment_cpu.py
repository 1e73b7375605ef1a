"""Host side of the training augmentation (nerf_rpn_b200/augment.py): the random draws and the box transforms against the reference's own
BaseDataset.augment_rpn_inputs (datasets.py:109-163) from the staged copy; the grid kernel is covered by tests/test_gpu_augment.py."""
import os
import random
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref():
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip("oracle/_ref not staged")
    stub = os.path.join(ROOT, "tools", "ref_stub")                 # import-time stand-in for the reference's native op on a CPU-only host
    sys.path.insert(0, stub)
    try:
        return ref_gpu.load(need_k1=False)
    finally:
        sys.path.remove(stub)


@pytest.mark.parametrize("obb", [True, False])
def test_draws_and_boxes_match_reference(obb):
    from nerf_rpn_b200 import augment
    ref = _ref()
    dims = (20, 26, 12)
    g = torch.Generator().manual_seed(5)
    for seed in range(40):
        grid = torch.rand(4, *dims, generator=g)
        ctr = torch.rand(9, 3, generator=g) * torch.tensor(dims, dtype=torch.float32)
        size = 2 + torch.rand(9, 3, generator=g) * 6
        boxes = torch.cat([ctr, size, (torch.rand(9, 1, generator=g) - 0.5) * 3], 1) if obb else torch.cat([ctr - size / 2, ctr + size / 2], 1)
        random.seed(seed)
        _, want = ref.datasets.BaseDataset.augment_rpn_inputs(grid, boxes, 0.5, 0.5, 0.6)
        state = random.getstate()
        random.seed(seed)
        aug = augment.draw_augmentation(0.5, 0.5, 0.6, obb)
        assert random.getstate() == state
        assert (aug.angle is not None) <= obb
        got = augment.augment_boxes(boxes, aug, dims)
        assert torch.allclose(got, want, rtol=0, atol=1e-5), (seed, aug)


def test_probability_validation_and_identity():
    from nerf_rpn_b200 import augment
    with pytest.raises(ValueError):
        augment.draw_augmentation(1.5, 0.0, 0.0, True)
    with pytest.raises(ValueError):
        augment.draw_augmentation(0.0, -0.1, 0.0, True)
    a = augment.draw_augmentation(0.0, 0.0, 0.0, True)
    assert a.identity
    b = torch.rand(3, 7)
    assert torch.equal(augment.augment_boxes(b, a, (8, 8, 8)), b)
    assert augment.augment_boxes(None, a, (8, 8, 8)) is None
