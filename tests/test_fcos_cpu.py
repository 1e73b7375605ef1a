"""CPU: FCOS oracle (oracle/fcos_post.py) against the reference's golden outputs, and the FCOS module mirror's seeds/keys."""
import os

import numpy as np
import pytest
import torch

from oracle import fcos_post as fp
from tests import recipes

GRIDS = [(8, 12, 10), (4, 6, 5), (2, 3, 3), (1, 2, 2)]
STRIDES = [4, 8, 16, 32]


@pytest.mark.parametrize("name,obb,pre,post", [("fcos_small_aabb", False, 2500, 2500), ("fcos_small_aabb_tight", False, 300, 150),
                                               ("fcos_small_obb", True, 2500, 2500), ("fcos_small_obb_tight", True, 300, 150)])
def test_fcos_post_oracle_matches_reference(golden_dir, name, obb, pre, post):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    code = 8 if obb else 6
    cls = [g[f"logits{l}"].transpose(1, 2, 3, 0).reshape(-1) for l in range(4)]
    reg = [g[f"reg{l}"].transpose(1, 2, 3, 0).reshape(-1, code) for l in range(4)]
    ctr = [g[f"ctr{l}"].transpose(1, 2, 3, 0).reshape(-1) for l in range(4)]
    b, s = fp.fcos_proposals(cls, reg, ctr, None, GRIDS, STRIDES, (32, 48, 40), obb, 0.0, pre, 0.3, post, 0.0, reg_is_raw=False)
    assert b.shape == g["boxes"].shape                         # same count, incl. the k-th value cut of the tight setting
    np.testing.assert_array_equal(b[:, 0], g["boxes"][:, 0])   # same order (level ids line up)
    np.testing.assert_allclose(b, g["boxes"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(s, g["scores"], rtol=0, atol=2e-7)


def test_fcos_module_mirror_seeds_and_keys(golden_dir):
    from nerf_rpn_b200.model import feature_extractor
    from nerf_rpn_b200.model.fcos import fcos

    class NS:
        ResNet_FPN_256 = feature_extractor.ResNet_FPN_256
        Bottleneck = feature_extractor.Bottleneck
        FCOSOverNeRF = fcos.FCOSOverNeRF
    g = np.load(os.path.join(golden_dir, "fcos_small_obb.npz"))
    model = recipes.build_fcos_small(NS, True, g)
    keys = list(model.fcos_module.state_dict())
    assert "head.cls_tower.0.weight" in keys and "head.bbox_tower.10.bias" in keys and "head.scales.4.scale" in keys
    assert len(keys) == 2 * 4 * 4 + 6 + 5                      # 2 towers x 4 x (conv w,b + GN w,b) + 3 predictors + 5 scales
    with pytest.raises(NotImplementedError):
        model.train()(None)


def test_output_objectness_export(tmp_path):
    """FCOSModule.output_objectness (--output_voxel_scores of run_fcos.py, fcos.py:268-284): npz per scene, one array per level,
    sqrt(sigmoid(cls) * sigmoid(centerness)) cropped to ceil(size / stride)."""
    import argparse
    import numpy as np
    import torch
    from nerf_rpn_b200.model.fcos.fcos import FCOSModule
    args = argparse.Namespace(num_convs=1, norm_reg_targets=True, centerness_on_reg=True, rotated_bbox=True, pre_nms_thresh=0.0, pre_nms_top_n=10,
                              nms_thresh=0.3, fpn_post_nms_top_n=10, min_size=0.0)
    mod = FCOSModule(args, 256, [4, 8, 16, 32])
    g = torch.Generator().manual_seed(0)
    grids = [(10, 12, 8), (5, 6, 4), (3, 3, 2), (2, 2, 1)]
    cls = [torch.randn(2, 1, *d, generator=g) for d in grids]
    ctr = [torch.randn(2, 1, *d, generator=g) for d in grids]
    sizes = [(40, 48, 32), (33, 20, 30)]
    paths = [str(tmp_path / f"s{i}.npz") for i in range(2)]
    mod.output_objectness(cls, ctr, sizes, paths)
    for i, p in enumerate(paths):
        z = np.load(p)
        assert sorted(z.files) == ["0", "1", "2", "3"]
        for l, s in enumerate([4, 8, 16, 32]):
            w, ll, h = [int(np.ceil(v / s)) for v in sizes[i]]
            want = np.sqrt(1 / (1 + np.exp(-cls[l][i, 0].numpy().astype(np.float64))) / (1 + np.exp(-ctr[l][i, 0].numpy().astype(np.float64))))[:w, :ll, :h]
            assert z[str(l)].shape == (w, ll, h)
            np.testing.assert_allclose(z[str(l)], want, rtol=1e-5)
