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
