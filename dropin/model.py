"""`import model` shim: makes the reference's bare `from model.xxx import ...` (run_rpn.py:17-22) resolve to
nerf_rpn_b200.model when this directory precedes /root/reference/nerf_rpn on sys.path."""
import sys

import nerf_rpn_b200.model as _m
import nerf_rpn_b200.model.anchor  # noqa: F401
import nerf_rpn_b200.model.feature_extractor  # noqa: F401
import nerf_rpn_b200.model.nerf_rpn  # noqa: F401
import nerf_rpn_b200.model.rpn  # noqa: F401
import nerf_rpn_b200.model.utils  # noqa: F401
import nerf_rpn_b200.model.rotated_iou.oriented_iou_loss  # noqa: F401
import nerf_rpn_b200.model.fpn  # noqa: F401
import nerf_rpn_b200.model.fcos.fcos  # noqa: F401

sys.modules["model"] = _m
for _name, _mod in list(sys.modules.items()):
    if _name.startswith("nerf_rpn_b200.model."):
        sys.modules["model." + _name[len("nerf_rpn_b200.model."):]] = _mod
