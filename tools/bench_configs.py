#!/usr/bin/env python
"""Device-resident throughput of the other BASELINE.json configurations (parity-test cases, not the bench line):
  config 1: VGG19-FPN + anchor head, one 32^3 grid
  config 3: Swin-S + FPN + FCOS head (OBB), 200x200x130 grid
  extra   : ResNet50-FPN + FCOS head (OBB), 160x256x256
Random-init weights (seed 0), synthetic U[0,1) grids, CUDA events around K graph replays, inputs resident in HBM."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_rpn_b200.model.anchor import AnchorGenerator3D, RPNHead  # noqa: E402
from nerf_rpn_b200.model.fcos.fcos import FCOSOverNeRF  # noqa: E402
from nerf_rpn_b200.model.feature_extractor import Bottleneck, ResNet_FPN_256, SwinTransformer_FPN, VGG_FPN  # noqa: E402
from nerf_rpn_b200.model.nerf_rpn import NeRFRegionProposalNetwork  # noqa: E402

SIZES = ((8,), (16,), (32,), (64,),)
ASPECT = (((1., 1., 1.), (1., 1., 2.), (1., 2., 2.), (1., 1., 3.), (1., 3., 3.)),) * 4


def fcos_args():
    return argparse.Namespace(num_convs=4, norm_reg_targets=True, centerness_on_reg=True, rotated_bbox=True, pre_nms_thresh=0.0,
                              pre_nms_top_n=2500, nms_thresh=0.3, fpn_post_nms_top_n=2500, min_size=0.0)


def build(name):
    torch.manual_seed(0)
    if name == "config1_vgg19_anchor_32":
        bb = VGG_FPN("EF", 4, True, 32)
        ag = AnchorGenerator3D(SIZES, ASPECT)
        m = NeRFRegionProposalNetwork(bb, ag, RPNHead(256, 13, 4), rpn_pre_nms_top_n_test=2500, rpn_post_nms_top_n_test=2500, rpn_nms_thresh=0.3)
        return m, (32, 32, 32)
    if name == "config3_swin_s_fcos_200x200x130":
        bb = SwinTransformer_FPN(patch_size=[4, 4, 4], embed_dim=96, depths=[2, 2, 18, 2], num_heads=[3, 6, 12, 24], window_size=[4, 4, 4],
                                 stochastic_depth_prob=0.0, expand_dim=True)
        return FCOSOverNeRF(fcos_args(), bb, [4, 8, 16, 32]), (200, 200, 130)
    if name == "resnet50_fcos_160x256x256":
        bb = ResNet_FPN_256(Bottleneck, [3, 4, 6, 3], input_dim=4, is_max_pool=True)
        return FCOSOverNeRF(fcos_args(), bb, [4, 8, 16, 32]), (160, 256, 256)
    raise ValueError(name)


def main():
    out = []
    for name in ("config1_vgg19_anchor_32", "config3_swin_s_fcos_200x200x130", "resnet50_fcos_160x256x256"):
        model, dims = build(name)
        model = model.cuda().eval()
        eng = model.engine()
        g = torch.Generator().manual_seed(1000)
        xs = [torch.rand(*dims, 4, generator=g).permute(3, 0, 1, 2).contiguous().cuda()[None] for _ in range(3)]
        with torch.no_grad():
            for i in range(5):
                plan = eng.forward_device(xs[i % 3])
            torch.cuda.synchronize()
            K = 20
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(K):
                plan = eng.forward_device(xs[i % 3])
            torch.cuda.current_stream().wait_event(plan.done)
            b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / K
        row = dict(config=name, dims=dims, ms_per_scene=ms, scenes_per_s=1000.0 / ms, algorithmic_tflop=plan.algorithmic_flops / 1e12,
                   tflops=plan.algorithmic_flops / (ms * 1e-3) / 1e12, proposals=int(plan.out_count[0].item()), launches=plan.num_launches())
        print(json.dumps(row), flush=True)
        out.append(row)
        del model, eng, plan
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/other_configs.json", "w"), indent=1)


if __name__ == "__main__":
    main()
