#!/usr/bin/env python
"""FCOS training loss at BASELINE config 3's size (200x200x130 grid, strides 4..32 = 94 358 locations per scene, 30 ground-truth boxes):
our two kernels (targets + loss forward / backward through FCOSModule.loss_evaluator) against the staged reference's FCOSLossComputation
(oracle/_ref) on the same GPU.  CUDA events, 3 warm-up + K timed calls, every call on fresh head outputs (3 sets cycled).
    python tools/fcos_loss_bench.py [--no-ref] [--steps K]      -> one JSON line per (head, loss type), also gpurun_out/fcos_loss_bench.json"""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STRIDES = [4, 8, 16, 32]
MESH = (200, 200, 130)


def fcos_args(rotated, loss_type):
    return argparse.Namespace(num_convs=1, norm_reg_targets=True, centerness_on_reg=True, rotated_bbox=rotated, pre_nms_thresh=0.0, pre_nms_top_n=100,
                              nms_thresh=0.3, fpn_post_nms_top_n=100, min_size=0.0, center_sampling_radius=1.5, iou_loss_type=loss_type,
                              use_additional_l1_loss=rotated and loss_type != "smooth_l1", proj2d_loss_weight=0.0)


def inputs(rotated, seed, n_gt=30):
    g = torch.Generator().manual_seed(seed)
    grids = [tuple(int(math.ceil(m / s)) for m in MESH) for s in STRIDES]
    cls = [(torch.randn(1, 1, *gr, generator=g) * 2 - 2).cuda().requires_grad_(True) for gr in grids]
    reg = [torch.cat([torch.rand(1, 6, *gr, generator=g) * 3 + 0.1] + ([torch.randn(1, 2, *gr, generator=g) * 0.3] if rotated else []), 1)
           .cuda().requires_grad_(True) for gr in grids]
    ctr = [torch.randn(1, 1, *gr, generator=g).cuda().requires_grad_(True) for gr in grids]
    ext = torch.rand(n_gt, 3, generator=g) * torch.tensor([90.0, 90.0, 60.0]) + 6.0
    ctrs = torch.rand(n_gt, 3, generator=g) * torch.tensor(MESH, dtype=torch.float32)
    gt = torch.cat([ctrs, ext, (torch.rand(n_gt, 1, generator=g) - 0.5) * math.pi], 1) if rotated else torch.cat([ctrs - ext / 2, ctrs + ext / 2], 1)
    return cls, reg, ctr, [gt.cuda()]


def timed(fn, sets, steps):
    for i in range(3):
        fn(*sets[i % len(sets)])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        fn(*sets[i % len(sets)])
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    from nerf_rpn_b200._lib import lib
    from nerf_rpn_b200.model.fcos.fcos import FCOSModule
    ref = None
    if not a.no_ref:
        from oracle import ref_gpu
        if ref_gpu.available():
            ref = ref_gpu.load()
    rows = []
    for rotated, loss_type in ((False, "iou"), (False, "giou"), (True, "smooth_l1"), (True, "iou")):
        sets = [inputs(rotated, 10 + k) for k in range(3)]
        mod = FCOSModule(fcos_args(rotated, loss_type), 256, STRIDES)
        locs = mod.compute_locations(sets[0][0])
        n_loc = sum(int(p.shape[0]) for p in locs)

        def step(ev):
            def run(cls, reg, ctr, gts):
                l = ev(locs, cls, reg, ctr, gts, None)
                (l[0] + l[1] + l[2]).backward()
                for t in cls + reg + ctr:
                    t.grad = None
            return run
        n0 = lib().nrpn_launch_count()
        ours = timed(step(mod.loss_evaluator), sets, a.steps)
        launches = (lib().nrpn_launch_count() - n0) / (a.steps + 3)
        D = 8 if rotated else 6
        bytes_per_scene = n_loc * (12 + 4 + 4 * D) + n_loc * ((4 + 4 * D + 4) * 2 + 4 + 4 * D + 4)          # targets kernel + loss kernel (outputs, targets, gradients, ct)
        row = dict(head="obb" if rotated else "aabb", iou_loss_type=loss_type, locations=n_loc, n_gt=30, ours_ms=round(ours, 4),
                   our_kernel_launches_per_call=launches, algorithmic_mb=round(bytes_per_scene / 1e6, 2))
        if ref is not None:
            rmod = ref.fcos.FCOSModule(fcos_args(rotated, loss_type), 256, STRIDES).cuda()
            row["reference_gpu_ms"] = round(timed(step(rmod.loss_evaluator), sets, max(3, a.steps // 4)), 3)
            row["speedup"] = round(row["reference_gpu_ms"] / ours, 1)
        print(json.dumps(row), flush=True)
        rows.append(row)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "fcos_loss_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
