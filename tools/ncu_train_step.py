#!/usr/bin/env python
"""Target for `ncu --nvtx --nvtx-include "step/" --metrics gpu__time_duration.sum,...`: ONE training step of BASELINE config 4 (ResNet50-FPN + anchor head,
--rotated_bbox, one 160x256x256 scene), after two untimed steps.  Prints the CUDA-event time of three more steps first (outside the range)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

model = bench.build_model(rotated=True, spread=0.0).cuda().train()
eng = model.train_engine(precision="bf16", lr=1e-4, weight_decay=0.01, clip_grad_norm=0.1, reg_loss_weight=5.0)
grid = bench.synth_scene(0, "dataset").permute(1, 2, 3, 0).contiguous().cuda().permute(3, 0, 1, 2)[None]
gt = [bench.planted_boxes(0).cuda()]
for _ in range(4):                       # two eager steps, the graph capture, one replay
    eng.train_step(grid, gt)
torch.cuda.synchronize()
if os.environ.get("NCU_TRAIN_TIME", "1") == "1":
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        eng.train_step(grid, gt)
    b.record(); torch.cuda.synchronize()
    print(f"train step: {a.elapsed_time(b) / 3:.2f} ms", flush=True)
torch.cuda.nvtx.range_push("step")
eng.train_step(grid, gt)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
