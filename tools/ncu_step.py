#!/usr/bin/env python
"""Target for a metrics-limited ncu pass over ONE full-size scene (eager launches, every kernel of the step):
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,\
lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --nvtx --nvtx-include "step/" ...
Config via argv[1]: anchor (default) | config3_swin_s_fcos_200x200x130 | resnet50_fcos_160x256x256 | config1_vgg19_anchor_32."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench  # noqa: E402
from nerf_rpn_b200.model.nerf_rpn import NeRFRegionProposalNetwork  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "anchor"
if cfg == "anchor":
    backbone, ag, head = bench.build_modules()
    model = NeRFRegionProposalNetwork(backbone, ag, head, rpn_pre_nms_top_n_test=2500, rpn_post_nms_top_n_test=2500,
                                      rpn_nms_thresh=0.3).cuda().eval()
    x = bench.synth_scene(0).cuda()[None]
else:
    import bench_configs
    model, dims = bench_configs.build(cfg)
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(1000)
    x = torch.rand(*dims, 4, generator=g).permute(3, 0, 1, 2).contiguous().cuda()[None]
eng = model.engine()
eng.use_graph = False
with torch.no_grad():
    plan = eng.forward_device(x)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("step")
    plan = eng.forward_device(x)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
