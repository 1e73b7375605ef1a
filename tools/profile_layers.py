#!/usr/bin/env python
"""Per-launch CUDA-event timing of one full-size scene (eager launches, L2 not flushed between layers, like the real
step). Prints a table and writes gpurun_out/layers.csv. Run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

from nerf_rpn_b200.model.nerf_rpn import NeRFRegionProposalNetwork  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "anchor"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    if cfg == "anchor":
        backbone, ag, head = bench.build_modules()
        model = NeRFRegionProposalNetwork(backbone, ag, head, rpn_pre_nms_top_n_test=2500, rpn_post_nms_top_n_test=2500,
                                          rpn_nms_thresh=0.3).cuda().eval()
        x = torch.stack([bench.synth_scene(i) for i in range(batch)]).cuda()
    else:                       # a name from tools/bench_configs.py
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import bench_configs
        model, dims = bench_configs.build(cfg)
        model = model.cuda().eval()
        g = torch.Generator().manual_seed(1000)
        x = torch.stack([torch.rand(*dims, 4, generator=g).permute(3, 0, 1, 2).contiguous() for _ in range(batch)]).cuda()
    eng = model.engine()
    eng.use_graph = False
    for _ in range(2):
        plan = eng.forward_device(x)
    torch.cuda.synchronize()
    rows = []
    reps = 5
    for stage, fs in (("backbone", plan.launches), ("head", plan.head_launches + [plan.pred_launch[0]]), ("post", plan._post[0])):
        for f in fs:
            name, fl = plan.names.get(id(f), ("rpn_proposals (top-k, decode, NMS; ~45 kernels)", 0.0))
            ts = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); f(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            t = sorted(ts)[len(ts) // 2]
            rows.append((stage, name, t, fl))
    total = sum(r[2] for r in rows)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/layers_{cfg}_B{batch}.csv", "w") as fo:
        fo.write("stage,layer,ms,share,gflop,tflops\n")
        for stage, name, t, fl in rows:
            fo.write(f"{stage},{name},{t:.4f},{t / total:.4f},{fl / 1e9:.2f},{fl / (t * 1e-3) / 1e12 if t > 0 else 0:.1f}\n")
    for stage, name, t, fl in sorted(rows, key=lambda r: -r[2])[:30]:
        print(f"{t:8.3f} ms {100 * t / total:5.1f}%  {fl / (t * 1e-3) / 1e12 if t > 0 else 0:7.1f} TF/s  {stage:8s} {name}")
    print(f"sum of launches {total:.3f} ms")


if __name__ == "__main__":
    main()
