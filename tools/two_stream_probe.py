#!/usr/bin/env python
"""Probe: does running two engine instances on two CUDA streams (S scenes per launch each) beat one engine with 2 S scenes per launch?
The small late-stage layers (layer3 / layer4 / laterals: a few hundred voxels) leave most SMs idle; a second stream could fill them."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def run(streams, per, K=40):
    models = [bench.build_model().cuda().eval() for _ in range(streams)]
    xs = [torch.stack([bench.synth_scene(i + 4 * s, "dataset").permute(1, 2, 3, 0).contiguous().cuda() for i in range(per)], 0).permute(0, 4, 1, 2, 3)
          for s in range(streams)]
    sts = [torch.cuda.Stream() for _ in range(streams)]
    engs = [m.engine() for m in models]
    with torch.no_grad():
        for _ in range(5):
            for s in range(streams):
                with torch.cuda.stream(sts[s]):
                    plan = engs[s].forward_device(xs[s])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        plans = []
        for _ in range(K):
            for s in range(streams):
                with torch.cuda.stream(sts[s]):
                    plans.append(engs[s].forward_device(xs[s]))
        for p in plans[-streams:]:
            torch.cuda.current_stream().wait_event(p.done)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{streams} stream(s) x {per} scenes per launch: {K * streams * per / dt:.1f} scenes/s", flush=True)
    del models, engs, xs
    torch.cuda.empty_cache()


run(1, 4)
run(2, 2)
run(2, 4)
run(1, 4)
run(4, 1)
