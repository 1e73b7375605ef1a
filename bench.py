#!/usr/bin/env python
"""bench.py -- scenes/sec of the NeRF-RPN hot path (BASELINE.json config 2: ResNet50-3D + FPN + anchor head,
160x256x256 RGB-sigma grids, bf16, 13 anchors/location, top-2500 per level, NMS 0.3).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one scene through backbone -> FPN -> head -> decode/top-k -> NMS -> proposals on each rank (weak scaling:
scenes are independent, no collective on the data path).  Prints ONE JSON line (contract in the task
statement): `value` = whole-job scenes/s with inputs resident in HBM (device-timed, max over ranks), `e2e` = the same
metric through the public pipeline with pinned HOST grids (H2D + D2H inside the timed region), `roofline` for the
dominant kernel (tcgen05 implicit-GEMM conv, RPN-head layer over P2..P5) measured live with CUDA events,
`cpu_baseline` = the oracle's fp32 CPU port of the reference on this box's host cores.

`--impl reference` times that CPU port only (the reference is pure Python/PyTorch and cannot travel to the GPU box
together with /root/reference; the port under oracle/net.py restates it and is pinned by its golden vectors).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = (160, 256, 256)
ANCHOR_SIZES = ((8,), (16,), (32,), (64,),)
ASPECT = (((1., 1., 1.), (1., 1., 2.), (1., 2., 2.), (1., 1., 3.), (1., 3., 3.)),) * 4
WORKLOAD = "ResNet50-3D+FPN+anchor-head(AABB) 160x256x256x4 RGBsigma, 13 anchors/loc, pre/post-NMS top 2500, NMS 0.3"
FLOPS_PER_SCENE = 3.913e12        # conv FLOPs (2*MAC) of the reference's layers


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=None, choices=["bf16", "fp16", "fp16_w2"],
                    help="numeric mode (nerf_rpn_b200/precision.py); default fp16_w2 = the mode that meets north_star's <= 1e-3 on feature maps")
    ap.add_argument("--input-layout", default="dataset", choices=["dataset", "ncdhw", "dataset_u8"],
                    help="input grids: the dataset's fp32 channels-last view (default), contiguous fp32 (4,W,L,H), or the raw uint8 "
                         "channels-last view (uint8 npz files; normalised on the device instead of by datasets.py:59-61 on the host)")
    ap.add_argument("--skip-cpu-baseline", action="store_true", help="exploration runs only: omit the ~40 s CPU port timing")
    ap.add_argument("--scenes-per-step", type=int, default=int(os.environ.get("NRPN_SCENES_PER_STEP", "4")),
                    help="scenes per rank per step (one engine launch); weights are read once per step")
    return ap.parse_args()


def synth_scene(i, layout="ncdhw"):
    """Scene i of SURVEY.md 8(d): U[0,1) RGB + alpha, channels-last on disk like the real npz -> (4,W,L,H) fp32.
    layout "dataset": the (4,W,L,H) VIEW of the (W,L,H,4) array, exactly what datasets.py:55-56 hands to the model;
    layout "ncdhw": the same values as a contiguous (4,W,L,H) tensor (what the reference's torch.stack makes of it)."""
    import torch
    g = torch.Generator().manual_seed(1000 + i)
    grid = torch.rand(*DIMS, 4, generator=g)
    return grid.permute(3, 0, 1, 2) if layout == "dataset" else grid.permute(3, 0, 1, 2).contiguous()


def build_modules():
    import torch
    from nerf_rpn_b200.model.anchor import AnchorGenerator3D, RPNHead
    from nerf_rpn_b200.model.feature_extractor import Bottleneck, ResNet_FPN_256
    torch.manual_seed(0)
    backbone = ResNet_FPN_256(Bottleneck, [3, 4, 6, 3], input_dim=4, is_max_pool=True)
    ag = AnchorGenerator3D(ANCHOR_SIZES, ASPECT)
    head = RPNHead(256, ag.num_anchors_per_location()[0], 4, rotate=False)
    return backbone, ag, head


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU port
_CPU_PORT = {}


def _cpu_port_state():
    """Weights, anchors and the synthetic scene of the CPU port, built once per process."""
    if not _CPU_PORT:
        backbone, ag, head = build_modules()
        _CPU_PORT["sd"] = {k: v.detach() for k, v in backbone.state_dict().items()}
        _CPU_PORT["hsd"] = {k: v.detach() for k, v in head.state_dict().items()}
        _CPU_PORT["cells"] = ag.cell_anchors_np()
        _CPU_PORT["scene"] = synth_scene(0)
    return _CPU_PORT


def cpu_port_run(x_extent, n_threads, repeats=1, y_extent=None):
    """Time the oracle's CPU port on an (x_extent x y_extent x 256) block of the scene (x_extent*y_extent/(160*256) of a scene).
    Returns seconds per run."""
    import torch
    from oracle import net as onet
    torch.set_num_threads(n_threads)
    st = _cpu_port_state()
    y_extent = DIMS[1] if y_extent is None else y_extent
    x = st["scene"][:, :x_extent, :y_extent].contiguous()[None]
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        onet.full_forward(st["sd"], st["hsd"], x, st["cells"], False)
        times.append(time.perf_counter() - t0)
    return times


def best_cpu_threads():
    """Intra-op thread count that makes the CPU port fastest on this host: all cores is NOT it on a 128-thread box (measured on
    the B200 host: 32x128x256 block 0.37 s with 16 threads, 0.43 s with 32, 0.80 s with 64, 5.6 s with 128; full scene 4.2 s
    with 32 threads vs 33 s with 128 -- profiles/r01_cpu_port_threads.txt). Sweeps {all, 64, 32, 16, 8} on a small block and
    returns (threads, {threads: seconds})."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (cores, 64, 32, 16, 8) if 1 <= c <= cores}, reverse=True)
    cpu_port_run(32, min(cands), y_extent=64)                  # library warm-up (oneDNN JIT, thread pool), not timed
    sweep = {}
    for th in cands:
        sweep[th] = min(cpu_port_run(32, th, y_extent=128, repeats=2))
    best = min(sweep, key=sweep.get)
    return best, {k: round(v, 2) for k, v in sweep.items()}


def run_reference(args):
    """--impl reference: the CPU port on all host cores; each step = a bounded block of the workload, sized so that the whole
    (warmup + steps) run stays within a few minutes whatever K is."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores, sweep = best_cpu_threads()
    torch.set_num_threads(cores)
    per_full = cpu_port_run(160, cores)[0]                     # one full scene with the chosen thread count
    budget = 150.0
    frac = budget / (per_full * (args.steps + args.warmup))
    # candidate blocks (x extent multiple of 32 for the 5 stride-2 stages, y extent 64/128/256), largest one within the budget
    cands = sorted(((ex * ey) / float(DIMS[0] * DIMS[1]), ex, ey) for ex in (32, 64, 96, 128, 160) for ey in (64, 128, 256))
    pick = cands[0]
    for c in cands:
        if c[0] <= frac:
            pick = c
    share, ex, ey = pick
    for _ in range(args.warmup):
        cpu_port_run(ex, cores, y_extent=ey)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_run(ex, cores, y_extent=ey)
    dt = time.perf_counter() - t0
    scenes = args.steps * share
    value = scenes / dt
    sample = (f"{ex}x{ey}x{DIMS[2]} block per step = {share:.3f} scene (oracle/net.py fp32 port of the reference, torch {torch.__version__}, "
              f"{cores} of {os.cpu_count()} threads = the fastest of the sweep {sweep})")
    out = {"impl": "reference", "metric": "scenes/sec", "value": value, "unit": "scenes/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
           "cpu_baseline": {"value": value, "unit": "scenes/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


def time_dominant_kernel(plan, reps=10):
    """CUDA-event time of the dominant kernel: the first RPN-head layer (3^3 256->256 + ReLU over P2..P5 in one launch)."""
    import torch
    f = plan.head_launches[0]
    f(); torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    times = []
    for _ in range(reps):
        flush.zero_()                       # evict L2 between repetitions
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    vox = plan.n * sum(d[0] * d[1] * d[2] for d in plan.feat_dims)
    flops = 2.0 * vox * 256 * 256 * 27
    return statistics.mean(times), min(times), flops


def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from nerf_rpn_b200._lib import lib
    from nerf_rpn_b200.model.nerf_rpn import NeRFRegionProposalNetwork
    from nerf_rpn_b200.runtime import ScenePipeline

    backbone, ag, head = build_modules()
    model = NeRFRegionProposalNetwork(backbone, ag, head, rpn_pre_nms_top_n_test=2500, rpn_post_nms_top_n_test=2500,
                                      rpn_nms_thresh=0.3, rpn_score_thresh=0.0).cuda().eval()
    from nerf_rpn_b200 import precision as nprec
    args.precision = nprec.resolve(args.precision)
    model.precision = args.precision
    eng = model.engine()
    B = max(1, args.scenes_per_step)
    n_pool = 4                                             # 4 x 168 MB of distinct inputs (> 126 MB L2)
    if args.input_layout == "dataset_u8":   # raw uint8 (W,L,H,4) arrays as stored in uint8 npz files
        host = [(synth_scene(rank * 1000 + i, "dataset").permute(1, 2, 3, 0) * 255.0).round().to(torch.uint8).contiguous().pin_memory()
                .permute(3, 0, 1, 2) for i in range(n_pool)]
    elif args.input_layout == "dataset":    # pinned (W,L,H,4) arrays, handed over as (4,W,L,H) views like the reference's dataset does
        host = [synth_scene(rank * 1000 + i, "dataset").permute(1, 2, 3, 0).contiguous().pin_memory().permute(3, 0, 1, 2) for i in range(n_pool)]
    else:
        host = [synth_scene(rank * 1000 + i).pin_memory() for i in range(n_pool)]
    hdev = [h.cuda() for h in host]
    if args.input_layout in ("dataset", "dataset_u8"):      # keep the (B,X,Y,Z,4) memory order: logical (B,4,X,Y,Z) views
        dev = [torch.stack([hdev[(i + b) % n_pool].permute(1, 2, 3, 0) for b in range(B)], 0).permute(0, 4, 1, 2, 3) for i in range(n_pool)]
    else:
        dev = [torch.stack([hdev[(i + b) % n_pool] for b in range(B)], 0) for i in range(n_pool)]   # (B,4,X,Y,Z) batches
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    with torch.no_grad():
        for i in range(max(W, 4)):                            # >= 4: both buffer parities warmed and captured
            plan = eng.forward_device(dev[i % n_pool])
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            plan = eng.forward_device(dev[i % n_pool])
        torch.cuda.current_stream().wait_event(plan.done)     # the last scene's post-processing (side stream)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        count = int(plan.out_count[0].item())

        # ---- end to end through the streaming pipeline (pinned host grids in, proposals out on the host)
        pipe = ScenePipeline(model, DIMS, batch=B)
        pipe.run([host[i % n_pool] for i in range(W * B)], collect=True)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e2.record()
        res = pipe.run([host[i % n_pool] for i in range(K * B)], collect=True)
        e3.record()
        barrier()
        wall_ms = 1000.0 * (time.perf_counter() - t0)
        ms_e2e = max(e2.elapsed_time(e3), wall_ms)         # host-side collection included
        assert len(res) == K * B

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    out = None
    if rank == 0:
        launches_per_step = plan.num_launches()
        torch.cuda.synchronize()
        k_mean, k_min, k_flops = time_dominant_kernel(plan)
        burst, sustained, how = measured_peaks()
        achieved = k_flops / (k_mean * 1e-3) / 1e12
        # DRAM traffic of that kernel from the committed `ncu --set full` capture (one scene per launch), scaled to this run's
        # scenes per launch; None when the summary is absent.  It is a profile number, not measured in this (unprofiled) run.
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_ncu_full_summary.json")) as f:
                for row in json.load(f):
                    if row.get("capture") == "head":
                        traffic = (row["dram_read_B"] + row["dram_write_B"]) * B
        except (OSError, ValueError, KeyError):
            traffic = None
        roofline = {"bound": "tensor", "kernel": "conv3d_igemm_kernel<256,4> (RPN head layer, 3x3x3 256->256 + bias + ReLU over P2..P5)",
                    "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst, "traffic": traffic,
                    "traffic_source": "profiles/r01_ncu_full_summary.json (dram__bytes_read.sum + dram__bytes_write.sum, 1 scene/launch) x scenes per launch",
                    "peak_source": how + ", burst figure (kernel timed alone, L2 flushed between launches)",
                    "launch_ms": k_mean, "flops_per_launch": k_flops,
                    "whole_step_frac_of_sustained": (FLOPS_PER_SCENE * world * K * B / (ms * 1e-3) / 1e12) / (sustained * world)}
        value = world * K * B / (ms * 1e-3)
        cores = os.cpu_count() or 1
        sweep = {}
        torch.cuda.empty_cache()
        if args.skip_cpu_baseline or world > 1:            # the CPU baseline is timed on rank 0 of the single-GPU run only
            cpu_t = float("nan")
        else:
            cores, sweep = best_cpu_threads()              # all 128 hardware threads are 8x SLOWER than 32 on the B200 host
            cpu_t = statistics.mean(cpu_port_run(160, cores, repeats=3))
        out = {"metric": "scenes/sec", "value": value, "unit": "scenes/s", "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": nprec.bench_dtype(args.precision),
               "data": "synthetic",
               "config": {"workload": WORKLOAD, "scenes_per_step_per_gpu": B, "parallelism": f"dp{world} (one scene per rank, no collective)",
                          "l2": "4 distinct 168 MB input grids per rank cycled (> 126 MB L2); activations stream ~1.5 GB/scene",
                          "weights": "reference init, torch.manual_seed(0)", "proposals_last_scene": count,
                          "input_layout": {"dataset": "fp32 (4,W,L,H) views of (W,L,H,4) arrays, as datasets.py:55-56 yields",
                                           "dataset_u8": "raw uint8 (4,W,L,H) views of (W,L,H,4) arrays, normalised on the device",
                                           "ncdhw": "fp32 contiguous (4,W,L,H)"}[args.input_layout]},
               "clocks": clocks,
               "e2e": {"value": world * K * B / (ms_e2e * 1e-3), "unit": "scenes/s", "h2d_bytes_per_step": pipe.h2d_bytes_per_scene * B,
                       "d2h_bytes_per_step": pipe.d2h_bytes_per_scene * B, "ms_per_step": ms_e2e / K,
                       "api": "nerf_rpn_b200.runtime.ScenePipeline.run (pinned host grids -> host proposals)"},
               "gpu_launches": launches_per_step * K,
               "roofline": roofline,
               "cpu_baseline": {"value": (1.0 / cpu_t) if cpu_t == cpu_t else None, "unit": "scenes/s", "cores": cores, "kind": "port",
                                "sample": f"3 scenes 160x256x256, mean (oracle/net.py fp32 CPU port of the reference; {cores} of {os.cpu_count()} threads = "
                                          f"the fastest of the sweep {sweep})"
                                if cpu_t == cpu_t else "not timed in this run (N > 1 or --skip-cpu-baseline)"}}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
