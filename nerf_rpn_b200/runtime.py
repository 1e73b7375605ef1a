"""Scene streaming runtime: host (pinned) RGB-sigma grids -> proposals, with the host->device copy of scene i+1
overlapped with the compute of scene i on a second CUDA stream, and the small device->host result copy issued
asynchronously.  This is the end-to-end entry point bench.py times (`e2e`): what run_rpn.py's eval loop does per
batch (run_rpn.py:470-517: .cuda(), model(...), .cpu()), minus its per-iteration synchronisations.

Scenes shard across ranks by index (rank r takes scenes r, r+world, ...): inference needs no collective
(SURVEY.md section 8e).
"""
from typing import Iterable, List, Optional, Tuple

import torch


class ScenePipeline:
    def __init__(self, model, dims: Tuple[int, int, int], device=None, batch: int = 1):
        """model: nerf_rpn_b200.model.nerf_rpn.NeRFRegionProposalNetwork in eval mode; dims: (W, L, H) of every scene;
        batch: scenes per engine launch (weights are read once per batch, small layers get more tiles)."""
        self.model = model
        self.batch = int(batch)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.dims = tuple(dims)
        self.eng = model.engine()
        self.copy_stream = torch.cuda.Stream(device=self.device)
        # two device staging buffers; the engine's graph reads its own static input, filled by a D2D copy.  Allocated on the
        # first run() in the memory order of the host grids (contiguous (4,W,L,H), or the dataset's channels-last view).
        self.stage = None
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        k = self.eng.post_n
        bd = 7 if self.eng.rotated else 6
        B = self.batch
        self.h_boxes = torch.empty((2, B, k, bd), dtype=torch.float32).pin_memory()
        self.h_scores = torch.empty((2, B, k), dtype=torch.float32).pin_memory()
        self.h_levels = torch.empty((2, B, k), dtype=torch.float32).pin_memory()
        self.h_count = torch.empty((2, B), dtype=torch.int32).pin_memory()
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.h2d_bytes_per_scene = 4 * self.dims[0] * self.dims[1] * self.dims[2] * 4
        self.d2h_bytes_per_scene = (k * bd + 2 * k + 1) * 4

    def _alloc_stage(self, sample: torch.Tensor):
        cl = sample.dim() == 4 and not sample.is_contiguous() and sample.permute(1, 2, 3, 0).is_contiguous()
        dt = sample.dtype                   # fp32, or raw uint8 (normalised by the stem packing kernel on the device)
        if cl:
            self.stage = [torch.empty((self.batch, *self.dims, 4), dtype=dt, device=self.device).permute(0, 4, 1, 2, 3)
                          for _ in range(2)]
        else:
            self.stage = [torch.empty((self.batch, 4, *self.dims), dtype=dt, device=self.device) for _ in range(2)]
        self.stage_channels_last = cl
        self.stage_dtype = dt
        self.h2d_bytes_per_scene = 4 * self.dims[0] * self.dims[1] * self.dims[2] * sample.element_size()

    def _prefetch(self, slot: int, host_grids):
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])
            for b, g in enumerate(host_grids):
                self.stage[slot][b].copy_(g, non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def run(self, host_grids: Iterable[torch.Tensor], collect: bool = True):
        """host_grids: iterable of pinned fp32 (4,W,L,H) tensors -- contiguous, or the dataset's views of (W,L,H,4) arrays (then the
        H2D copy is a plain memcpy of the on-disk layout and the stem packing reads it channels-last). Returns a list of (boxes, scores, levels) CPU tensors
        (or only the number of scenes processed when collect=False)."""
        if self.model.engine() is not self.eng:            # hyper-parameters / precision of the model changed: its engine was rebuilt (engine() key)
            raise RuntimeError("nerf_rpn_b200: the model's engine changed since this ScenePipeline was built (its pinned result buffers are sized "
                               "for the old post_nms_top_n / box type): create a new ScenePipeline")
        cur = torch.cuda.current_stream(self.device)
        grids = list(host_grids)
        if len(grids) % self.batch:
            raise ValueError(f"number of scenes ({len(grids)}) must be a multiple of the pipeline batch ({self.batch})")
        it = iter([grids[i:i + self.batch] for i in range(0, len(grids), self.batch)])
        nxt = next(it, None)
        if nxt is None:
            return []
        cl = grids[0].dim() == 4 and not grids[0].is_contiguous() and grids[0].permute(1, 2, 3, 0).is_contiguous()
        if self.stage is None or cl != self.stage_channels_last or grids[0].dtype != self.stage_dtype:
            self._alloc_stage(grids[0])
        for e in self.consumed:
            e.record(cur)
        self._prefetch(0, nxt)
        results: List = []
        pending: Optional[int] = None
        i = 0
        while nxt is not None:
            slot = i & 1
            nxt = next(it, None)
            if nxt is not None:
                self._prefetch(slot ^ 1, nxt)
            cur.wait_event(self.ready[slot])
            plan = self.eng.forward_device(self.stage[slot])
            self.consumed[slot].record(cur)
            # results of the previous scene were copied out while this one was being enqueued
            if pending is not None and collect:
                results.extend(self._collect(pending))
            with torch.cuda.stream(plan.side):                    # ordered after this batch's post-processing
                self.h_boxes[slot].copy_(plan.out_boxes, non_blocking=True)
                self.h_scores[slot].copy_(plan.out_scores, non_blocking=True)
                self.h_levels[slot].copy_(plan.out_levels, non_blocking=True)
                self.h_count[slot].copy_(plan.out_count, non_blocking=True)
                self.done[slot].record(plan.side)
            pending = slot
            i += 1
        if pending is not None and collect:
            results.extend(self._collect(pending))
        else:
            self.done[(i - 1) & 1].synchronize()
        return results if collect else i * self.batch

    def _collect(self, slot: int):
        self.done[slot].synchronize()
        out = []
        for b in range(self.batch):
            k = int(self.h_count[slot, b])
            out.append((self.h_boxes[slot, b, :k].clone(), self.h_scores[slot, b, :k].clone(), self.h_levels[slot, b, :k].clone()))
        return out


# ---------------------------------------------------------------------------------------------- multi-GPU
def shard_indices(n_scenes: int, rank: int, world: int) -> List[int]:
    """Scene i is processed by rank i % world (the reference shards the same way through DistributedSampler with
    batch_size // world_size scenes per rank, run_rpn.py:336-339). No data-path collective is needed at inference."""
    return list(range(rank, n_scenes, world))


def max_over_ranks(value: float, device=None) -> float:
    """Elapsed time of a multi-rank job = the slowest rank (bench.py timing rule)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_to_rank0(local: dict):
    """Collect {scene index: result} dicts on rank 0 (the reference evaluates on rank 0 only, run_rpn.py:359-363)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local, out, dst=0)
    if dist.get_rank() != 0:
        return None
    merged = {}
    for d in out:
        merged.update(d)
    return merged
