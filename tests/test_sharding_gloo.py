"""CPU, world_size 2 over gloo: the N>1 host logic (scene sharding, max-over-ranks timing, rank-0 gather)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf_rpn_b200.runtime import gather_to_rank0, max_over_ranks, shard_indices
    mine = shard_indices(7, rank, world)
    local = {i: (i * i, rank) for i in mine}
    elapsed = max_over_ranks(10.0 + 5.0 * rank)
    merged = gather_to_rank0(local)
    if rank == 0:
        q.put((mine, elapsed, merged))
    else:
        q.put((mine, elapsed, None))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = sorted(g[0] for g in got)
    assert shards == [[0, 2, 4, 6], [1, 3, 5]]                  # every scene exactly once, round-robin
    assert all(abs(g[1] - 15.0) < 1e-9 for g in got)             # both ranks agree on the slowest rank's time
    merged = [g[2] for g in got if g[2] is not None][0]
    assert sorted(merged) == list(range(7)) and merged[3] == (9, 1) and merged[4] == (16, 0)


def _grad_worker(rank, world, port, q):
    """The training step's data-parallel half on CPU tensors: every rank holds a flat gradient bucket, the ranges of
    nerf_rpn_b200.train.bucket_schedule are all-reduced one by one (as the backward pass would launch them), and the result must equal
    ONE all-reduce of the whole bucket; the 1/world factor is applied afterwards like the fused optimiser does."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf_rpn_b200.train import bucket_schedule
    n = 100_003
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(n, generator=g)
    whole = flat.clone()
    dist.all_reduce(whole)
    # stage offsets like a ResNet: stem at 0, blocks ascending, FPN, head last (largest offsets)
    bwd_lo = [0, 5_000, 9_000, 20_000, 21_000, 50_000, 77_000, 90_000]
    sched = bucket_schedule(bwd_lo, n, bucket_elems=15_000)
    covered = torch.zeros(n, dtype=torch.int32)
    prev_k = -1
    for k, lo, hi in sched:
        assert k > prev_k and 0 <= lo < hi <= n
        prev_k = k
        covered[lo:hi] += 1
        dist.all_reduce(flat[lo:hi])
    ok = bool((covered == 1).all()) and torch.equal(flat, whole)
    # ranges launch only once >= bucket_elems elements are final, except the last one
    sizes = [hi - lo for _, lo, hi in sched]
    q.put((ok, sizes, [k for k, _, _ in sched]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_bucket_allreduce_schedule():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for ok, sizes, stages in got:
        assert ok
        assert all(s >= 15_000 for s in sizes[:-1]) and sum(sizes) == 100_003
        assert stages == sorted(stages) and stages[-1] == 7          # the last range goes out after the stem's backward (execution index 7)
