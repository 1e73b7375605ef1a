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
