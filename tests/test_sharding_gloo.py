"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py — tiles are partitioned t mod G with no data-path collective,
counts are summed and times max-reduced exactly as the NCCL run does."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tiles, offs, cat = bench.shard(rank, world, 24)
    cnt = torch.tensor([float(offs[-1])], dtype=torch.float64)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, tiles)
    if rank == 0:
        out.put((cnt.item(), t.item(), gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_tile_sharding_two_ranks():
    sys.path.insert(0, ROOT)
    import bench
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, tmax, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(gathered[0] + gathered[1]) == list(range(24))
    assert all(t % 2 == 0 for t in gathered[0]) and all(t % 2 == 1 for t in gathered[1])
    _, offs, _ = bench.shard(0, 1, 24)
    assert total == float(offs[-1]) and tmax == 11.0
    # a shard is a pure function of (tile id): the same tile has the same heads whatever the world size
    a = bench.tile_heads(5)
    tiles1, offs1, cat1 = bench.shard(1, 2, 24)
    g = tiles1.index(5)
    assert np.array_equal(cat1[1][offs1[g]:offs1[g + 1]], a[1])
