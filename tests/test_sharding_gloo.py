"""CPU, world_size 2 over gloo: the multi-GPU host logic (batch ownership, ordered merge, max-over-ranks timing)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bitmapperbs_b200 import shard


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.batches_of_rank(7, rank, world)
    assert all(shard.owner_of_batch(b, world) == rank for b in mine)
    results = [(b, f"payload{b}") for b in mine]
    gathered = [None] * world
    dist.all_gather_object(gathered, results)
    merged = shard.merge_in_order(gathered)
    # rank 1 is the slow one: throughput must use ITS time and the SUM of the units
    thr, t, u = shard.aggregate_throughput(dist, units_this_rank=1000 * (rank + 1), seconds_this_rank=0.5 * (rank + 1))
    q.put((rank, mine, merged, thr, t, u, shard.rank_seed(2002, rank)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, merged0, thr0, t0, u0, s0), (r1, m1, merged1, thr1, t1, u1, s1) = out
    assert m0 == [0, 2, 4, 6] and m1 == [1, 3, 5]
    assert merged0 == merged1 == [f"payload{b}" for b in range(7)]
    assert t0 == t1 == 1.0 and u0 == u1 == 3000 and thr0 == thr1 == 3000.0
    assert (s0, s1) == (2002, 2003)


def test_single_process_fallback():
    thr, t, u = shard.aggregate_throughput(None, 10, 2.0)
    assert (thr, t, u) == (5.0, 2.0, 10)
