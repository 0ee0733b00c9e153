"""CPU, world_size 2 over gloo: the multi-GPU bookkeeping bench.py runs under torchrun (bitmapperbs_b200/shard.py) -- per-rank
read seeds, the MAX reduction of the per-rank times, the whole-job rate -- and the reference arm's rule that rank 0 alone
works.  (The command line's sharding over GPUs is C++ and is tested through its SAM output on the GPU box.)"""
import os
import subprocess
import sys

import torch.distributed as dist
import torch.multiprocessing as mp

from bitmapperbs_b200 import shard
from conftest import ROOT


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 1 is the slow one on the device, rank 0 end to end: every rank must end up with the slowest times
    mine = [10.0 * (rank + 1), 50.0 - 10.0 * rank, 7.0]
    got = shard.max_over_ranks(dist, mine)
    q.put((rank, got, shard.rank_seed(2003, rank)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_agree_on_the_slowest_times():
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
    (r0, t0, s0), (r1, t1, s1) = out
    assert t0 == t1 == [20.0, 50.0, 7.0]
    assert (s0, s1) == (2003, 2004)                      # every rank maps its own reads
    # 2 ranks x 20 steps x 1 M reads, slowest rank 125 ms in total -> 320 M reads/s for the whole job
    assert shard.whole_job_rate(1_000_000, 2, 20, 125.0) == 320e6


def test_single_process_needs_no_process_group():
    assert shard.max_over_ranks(None, [3, 4.5]) == [3.0, 4.5]


def test_reference_arm_runs_on_rank_zero_only():
    """bench.py --impl reference under torchrun: ranks other than 0 exit 0 without work or output"""
    env = {**os.environ, "RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
