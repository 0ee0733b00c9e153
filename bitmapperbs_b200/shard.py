"""Multi-GPU bookkeeping of bench.py: the index is replicated, every rank maps its own reads (weak scaling),
there is no collective on the data path.  The only cross-rank communication is control plane -- a barrier and one MAX
reduction of the per-rank times (torch.distributed: NCCL on GPUs, gloo in the CPU tests); the whole-job value is the units
all ranks processed over the slowest rank's time.  (The command line's own sharding -- sub-blocks to whichever GPU thread is
free, results put back in input order -- is C++, csrc/host/bmbs_main.cpp, and is tested through its SAM output.)"""
from __future__ import annotations


def rank_seed(base_seed: int, rank: int) -> int:
    """weak scaling: every rank simulates its own reads"""
    return base_seed + rank


def max_over_ranks(dist, values, device=None):
    """element-wise maximum of `values` (floats) over all ranks; the values themselves without a process group"""
    values = [float(v) for v in values]
    if dist is None or not dist.is_initialized():
        return values
    import torch
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def whole_job_rate(units_per_rank: int, world: int, steps: int, max_ms: float) -> float:
    """units/s of the whole job: every rank did `steps` steps of `units_per_rank`, the slowest took `max_ms`"""
    return units_per_rank * world * steps / (max_ms / 1000.0)
