"""Sharding of read batches over GPUs (index replicated, no data-path collective) -- SURVEY.md §8e.

The host mapper (`bmbs --gpus G`) and bench.py use the same rule: batch b goes to GPU b mod G and results are
merged back by batch sequence number.  The only cross-rank communication is control plane: a barrier and a MAX
reduction of the per-rank device time (torch.distributed, NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def owner_of_batch(batch_no: int, world: int) -> int:
    return batch_no % world


def batches_of_rank(n_batches: int, rank: int, world: int):
    return list(range(rank, n_batches, world))


def merge_in_order(per_rank_results):
    """per_rank_results[r] = list of (batch_no, payload) -> payloads in batch order"""
    merged = sorted((b, p) for res in per_rank_results for b, p in res)
    return [p for _, p in merged]


def rank_seed(base_seed: int, rank: int) -> int:
    """weak scaling: every rank simulates its own reads"""
    return base_seed + rank


def aggregate_throughput(dist, units_this_rank: int, seconds_this_rank: float, device=None):
    """whole-job units/s = sum of units over ranks / max of time over ranks"""
    if dist is None or not dist.is_initialized():
        return units_this_rank / seconds_this_rank, seconds_this_rank, units_this_rank
    import torch
    t = torch.tensor([seconds_this_rank], dtype=torch.float64, device=device)
    u = torch.tensor([float(units_this_rank)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()) / float(t.item()), float(t.item()), int(u.item())
