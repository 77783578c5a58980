"""Pair-level sharding of independent registrations over ranks (one process per GPU).

Registrations are independent (loop-closure candidates, odometry pairs), so there is no data-path collective: every rank
aligns its contiguous shard, the elapsed time is the max over ranks, results are gathered on rank 0, which owns the graph
(ScanSensor.cpp:157-166 inserts edges from one thread).  Works with NCCL (GPU) and gloo (CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of rank `rank`; same formula as s3d_gicp_align_batch uses across devices."""
    lo = n_items * rank // world
    hi = n_items * (rank + 1) // world
    return lo, hi


def max_over_ranks(value, device="cpu"):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_to_rank0(local_items):
    """Concatenates every rank's list of picklable results in rank order on rank 0 (None elsewhere)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local_items)
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(list(local_items), out, dst=0)
    if dist.get_rank() != 0:
        return None
    return [x for part in out for x in part]


def align_sharded(align_batch, sources, targets, guesses, params):
    """Runs `align_batch(sources, targets, guesses, params)` on this rank's shard; returns (lo, hi, results)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(len(sources), rank, world)
    g = None if guesses is None else guesses[lo:hi]
    res = align_batch(sources[lo:hi], targets[lo:hi], g, params) if hi > lo else []
    return lo, hi, res
