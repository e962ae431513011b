"""Read sharding across ranks and the end-of-run statistics reduction.

The path shards embarrassingly: every rank holds the full index in its own HBM and aligns a contiguous range of
reads (the reference does the same with threads over input ranges, RangeSplitter.h:14-20).  The only collective is
one all-reduce(SUM) of the AlignerStats counters at the end (AlignerStats.h:41-84) so rank 0 can print them.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of reads for `rank`: rank r gets [r*n/world, (r+1)*n/world)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (n * rank) // world, (n * (rank + 1)) // world


def allreduce_counters(counters: np.ndarray, device=None) -> np.ndarray:
    """Sums an int64 counter vector over all ranks with torch.distributed (NCCL on GPUs, gloo in the CPU tests).
    Returns the reduced vector on every rank."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(counters, dtype=np.int64).copy())
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (timings are reported as the max over ranks)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
