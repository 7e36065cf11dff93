"""Multi-GPU = replicas only.

A unit of work is one reference view with its source views; units share nothing but the
read-only weights, so the path shards with NO data-path collective: unit u runs on
rank u mod N (the reference does the same split with nn.DataParallel scatter/gather inside one
process, eval.py:118-120).  The only communication is bookkeeping: a barrier around timed regions
and a MAX-reduce of per-rank device times.  Works on any torch.distributed backend (nccl on GPUs,
gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def world() -> tuple:
    """(rank, world_size) -- (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_units(n_units: int, rank: int, world_size: int) -> List[int]:
    """Round-robin assignment: unit u -> rank u mod N (weak scaling: per-rank work fixed per unit)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, n_units, world_size))


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(values: Sequence[float], device=None) -> List[float]:
    """Element-wise MAX of per-rank scalars (device times) over all ranks."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_over_ranks(values: Sequence[float], device=None) -> List[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


def aggregate_throughput(units_this_rank: int, elapsed_ms_this_rank: float, device=None) -> float:
    """Whole-job units/s = (units processed by all ranks) / (max over ranks of the elapsed time)."""
    total_units = sum_over_ranks([units_this_rank], device)[0]
    worst_ms = max_over_ranks([elapsed_ms_this_rank], device)[0]
    return total_units / (worst_ms / 1000.0)
