"""Batch sharding for multi-GPU inference: images are independent, so ranks split the batch and never exchange
activations; the only collective is the MAX-reduction of the per-rank timings for reporting.
torch.distributed is optional plumbing here (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple


def shard_batch(global_batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[start, start+count) of the images rank `rank` owns: contiguous, sizes differ by at most one."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError('bad rank %d / world size %d' % (rank, world_size))
    base, rem = divmod(global_batch, world_size)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def all_shards(global_batch: int, world_size: int) -> List[Tuple[int, int]]:
    return [shard_batch(global_batch, world_size, r) for r in range(world_size)]


def max_over_ranks(values, device=None):
    """Element-wise MAX of a small list of floats over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def aggregate_throughput(units_per_rank_per_step: int, steps: int, world_size: int, max_ms: float) -> float:
    """Whole-job units/s for weak scaling: every rank processed units_per_rank_per_step * steps in <= max_ms."""
    return units_per_rank_per_step * world_size * steps / (max_ms / 1000.0)
