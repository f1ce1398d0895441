"""Batch sharding for multi-GPU inference: images are independent, so ranks split the batch and never exchange
activations; plus the host arithmetic of SyncBatchNormalization statistics.  Pure Python / numpy: the process-group
plumbing the benchmarks and tests use (max over ranks, all-reduce of a statistics vector) lives outside the package, in
tools/torch_plumbing.py."""
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


def aggregate_throughput(units_per_rank_per_step: int, steps: int, world_size: int, max_ms: float) -> float:
    """Whole-job units/s for weak scaling: every rank processed units_per_rank_per_step * steps in <= max_ms."""
    return units_per_rank_per_step * world_size * steps / (max_ms / 1000.0)


# ---------------------------------------------------------------------------------------------------------------------
# SyncBatchNormalization statistics (reference layers.py:63-70 with MirroredStrategy, train.py:143-158): the FORWARD half
# of the cfg-5 exchange.  Every replica reduces its own rows to [sum_x | sum_x2 | row count], the replicas all-reduce (SUM)
# that one small vector (<= 16 KB per layer: latency bound, NCCL LL protocol over NVLink), and normalise with the global
# mean / biased variance.  The backward half ([sum g | sum g*xhat]) and the gradient all-reduce live in train.py (HeadTrainer).
def moments_from_stats(stats, C: int):
    """(mean, biased variance) from a reduced stats vector (numpy or torch; SURVEY.md §8(c): var = E[x^2] - mean^2)."""
    n = stats[2 * C]
    mean = stats[:C] / n
    var = stats[C:2 * C] / n - mean * mean
    return mean, var.clip(0) if hasattr(var, 'clip') else var.clamp(min=0)


def update_moving(moving_mean, moving_var, mean, var, momentum: float = 0.99):
    """Keras moving statistics: moving <- moving * momentum + batch * (1 - momentum) (momentum 0.99 default, biased variance)."""
    return moving_mean * momentum + mean * (1.0 - momentum), moving_var * momentum + var * (1.0 - momentum)
