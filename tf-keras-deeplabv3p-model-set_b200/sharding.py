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


# ---------------------------------------------------------------------------------------------------------------------
# SyncBatchNormalization statistics (reference layers.py:63-70 with MirroredStrategy, train.py:143-158): the FORWARD half
# of the cfg-5 exchange.  Every replica reduces its own rows to [sum_x | sum_x2 | row count], the replicas all-reduce (SUM)
# that one small vector (<= 16 KB per layer: latency bound, NCCL LL protocol over NVLink), and normalise with the global
# mean / biased variance.  The backward half ([sum g | sum g*xhat]) and the gradient all-reduce live in train.py (HeadTrainer).
def allreduce_stats(stats):
    """In-place SUM all-reduce of a stats tensor [2C+1] = sum_x | sum_x2 | rows (torch tensor: cuda -> NCCL, cpu -> gloo).
    Identity when torch.distributed is not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def moments_from_stats(stats, C: int):
    """(mean, biased variance) from a reduced stats vector (numpy or torch; SURVEY.md §8(c): var = E[x^2] - mean^2)."""
    n = stats[2 * C]
    mean = stats[:C] / n
    var = stats[C:2 * C] / n - mean * mean
    return mean, var.clip(0) if hasattr(var, 'clip') else var.clamp(min=0)


def update_moving(moving_mean, moving_var, mean, var, momentum: float = 0.99):
    """Keras moving statistics: moving <- moving * momentum + batch * (1 - momentum) (momentum 0.99 default, biased variance)."""
    return moving_mean * momentum + mean * (1.0 - momentum), moving_var * momentum + var * (1.0 - momentum)


def sync_batch_norm_forward(x, gamma, beta, eps: float = 1e-5, relu: bool = True):
    """Training-mode SyncBN of a CUDA bf16 tensor x [..., C] (NHWC) on this rank: libdlv3p statistics kernel ->
    all-reduce over the process group -> libdlv3p normalisation kernel, all on the current CUDA stream.
    Returns (y bf16 like x, stats fp32 [2C+1] after the all-reduce)."""
    import torch
    from . import ffi
    C = x.shape[-1]
    M = x.numel() // C
    dev = x.device.index or 0
    stream = torch.cuda.current_stream(x.device).cuda_stream
    stats = torch.empty(2 * C + 1, dtype=torch.float32, device=x.device)
    scratch = torch.empty(ffi.bn_scratch_bytes(C), dtype=torch.uint8, device=x.device)
    ffi.bn_stats(x.data_ptr(), M, C, stats.data_ptr(), scratch.data_ptr(), stream, dev)
    allreduce_stats(stats)                                        # NCCL orders itself after the kernels on this stream
    y = torch.empty_like(x)
    g = gamma.to(device=x.device, dtype=torch.float32).contiguous()
    b = beta.to(device=x.device, dtype=torch.float32).contiguous()
    ffi.bn_apply(x.data_ptr(), M, C, stats.data_ptr(), g.data_ptr(), b.data_ptr(), eps, relu, y.data_ptr(), stream, dev)
    return y, stats
