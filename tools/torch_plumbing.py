"""torch.distributed plumbing for bench.py and the tests (NOT part of the product package, which never imports torch): max over
ranks of the timings, all-reduce of a SyncBN statistics vector, and the single-layer SyncBN forward the first-round tests drive."""
from __future__ import annotations


def max_over_ranks(values, device=None):
    """Element-wise MAX of a small list of floats over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]



def allreduce_stats(stats):
    """In-place SUM all-reduce of a stats tensor [2C+1] = sum_x | sum_x2 | rows (torch tensor: cuda -> NCCL, cpu -> gloo).
    Identity when torch.distributed is not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats



def sync_batch_norm_forward(x, gamma, beta, eps: float = 1e-5, relu: bool = True):
    """Training-mode SyncBN of a CUDA bf16 tensor x [..., C] (NHWC) on this rank: libdlv3p statistics kernel ->
    all-reduce over the process group -> libdlv3p normalisation kernel, all on the current CUDA stream.
    Returns (y bf16 like x, stats fp32 [2C+1] after the all-reduce)."""
    import torch
    from dlv3p_b200 import ffi
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
