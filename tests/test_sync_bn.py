"""SyncBatchNormalization statistics exchange (the forward half of the cfg-5 training step):
oracle vs the CUDA statistics / normalisation kernels, and the all-reduce logic at world size 2 (gloo on the CPU here,
NCCL on two GPUs when present)."""
import os
import socket

import numpy as np
import pytest

from oracle import head_ref as R
from tools import torch_plumbing


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from dlv3p_b200 import sharding
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    C = 48
    shards = [rng.standard_normal((3 + r, 4, 5, C)).astype(np.float32) for r in range(world)]   # ragged replicas
    st = torch.from_numpy(R.bn_train_stats(shards[rank]))
    torch_plumbing.allreduce_stats(st)
    mean, var = sharding.moments_from_stats(st.numpy(), C)
    q.put((rank, mean, var))
    dist.destroy_process_group()


def test_stats_allreduce_world2_gloo():
    """Each replica reduces its own (ragged) shard; after the SUM all-reduce every rank holds the GLOBAL moments."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    rng = np.random.default_rng(5)
    shards = [rng.standard_normal((3 + r, 4, 5, 48)).astype(np.float32) for r in range(2)]
    allx = np.concatenate([s.reshape(-1, 48) for s in shards]).astype(np.float64)
    for _, mean, var in got:
        assert np.allclose(mean, allx.mean(0), atol=1e-12)
        assert np.allclose(var, allx.var(0), atol=1e-12)            # biased variance


def test_moving_statistics_update():
    from dlv3p_b200 import sharding
    mm, mv = sharding.update_moving(np.zeros(3), np.ones(3), np.array([1.0, 2.0, 3.0]), np.array([4.0, 4.0, 4.0]))
    assert np.allclose(mm, [0.01, 0.02, 0.03]) and np.allclose(mv, [1.03, 1.03, 1.03])    # momentum 0.99 (Keras default)


def test_oracle_sync_bn_equals_single_replica_bn():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((4, 6, 6, 16)).astype(np.float32)
    g, b = rng.uniform(0.5, 1.5, 16), rng.normal(0, 0.1, 16)
    whole, m, v = R.sync_bn_train([x], g, b)
    parts, m2, v2 = R.sync_bn_train([x[:1], x[1:]], g, b)
    assert np.allclose(np.concatenate(parts), whole[0], atol=1e-12) and np.allclose(m, m2) and np.allclose(v, v2)


@pytest.mark.gpu
@pytest.mark.parametrize('M,C,relu', [(32 * 128 * 128 // 8, 256, True), (1000, 48, False), (77, 304, True), (5, 8, True)])
def test_bn_stats_and_apply_match_the_oracle(gpu, M, C, relu):
    from dlv3p_b200 import ffi
    rng = np.random.default_rng(M + C)
    x = R.bf16_round((rng.standard_normal((M, C)) * 1.5 + 0.3).astype(np.float32))
    g, b = rng.uniform(0.5, 1.5, C).astype(np.float32), rng.normal(0, 0.1, C).astype(np.float32)
    y_bits, stats = ffi.op_bn_train(R.to_bf16_bits(x), g, b, 1e-5, relu)
    ref = R.bn_train_stats(x)
    assert stats[2 * C] == M
    assert np.allclose(stats[:C], ref[:C], rtol=2e-5, atol=2e-3 * np.sqrt(M))          # fp32 tree sums of M terms
    assert np.allclose(stats[C:2 * C], ref[C:2 * C], rtol=2e-5, atol=2e-3 * np.sqrt(M))
    outs, _, _ = R.sync_bn_train([x], g, b, 1e-5, relu)
    y = ffi.bf16_bits_to_f32(y_bits)
    assert np.abs(y - outs[0]).max() <= 2.0 ** -8 * max(1.0, np.abs(outs[0]).max())      # one bf16 rounding of the output


def _nccl_worker(rank, world, port, q, one_gpu=False):
    import torch
    import torch.distributed as dist
    from dlv3p_b200 import sharding
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    if one_gpu:     # both replicas on cuda:0, gloo carries the CUDA statistics vector (NCCL refuses two ranks on one device)
        torch.cuda.set_device(0)
        dist.init_process_group('gloo', rank=rank, world_size=world)
    else:
        torch.cuda.set_device(rank)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    rng = np.random.default_rng(9)
    C = 256
    shards = [R.bf16_round(rng.standard_normal((2 + r, 16, 16, C)).astype(np.float32)) for r in range(world)]
    g, b = rng.uniform(0.5, 1.5, C).astype(np.float32), rng.normal(0, 0.1, C).astype(np.float32)
    x = torch.from_numpy(shards[rank]).cuda().to(torch.bfloat16)
    y, stats = torch_plumbing.sync_batch_norm_forward(x, torch.from_numpy(g), torch.from_numpy(b))
    torch.cuda.synchronize()
    q.put((rank, y.float().cpu().numpy(), stats.cpu().numpy()))
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sync_bn_forward_two_gpus_nccl(gpu):
    """Two replicas with different row counts: statistics kernel -> all-reduce -> normalisation kernel.  NCCL over NVLink on a box
    with two GPUs; on a single-GPU box the same exchange runs with both replicas on cuda:0 over gloo (never skipped)."""
    import torch
    one_gpu = torch.cuda.device_count() < 2
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q, one_gpu)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted([q.get(timeout=300) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
    rng = np.random.default_rng(9)
    C = 256
    shards = [R.bf16_round(rng.standard_normal((2 + r, 16, 16, C)).astype(np.float32)) for r in range(2)]
    g, b = rng.uniform(0.5, 1.5, C).astype(np.float32), rng.normal(0, 0.1, C).astype(np.float32)
    outs, _, _ = R.sync_bn_train(shards, g, b)
    for (rank, y, stats), ref in zip(got, outs):
        assert stats[2 * C] == sum(s.shape[0] * 256 for s in shards)
        assert np.abs(y - ref).max() <= 2.0 ** -8 * max(1.0, np.abs(ref).max())
