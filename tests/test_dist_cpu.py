"""N>1 host logic on CPU: world_size-2 gloo process group (no GPU): batch sharding covers the batch exactly once and
the timing reduction is a MAX over ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dlv3p_b200 import sharding
from tools import torch_plumbing


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, global_batch, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    start, count = sharding.shard_batch(global_batch, world, rank)
    # every rank "processes" its shard: mark ownership, then check the union over ranks
    owned = torch.zeros(global_batch, dtype=torch.int32)
    owned[start:start + count] = 1
    dist.all_reduce(owned)
    ms = torch_plumbing.max_over_ranks([10.0 + rank, 5.0 - rank])
    q.put((rank, start, count, owned.tolist(), ms))
    dist.destroy_process_group()


@pytest.mark.parametrize('global_batch', [32, 33, 7])
def test_shards_partition_the_batch_and_times_reduce_with_max(global_batch):
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, global_batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, start, count, owned, ms in out:
        assert owned == [1] * global_batch            # each image owned by exactly one rank
        assert ms == [11.0, 5.0]                      # max over ranks of (10+rank, 5-rank)
    sizes = sorted(c for _, _, c, _, _ in out)
    assert sum(sizes) == global_batch and sizes[-1] - sizes[0] <= 1


def test_shard_arithmetic_and_throughput():
    assert sharding.all_shards(32, 8) == [(4 * r, 4) for r in range(8)]
    assert sharding.all_shards(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]
    with pytest.raises(ValueError):
        sharding.shard_batch(8, 2, 2)
    # weak scaling: 32 images per rank per step, 20 steps, 8 ranks, slowest rank 25 ms -> 204800 img/s
    assert sharding.aggregate_throughput(32, 20, 8, 25.0) == pytest.approx(204800.0)
    assert torch_plumbing.max_over_ranks([1.5, 2.5]) == [1.5, 2.5]   # no process group: identity


# ---------------------------------------------------------------------------------------------------- training step, N > 1 host logic
def _train_exchange_worker(rank, world, port, q):
    """The collectives of one training step on CPU tensors over gloo, through the very spans HeadTrainer all-reduces on the GPU:
    grouped SyncBN forward statistics, grouped SyncBN backward sums (= BN gradients), one gradient bucket."""
    import numpy as np
    from dlv3p_b200.train import TrainLayout
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    L = TrainLayout(64, 32, 21)
    rng = np.random.default_rng(100 + rank)
    stats = torch.zeros(L.nstats, dtype=torch.float64)
    shards = {}
    for name, (o, C) in L.stat_off.items():
        x = rng.standard_normal((5 + 3 * rank, C))            # ragged replicas: 5 and 8 rows
        shards[name] = x
        stats[o:o + C] = torch.from_numpy(x.sum(0))
        stats[o + C:o + 2 * C] = torch.from_numpy((x * x).sum(0))
        stats[o + 2 * C] = x.shape[0]
    for grp in L.FWD_GROUPS:
        b, e = L.stats_span(grp)
        dist.all_reduce(stats[b:e])
    grads = torch.full((L.nparams,), float(rank + 1), dtype=torch.float64)
    for grp in L.BWD_GROUPS:
        b, e = L.bn_grad_span(grp)
        dist.all_reduce(grads[b:e])
    b, e = L.bucket_span()
    dist.all_reduce(grads[b:e])
    q.put((rank, stats.numpy(), grads.numpy(), {k: v for k, v in shards.items()}))
    dist.destroy_process_group()


def test_training_step_exchanges_world_size_2_gloo():
    import numpy as np
    from dlv3p_b200.train import TrainLayout
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    L = TrainLayout(64, 32, 21)
    # spans: every BN layer in exactly one forward group and one backward group, spans disjoint, BN gradients outside the bucket
    fwd = [n for g in L.FWD_GROUPS for n in g]
    bwd = [n for g in L.BWD_GROUPS for n in g]
    assert sorted(fwd) == sorted(bwd) == sorted(n for n, _ in L._bn_specs()) and len(set(fwd)) == 14
    spans = [L.stats_span(g) for g in L.FWD_GROUPS]
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= L.nstats
    gsp = [L.bn_grad_span(g) for g in L.BWD_GROUPS]
    assert gsp[0][0] == L.bucket_span()[1] == L.endB and all(a[1] <= b[0] for a, b in zip(gsp, gsp[1:])) and gsp[-1][1] <= L.nparams
    assert len(L.FWD_GROUPS) + len(L.BWD_GROUPS) == 14          # collectives for SyncBN per step (28 layer by layer)
    (_, s0, g0, x0), (_, s1, g1, x1) = out
    assert np.array_equal(s0, s1) and np.array_equal(g0, g1)    # replicas agree after the exchanges
    assert np.all(g0 == 3.0)                                    # 1 + 2: every gradient element summed exactly once
    for name, (o, C) in L.stat_off.items():
        x = np.concatenate([x0[name], x1[name]])                # SyncBN: moments of the GLOBAL batch
        n = s0[o + 2 * C]
        assert n == x.shape[0] == 13
        mean = s0[o:o + C] / n
        var = s0[o + C:o + 2 * C] / n - mean * mean
        assert np.allclose(mean, x.mean(0)) and np.allclose(var, x.var(0))
