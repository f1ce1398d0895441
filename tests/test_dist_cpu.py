"""N>1 host logic on CPU: world_size-2 gloo process group (no GPU): batch sharding covers the batch exactly once and
the timing reduction is a MAX over ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dlv3p_b200 import sharding


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
    ms = sharding.max_over_ranks([10.0 + rank, 5.0 - rank])
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
    assert sharding.max_over_ranks([1.5, 2.5]) == [1.5, 2.5]   # no process group: identity
