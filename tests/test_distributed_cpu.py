"""world_size-2 gloo test (CPU) of the multi-GPU host logic: static sharding + final result gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnngls_b200 import distributed as gd


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 100000, 100003):
        for world in (1, 2, 3, 8):
            spans = [gd.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, n, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = gd.shard_range(total, rank, world)
        g = torch.Generator().manual_seed(0)
        all_tours = torch.randint(0, n, (total, n + 1), generator=g, dtype=torch.int32)
        all_costs = torch.rand(total, generator=g, dtype=torch.float64)
        tours, costs = gd.gather_results(all_tours[lo:hi].clone(), all_costs[lo:hi].clone(), total)
        ok = torch.equal(tours, all_tours) and torch.equal(costs, all_costs)
        tours2, costs2 = gd.gather_results(all_tours[lo:hi].clone(), all_costs[lo:hi].clone())   # total inferred
        ok = ok and torch.equal(tours2, all_tours) and torch.equal(costs2, all_costs)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('total', [10, 11])
def test_gather_results_world2_gloo(total):
    world, n = 2, 6
    ctx = mp.get_context('spawn')
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, total, n, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert dict(ret) == {0: True, 1: True}


def test_single_process_passthrough():
    t, c = torch.zeros(3, 5, dtype=torch.int32), torch.zeros(3, dtype=torch.float64)
    a, b = gd.gather_results(t, c)
    assert a is t and b is c
