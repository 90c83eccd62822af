"""world_size-2 gloo tests (CPU) of the crop-stream sharding and the pose all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from egonet_b200 import sharding


def test_partition_covers_stream_exactly():
    for n in (0, 1, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 8):
            blocks = [sharding.partition(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
    assert sharding.partition(4096, 3, 8) == (1536, 2048)      # BASELINE configs[4]: 512 crops per rank
    with pytest.raises(ValueError):
        sharding.partition(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = sharding.partition(n_total, rank, world)
        idx = torch.arange(lo, hi, dtype=torch.float64)
        # stand-in for the per-crop pose record: any pure function of the GLOBAL crop index
        local = torch.stack([idx * 0.5 + k for k in range(7)], dim=1)
        full = sharding.all_gather_records(local, n_total, dist)
        ref_idx = torch.arange(n_total, dtype=torch.float64)
        ref = torch.stack([ref_idx * 0.5 + k for k in range(7)], dim=1)
        ok = full.shape == ref.shape and bool(torch.equal(full, ref))
        # a rank holding the wrong block must be rejected, not silently gathered
        try:
            sharding.all_gather_records(local[:-1] if len(local) else local.new_zeros((1, 7)), n_total, dist)
            bad = False
        except ValueError:
            bad = True
        q.put((rank, ok, bad))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_total', [64, 65, 3])
def test_two_rank_gather_gloo(n_total):
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = sorted(q.get() for _ in range(2))
    assert results == [(0, True, True), (1, True, True)]


def test_single_process_gather_is_identity():
    x = torch.arange(14, dtype=torch.float64).view(2, 7)
    assert sharding.all_gather_records(x, 2) is x
    with pytest.raises(ValueError):
        sharding.all_gather_records(x, 3)


def _grad_worker(rank, world, port, q):
    """Data-parallel gradient averaging of libs/trainer/trainer.py over gloo: flat-buffer path (one collective)
    and the per-parameter fallback."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from egonet_b200.libs.trainer.trainer import allreduce_gradients

        class Flat(torch.nn.Module):                       # stands in for a natively trained module
            def __init__(self):
                super().__init__()
                self.a = torch.nn.Parameter(torch.zeros(3))
                self.b = torch.nn.Parameter(torch.zeros(2, 2))
                flat = torch.arange(8, dtype=torch.float32) * (rank + 1)
                self._train = {'last_grads': flat}
                self.a.grad, self.b.grad = flat[:3], flat[4:8].view(2, 2)      # views, as the engine hands them out

        m = Flat()
        n1 = allreduce_gradients(m)
        mean_scale = sum(r + 1 for r in range(world)) / world
        ok = n1 == 1 and torch.allclose(m.a.grad, torch.arange(3.) * mean_scale) and \
            torch.allclose(m.b.grad, (torch.arange(4.) + 4).view(2, 2) * mean_scale)
        plain = torch.nn.Linear(2, 2)
        for p in plain.parameters():
            p.grad = torch.full_like(p, float(rank))
        n2 = allreduce_gradients(plain)
        ok = ok and n2 == 2 and all(torch.allclose(p.grad, torch.full_like(p, (world - 1) / 2)) for p in plain.parameters())
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]
