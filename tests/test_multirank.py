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
