"""CPU, world_size 2, gloo: the host-side sharding / gather logic of monoloco_b200.distributed (the kernel itself
has no CPU form, so each rank fabricates its shard's output rows deterministically)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_rows(start, stop, out_size):
    idx = torch.arange(start, stop, dtype=torch.float32)
    raw = idx[:, None] * 10 + torch.arange(out_size, dtype=torch.float32)[None]
    dec = idx[:, None] * 100 + torch.arange(8, dtype=torch.float32)[None]
    return raw, dec


def _worker(rank, world, port, n_total, out_size, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from monoloco_b200 import distributed as D
    start, stop = D.shard_range(n_total, world, rank)
    raw, dec = _fake_rows(start, stop, out_size)
    # gloo has no all_gather_into_tensor for uneven shards either: exercise the padded path through all_gather
    rows = D.pack_rows(raw, dec)
    sizes = D.shard_sizes(n_total, world)
    mx = max(sizes)
    padded = torch.zeros((mx, rows.shape[1]))
    padded[:rows.shape[0]] = rows
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    full = torch.cat([parts[r][:sizes[r]] for r in range(world)], dim=0)
    r2, d2 = D.unpack_rows(full, out_size)
    eraw, edec = _fake_rows(0, n_total, out_size)
    q.put((rank, bool(torch.equal(r2, eraw) and torch.equal(d2, edec)), (start, stop)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_total,out_size', [(10, 9), (7, 10), (1, 9)])
def test_sharded_gather_gloo(n_total, out_size):
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, out_size, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    ranges = sorted(r for _, _, r in res)
    assert ranges[0][0] == 0 and ranges[-1][1] == n_total and ranges[0][1] == ranges[1][0]


def test_shard_range_properties():
    from monoloco_b200 import distributed as D
    for n in (0, 1, 5, 4096, 1048576, 1000003):
        for w in (1, 2, 4, 8):
            rs = [D.shard_range(n, w, r) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1 and sizes == D.shard_sizes(n, w)
