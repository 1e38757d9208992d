"""CPU tests of the multi-GPU host logic (world_size 2, gloo): clip sharding and the single all-gather of metric rows."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spatialaudiogen_b200 import dist as D


def test_schedule_shape_and_sharding_cover_everything_once():
    L = D.yt_all_clip_lengths()
    assert len(L) == 285 and L.min() >= 20 and L.max() <= 1800
    assert abs(L.mean() - 355) < 60                      # 113 h / 1146 videos ~ 355 s
    sched = D.eval_schedule(L, 32)
    assert all(n == 32 for _, _, n in sched)
    assert len(sched) == sum(int(l) // 32 for l in L)
    for world in (1, 2, 8):
        parts = [D.shard(sched, r, world) for r in range(world)]
        assert sum(len(p) for p in parts) == len(sched)
        assert sorted(sum(parts, [])) == sorted(sched)
        for r, p in enumerate(parts):
            assert all(c % world == r for c, _, _ in p)  # whole clips (hence whole batches) stay on one rank


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        n = 3 + 2 * rank                                 # ragged counts: 3 and 5 rows
        rows = torch.arange(n * 4, dtype=torch.float32).reshape(n, 4) + 100 * rank
        ids = torch.stack([torch.full((n,), rank), torch.arange(n)], 1)
        for max_rows in (None, 5):
            r, i = D.gather_rows(rows, ids, max_rows=max_rows)
            q.put((rank, max_rows, r.numpy(), i.numpy()))
        with pytest.raises(ValueError):
            D.gather_rows(rows, ids, max_rows=2)
    finally:
        dist.destroy_process_group()


def test_gather_rows_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2 * world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp_rows = np.concatenate([np.arange(12, dtype=np.float32).reshape(3, 4), np.arange(20, dtype=np.float32).reshape(5, 4) + 100])
    exp_ids = np.concatenate([np.stack([np.zeros(3), np.arange(3)], 1), np.stack([np.ones(5), np.arange(5)], 1)]).astype(np.int64)
    for rank, max_rows, r, i in got:
        assert np.array_equal(r, exp_rows), (rank, max_rows)
        assert np.array_equal(i, exp_ids), (rank, max_rows)


def test_gather_rows_single_process_is_identity():
    rows = torch.randn(4, 3)
    ids = torch.tensor([[0, 1], [0, 2], [1, 0], [1, 1]])
    r, i = D.gather_rows(rows, ids)
    assert torch.equal(r, rows) and torch.equal(i, ids)
