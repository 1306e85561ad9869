"""Host-side multi-rank logic on CPU: world_size 2 over gloo (shard ranges, statistics gather, candidate arg-min +
broadcast).  The kernels themselves need no collective."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world_size, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from rlitg_b200 import dist as D
    g = torch.Generator(); g.manual_seed(0)
    cost = torch.rand(n_total, generator=g, dtype=torch.float64)
    reward = torch.randn(n_total, generator=g, dtype=torch.float64)
    done = (torch.rand(n_total, generator=g) < 0.3)
    rows = torch.arange(n_total * 4, dtype=torch.float64).reshape(n_total, 4)
    a, b = D.shard_range(n_total)
    st = D.gather_stats(D.local_stats(cost[a:b], reward[a:b], done[a:b], done[a:b], offset=a))
    best, gid, win = D.broadcast_best(rows[a:b], cost[a:b], a)
    q.put((rank, a, b, st["count"], st["sum_cost"], st["min_cost"], st["argmin"], st["sum_reward"], st["n_done"],
           best, gid, win.tolist()))
    dist.destroy_process_group()


def test_shard_range_partition():
    import sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    from rlitg_b200 import dist as D
    for n in (0, 1, 7, 8, 1024, 1_048_577):
        for w in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_stats_gather_world2_gloo():
    n_total, w = 1001, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, w, port, n_total, q)) for r in range(w)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(w)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator(); g.manual_seed(0)
    cost = torch.rand(n_total, generator=g, dtype=torch.float64)
    reward = torch.randn(n_total, generator=g, dtype=torch.float64)
    done = (torch.rand(n_total, generator=g) < 0.3)
    for r in res:
        rank, a, b, count, sum_cost, min_cost, argmin, sum_reward, n_done, best, gid, win = r
        assert count == n_total and abs(sum_cost - cost.sum().item()) < 1e-9
        assert argmin == int(cost.argmin()) and min_cost == cost.min().item()
        assert abs(sum_reward - reward.sum().item()) < 1e-9 and n_done == int(done.sum())
        assert gid == argmin and best == min_cost and win == [4.0 * gid + k for k in range(4)]
    assert sorted((r[1], r[2]) for r in res) == [(0, 501), (501, 1001)]
