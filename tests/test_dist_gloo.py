"""world_size-2 gloo tests (CPU): ray sharding, the flat gradient bucket and the row gather of a sharded render."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cfnerf_b200 import dist as D
        torch.manual_seed(0)
        rays = torch.arange(11 * 37, dtype=torch.float32).reshape(37, 11)      # ragged: 37 rays over 2 ranks
        mine = D.shard_rays(rays)
        lo, hi = D.shard_bounds(37, rank, world)
        assert torch.equal(mine, rays[lo:hi])
        # a per-ray "render": the gathered rows must equal the unsharded result exactly
        full = rays.sin().sum(-1, keepdim=True).repeat(1, 3)
        got = D.gather_rows(mine.sin().sum(-1, keepdim=True).repeat(1, 3), 37)
        assert torch.equal(got, full)
        # gradient bucket: mean of per-rank gradients, dead parameters untouched
        p1 = torch.nn.Parameter(torch.zeros(5, 3))
        p2 = torch.nn.Parameter(torch.zeros(7))
        dead = torch.nn.Parameter(torch.zeros(2))
        p1.grad = torch.full((5, 3), float(rank + 1))
        p2.grad = torch.arange(7.0) * (rank + 1)
        b = D.GradBucket([p1, p2, dead])
        b.all_reduce_mean_()
        assert torch.allclose(p1.grad, torch.full((5, 3), 1.5))
        assert torch.allclose(p2.grad, torch.arange(7.0) * 1.5)
        assert dead.grad is None
        # equal shards: mean of per-rank mean losses == global mean
        x = torch.arange(64.0)
        local = D.shard_rays(x).mean()
        t = local.clone()
        dist.all_reduce(t)
        assert abs(float(t) / world - float(x.mean())) < 1e-6
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharding_and_gradient_bucket_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_bounds_cover_everything():
    from cfnerf_b200.dist import shard_bounds
    for n in (0, 1, 7, 64, 262144):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
