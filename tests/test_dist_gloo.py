"""Multi-process host logic of the pose-sharded sampler on the gloo backend (world_size 2, CPU):
shards are disjoint and cover every pose, the weight broadcast replicates rank 0, the gathered result equals the
unsharded one bit for bit (per-pose noise streams are keyed by pose id)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from disco_diffdock_b200 import dist as ddist


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 400, 1601):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a, b = ddist.shard_range(n, r, world)
                seen += list(range(a, b))
                assert b - a in (n // world, n // world + 1)
            assert seen == list(range(n))


def test_shard_poses_balanced_by_cost():
    samples = [40, 40, 40, 8]
    cost = [60 * 300, 120 * 2000, 30 * 100, 60 * 300]
    for world in (1, 2, 4, 8):
        units = []
        loads = []
        for r in range(world):
            sh = ddist.shard_poses(samples, cost, r, world)
            loads.append(sum((b - a) * cost[c] for c, a, b in sh))
            units += [(c, k) for c, a, b in sh for k in range(a, b)]
        assert sorted(units) == [(c, k) for c, s in enumerate(samples) for k in range(s)]
        assert max(loads) - min(loads) <= 2 * max(cost)


def fake_sample(pose_id, n_atoms=5):
    g = torch.Generator().manual_seed(1000 + pose_id)      # noise stream keyed by pose id, not by rank
    return torch.randn(n_atoms, 3, generator=g)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(rank)                                # ranks start with different weights
    lin = torch.nn.Linear(4, 3)
    ddist.broadcast_module(lin, src=0)
    n = 11
    a, b = ddist.shard_range(n, rank, world)
    local = torch.stack([fake_sample(i) for i in range(a, b)]) if b > a else torch.zeros(0, 5, 3)
    counts = [ddist.shard_range(n, r, world)[1] - ddist.shard_range(n, r, world)[0] for r in range(world)]
    full = ddist.gather_poses(local, counts)
    q.put((rank, lin.weight.detach().clone(), full))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_roundtrip():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=90) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert torch.equal(res[0][1], res[1][1])                                  # weights replicated
    want = torch.stack([fake_sample(i) for i in range(11)])
    assert torch.equal(res[0][2], want) and torch.equal(res[1][2], want)      # sharded == unsharded, bit for bit
