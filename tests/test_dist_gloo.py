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


# ---- the pose-sharded product entry point (inference.run_inference_sharded) with a stand-in sampler: the plumbing (shards, per-pose
# noise and start poses, one gather) must give every rank the unsharded result bit for bit
def _fake_sampler(data_list, noise=None, **kw):
    off = 0
    for i, g in enumerate(data_list):
        R = int(g['ligand'].edge_mask.sum())
        # sums in float64 over contiguous copies: the stand-in itself must not depend on the strides of the batch it sits in
        z = (noise['tr'][:, i].contiguous().double().sum(0) + noise['rot'][:, i].contiguous().double().sum()
             + noise['tor'][:, off:off + R].contiguous().double().sum())
        g['ligand'].pos = (g['ligand'].pos.double() * 1.25 + z).float()
        off += R
    return data_list, None


def _sharded_inputs():
    from types import SimpleNamespace
    from disco_diffdock_b200 import synthetic
    gs = [synthetic.as_loader_item(synthetic.make_complex(60 + i, 8 + 3 * i, 12 + i)) for i in range(3)]
    return gs, SimpleNamespace(tr_sigma_max=19.0, no_torsion=False)


def _sharded_worker(rank, world, port, q):
    from disco_diffdock_b200 import inference
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    gs, args = _sharded_inputs()
    torch.manual_seed(100 + rank)                          # the global generators differ per rank: nothing may depend on them
    res = inference.run_inference_sharded(gs, None, args, 'cpu', None, samples_per_complex=5, inference_steps=4, seed=3,
                                          sampler=_fake_sampler, poses_per_call=4, broadcast_weights=False)
    q.put((rank, res['ligand_pos'], res['shard']))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_inference_two_ranks_equals_unsharded():
    from disco_diffdock_b200 import inference
    gs, args = _sharded_inputs()
    full = inference.run_inference_sharded(gs, None, args, 'cpu', None, samples_per_complex=5, inference_steps=4, seed=3, rank=0,
                                           world=1, sampler=_fake_sampler, poses_per_call=7)
    assert not any(torch.isnan(p).any() for p in full['ligand_pos'])
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=150) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    units = sorted((c, k) for r in res for c, a, b in r[2] for k in range(a, b))
    assert units == [(c, k) for c in range(3) for k in range(5)]              # the shards partition the poses
    for r in res:
        for ci in range(3):
            assert torch.equal(r[1][ci], full['ligand_pos'][ci])              # every rank holds the unsharded result
