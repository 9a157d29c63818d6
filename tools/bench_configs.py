"""BASELINE.json configs[2..4] through the PRODUCT entry point (inference.run_inference_sharded), strong-scaled over the ranks of
a torchrun job (one process per GPU, NCCL): a fixed job is split by pose ranges, one weight broadcast, one gather of the poses.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/bench_configs.py [--configs 3 4 5]

config 3: DiffDock-S architecture, 40 complexes (60 atoms / 300 residues) x 40 samples = 1600 poses, README temperatures
config 4: DisCo latent-conditioned score model (latent_dim 2, latent_vocab 1) + AR latent sampler + classifier-free guidance,
          10 complexes x 40 samples
config 5: large-receptor stress: 8 complexes of 120 atoms / 2000 residues x 40 samples, dynamic cross cut-off
Weights: seeded fresh initialisation (tests/helpers); the timed region is the whole call (start poses, H2D, 20 reverse steps, D2H,
gather).  Prints one JSON line per config on rank 0 (poses/s = poses of the whole job / max over ranks of the wall time)."""
import argparse
import json
import os
import sys
import time
from functools import partial

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from disco_diffdock_b200 import diffusion_utils as du, inference, latent as dlatent, synthetic  # noqa: E402
from tests import helpers  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--configs', type=int, nargs='*', default=[3, 4, 5])
ap.add_argument('--samples', type=int, default=40)
ap.add_argument('--repeat', type=int, default=2)
args = ap.parse_args()
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
real_stdout = os.dup(1)
os.dup2(2, 1)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)


def run(cfg_id):
    N = args.samples
    kw = dict(samples_per_complex=N, inference_steps=20, seed=5, no_final_step_noise=True, poses_per_call=400, **helpers.README_TEMPS)
    if cfg_id == 3:
        m, sd, cfg = helpers.make_model(0, gain=5.0)
        complexes = [synthetic.as_loader_item(synthetic.make_complex(3000 + c, 60, 300)) for c in range(40)]
        what = 'configs[2]: DiffDock-S, 40 complexes x %d samples, 60 atoms / 300 residues' % N
    elif cfg_id == 4:
        m, sd, cfg = helpers.make_model(3, latent_dim=2, latent_droprate=0.1, gain=5.0)
        complexes = [synthetic.as_loader_item(synthetic.make_complex(4000 + c, 60, 300)) for c in range(10)]
        what = 'configs[3]: DisCo score model + AR latent sampler + classifier-free guidance, 10 complexes x %d samples' % N
    else:
        m, sd, cfg = helpers.make_model(0, gain=5.0)
        complexes = [synthetic.as_loader_item(synthetic.make_complex(5000 + c, 120, 2000)) for c in range(8)]
        kw['poses_per_call'] = 80
        what = 'configs[4]: 8 complexes x %d samples, 120 atoms / 2000 residues, dynamic cross cut-off' % N
    m = m.to(dev)
    if cfg_id == 4:
        ar = dlatent.PretrainedScoreEncoder(pretrained_score_model=m, ns=cfg.ns, latent_dim=1, latent_vocab=1, latent_hidden_dim=128,
                                            input_latent_dim=cfg.latent_dim, apply_gumbel_softmax=True)
        ar.load_state_dict(helpers.make_ar_heads(21), strict=False)
        ar = ar.to(dev).eval()
        for g in complexes:
            g['ligand'].ar_pos = g['ligand'].pos.clone()
        kw.update(ar_model=ar, classifier_free_guidance_weight=0.6, cfg_start=1.0, cfg_end=0.3, softmax_latent_temperature=100.0)
    t2s = partial(du.t_to_sigma, args=cfg)
    best = None
    for r in range(args.repeat):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        res = inference.run_inference_sharded(complexes, m, cfg, dev, t2s, **kw)
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        best = float(dt.item()) if best is None else min(best, float(dt.item()))
    ok = all(bool(torch.isfinite(p).all()) for p in res['ligand_pos']) if rank == 0 or world == 1 else True
    n_poses = len(complexes) * N
    return {'config': what, 'n_gpus': world, 'poses': n_poses, 'seconds': best, 'poses_per_sec': n_poses / best, 'scaling': 'strong',
            'all_poses_gathered_and_finite': ok, 'entry_point': 'disco_diffdock_b200.inference.run_inference_sharded'}


lines = [run(c) for c in args.configs]
sys.stdout.flush()
os.dup2(real_stdout, 1)
if rank == 0:
    for l in lines:
        print(json.dumps(l), flush=True)
if world > 1:
    os.dup2(2, 1)
    dist.destroy_process_group()
