"""Where the end-to-end time of bench.py goes (sampling() with host buffers, one complex = 40 poses per call):
wall time of the set_batch / sample_host halves, CUDA-event time per kernel class at 40 poses per launch, and the same
kernel classes at 400 poses per launch (the resident configuration) for comparison.  Optional: --cprofile."""
import argparse, cProfile, os, pstats, sys, time
from functools import partial
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from disco_diffdock_b200 import data as ddata, diffusion_utils as du, sampling as dsampling
from tests import helpers

ap = argparse.ArgumentParser()
ap.add_argument('--complexes', type=int, default=10)
ap.add_argument('--cprofile', action='store_true')
ap.add_argument('--workload', choices=['dense', 'sparse'], default='dense')
args = ap.parse_args()
dev = torch.device('cuda')
m, sd, cfg = bench.make_weights(args.workload)
dense = args.workload == 'dense'
m = m.to(dev); eng = m.engine(dev)
t2s = partial(du.t_to_sigma, args=cfg)
sched = du.get_t_schedule(bench.REV_STEPS)
complexes = bench.build_workload(0, args.complexes, args.workload)
n_poses = args.complexes * bench.N_SAMPLES


def step(seed):
    g = torch.Generator().manual_seed(seed)
    for data_list in complexes:
        dl = [x.shallow_copy() for x in data_list]
        dsampling.sampling(dl, m, bench.REV_STEPS, sched, sched, sched, dev, t2s, cfg, no_random=dense, batch_size=bench.N_SAMPLES,
                           no_final_step_noise=True, generator=g, host_buffers=True, **helpers.README_TEMPS)


def timed(fn, *a):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(*a); torch.cuda.synchronize()
    return time.perf_counter() - t0


step(0); step(1)
print(f'e2e (no profiling events): {n_poses / timed(step, 2):.1f} poses/s')

# ---- halves of one call, wall clock
tab = dsampling.build_step_tables(m, cfg, t2s, sched, sched, sched, bench.REV_STEPS, bench.N_SAMPLES, **helpers.README_TEMPS)
ts, tr = [], []
for data_list in complexes:
    ts.append(timed(eng.set_batch_copies, data_list[0], bench.N_SAMPLES))
    R = eng.batch_info.RB
    z = None if dense else {'tr': torch.randn(bench.REV_STEPS, bench.N_SAMPLES, 3), 'rot': torch.randn(bench.REV_STEPS, bench.N_SAMPLES, 3),
                            'tor': torch.randn(bench.REV_STEPS, R)}
    pos = torch.cat([x['ligand'].pos for x in data_list], 0).float().contiguous()
    tr.append(timed(eng.sample_host, pos, tab, z))
print(f'per complex: set_batch_copies {1e3 * np.mean(ts):.2f} ms, sample_host {1e3 * np.mean(tr):.2f} ms')

# ---- kernel classes at 40 poses per launch
eng.profile_enable(True)
eng.profile_read()
w = timed(step, 3)
p40 = eng.profile_read()
print(f'e2e with profiling events: {n_poses / w:.1f} poses/s; kernel ms per complex (40 poses per launch):')
tot40 = sum(v[0] for v in p40.values())
for k, v in p40.items():
    print(f'  {k:16s} {v[0] / args.complexes:8.3f} ms  {v[1] // args.complexes:5d} launches')
print(f'  total            {tot40 / args.complexes:8.3f} ms of {1e3 * w / args.complexes:.2f} ms wall per complex')

# ---- the same classes at 400 poses per launch
flat = [g for c in complexes for g in c]
big = ddata.Batch.from_data_list(flat)
info = eng.set_batch(big)
tabN = dsampling.build_step_tables(m, cfg, t2s, sched, sched, sched, bench.REV_STEPS, n_poses, **helpers.README_TEMPS)
z = None if dense else {'tr': torch.randn(bench.REV_STEPS, n_poses, 3, device=dev), 'rot': torch.randn(bench.REV_STEPS, n_poses, 3, device=dev),
                        'tor': torch.randn(bench.REV_STEPS, info.RB, device=dev)}
pos0 = big['ligand'].pos.to(dev).contiguous()
eng.sample(pos0.clone(), tabN, z)
eng.profile_read()
w = timed(lambda: eng.sample(pos0.clone(), tabN, z))
pN = eng.profile_read()
print(f'resident: {n_poses / w:.1f} poses/s; kernel ms per complex ({n_poses} poses per launch):')
for k, v in pN.items():
    print(f'  {k:16s} {v[0] / args.complexes:8.3f} ms  (x{p40[k][0] / max(v[0], 1e-9):.2f} at 40 poses per launch)')
print(f'  total            {sum(v[0] for v in pN.values()) / args.complexes:8.3f} ms')

# ---- the batched call (one sampling() over all complexes, host buffers): halves
from disco_diffdock_b200.sampling import group_copies
from disco_diffdock_b200.engine import group_index_arrays
dl = [x.shallow_copy() for c in complexes for x in c]
t0 = time.perf_counter(); groups = group_copies(dl); t1 = time.perf_counter(); group_index_arrays(groups, False); t2 = time.perf_counter()
print(f'batched call: group_copies {1e3 * (t1 - t0):.1f} ms, group_index_arrays {1e3 * (t2 - t1):.1f} ms (host only)')
for _ in range(2):
    tsb = timed(eng.set_batch_groups, groups)
zh = None if z is None else {k: v.cpu() for k, v in z.items()}
posh = torch.cat([x['ligand'].pos for x in dl], 0).float().contiguous()
for _ in range(2):
    tsh = timed(eng.sample_host, posh.clone(), tabN, zh)
print(f'batched call: set_batch_groups {1e3 * tsb:.1f} ms, sample_host {1e3 * tsh:.1f} ms, resident sample {1e3 * w:.1f} ms')

def batched(seed):
    g = torch.Generator().manual_seed(seed)
    d = [x.shallow_copy() for c in complexes for x in c]
    dsampling.sampling(d, m, bench.REV_STEPS, sched, sched, sched, dev, t2s, cfg, no_random=dense, batch_size=len(d), no_final_step_noise=True,
                       generator=g, host_buffers=True, **helpers.README_TEMPS)
eng.profile_enable(False)
batched(0)
print(f'batched e2e: {n_poses / timed(batched, 1):.1f} poses/s')

if args.cprofile:
    eng.profile_enable(False)
    pr = cProfile.Profile(); pr.enable(); batched(4); torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
