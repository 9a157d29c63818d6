#!/usr/bin/env python
"""Benchmark of the reverse-diffusion docking sampler (BASELINE.json metric: docked poses/s, 20 reverse steps,
40 samples per complex).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload dense|sparse]

Workload at every rank = BASELINE.json configs[1]: 10 synthetic complexes (60-atom ligand, 300 C-alpha receptor, 24-nn
receptor graph) x 40 samples x 20 reverse steps, DiffDock-S architecture (ns=24 nv=6 L=5 sh_lmax=1), seeded fresh-initialised
weights, README low-temperature coefficients.  One "step" = the whole 400-pose job.

  dense  (default, the headline): the trajectory a TRAINED model produces -- the ligand stays inside the protein, so every
          reverse step sees a PDBBind-shaped graph (E ~ 44 k edges per pose-step).  Random weights cannot hold a ligand in a
          pocket against the sigma_max = 19 A noise, so this is `evaluate.py --no_random`: start poses from
          randomize_position(no_random=True) (random torsions and orientation, centred on the protein, utils/sampling.py:12-46),
          z = 0 (utils/sampling.py:146-165), and the last linear layer of the three score heads scaled by 0.05 so that the drift
          term keeps the ligand where it is.  Every kernel does the same work per edge as with noise on.
  sparse : the round-1 workload -- randomize_position(sigma = 19 A) starts, noise on, head gain 5: the ligand drifts off and the
          late steps have no cross edge (E ~ 13 k per pose-step).  Reported next to the dense line (`sparse`).

Both arms (--impl ours / reference) build the SAME complexes, start poses and weights from the same seeds; the reference arm
(the CPU port of the reference algorithm, oracle/restate.py, on all host cores) runs complete 20-step trajectories of a bounded
sample of those poses, so edges per pose-step agree between the arms (`work.edges_per_pose_step` in both lines).

Weak scaling: with N ranks the job is 10*N complexes sharded by rank, no collective inside a diffusion step (weights broadcast
once, final poses gathered at the end of each step).

value : poses/s with every input resident in HBM (one ddk_sample call over the rank's 400 poses)
e2e   : poses/s through the drop-in sampling() API with HOST buffers (H2D of the static complex data, start poses, noise and
        step tables, the 20-step run, D2H of the final poses) -- the headline number: one call over the 10 x 40 graphs of the
        workload; the evaluate.py-style loop (one call per complex) is reported next to it
roofline / cpu_baseline: see DESIGN.md ("Measurement").
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LIG, N_REC, N_COMPLEX, N_SAMPLES, REV_STEPS = 60, 300, 10, 40, 20
W_CONV = (720, 936, 1152, 1872, 1872)
C_TP = (1008, 1368, 1728, 2736, 2736)
U_LEVEL = (96, 138, 180, 276, 276)                                                   # basis rows of layer l (level min(l, 3))
FLOP_PER_EDGE_REF = 2 * sum(72 * 72 + 72 * w + c for w, c in zip(W_CONV, C_TP))     # 1.014 MFLOP per edge-step
HEAD_GAIN = {'dense': 0.05, 'sparse': 5.0}


def workload_config(workload, n_complex):
    """The `config` object, identical in both arms."""
    start = ('randomize_position(no_random=True) starts, z = 0 (evaluate.py --no_random), score-head gain 0.05: ligand stays in the protein'
             if workload == 'dense' else 'randomize_position(sigma_tr_max = 19 A) starts, noise on, score-head gain 5: ligand drifts off')
    return {'workload': f'{n_complex} synthetic complexes ({N_LIG}-atom ligand, {N_REC} C-alpha receptor) x {N_SAMPLES} samples x '
                        f'{REV_STEPS} reverse steps, DiffDock-S ns=24 nv=6 L=5 sh_lmax=1 (BASELINE.json configs[1]), {workload} trajectory',
            'trajectory': workload, 'start_and_noise': start, 'weights': 'fresh-initialised, seed 0 (tests/helpers.make_model)',
            'temperatures': 'README.md:15 (DiffDock-S low-temperature sampling), --no_final_step_noise',
            'poses_per_gpu': n_complex * N_SAMPLES, 'reverse_steps': REV_STEPS,
            'l2': 'inputs larger than L2: the per-layer working set (edge embeddings, hidden units and harmonics of the listed '
                  'edges of 400 poses) is 2 - 6 GB per pass'}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json)', d
    return 6650.0, 'fallback (B200_PROFILING.md)', {}


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': float(self.rows[0][1]) if self.rows else None,
                'reasons': reasons, 'samples': len(self.rows)}


def build_workload(rank, n_complex, workload):
    """The rank's complexes with their start poses: [[40 loader items of complex c] ...] (evaluate.py:229-233)."""
    from disco_diffdock_b200 import synthetic
    from disco_diffdock_b200.sampling import randomize_position
    np.random.seed(1234 + rank)
    torch.manual_seed(1234 + rank)
    complexes = []
    for c in range(n_complex):
        g = synthetic.make_complex(1000 + rank * n_complex + c, N_LIG, N_REC)
        item = synthetic.as_loader_item(g)
        data_list = [copy.deepcopy(item) for _ in range(N_SAMPLES)]
        randomize_position(data_list, False, workload == 'dense', 19.0)        # evaluate.py:232-233 (--no_random for dense)
        complexes.append(data_list)
    return complexes


def make_weights(workload):
    from tests import helpers
    return helpers.make_model(0, gain=HEAD_GAIN[workload])


# ---------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference_run(workload, complexes, sd, cfg, picks, n_rev_steps, threads):
    """The oracle port (oracle/restate.py = the reference algorithm on torch-CPU) on the poses `picks` = [(complex, sample), ...]
    of the GPU arm's own workload: trajectories of `n_rev_steps` reverse steps from the same start poses with the same weights.
    Returns (poses/s scaled to 20-step poses, seconds, edges per pose-step)."""
    from disco_diffdock_b200 import data as ddata
    from oracle import restate
    from tests import helpers
    from tests.test_oracle_golden import load_tables
    torch.set_num_threads(threads)
    sched = np.linspace(1, 0, REV_STEPS + 1)[:-1]
    tables = load_tables()
    t0 = time.perf_counter()
    edges = []
    for ci, si in picks:                                   # one complex per call, like the reference's sampling()
        batch = ddata.Batch.from_data_list([copy.deepcopy(complexes[ci][si])])
        R = int(batch['ligand'].edge_mask.sum())
        if workload == 'dense':
            noise = {'tr': torch.zeros(n_rev_steps, 1, 3), 'rot': torch.zeros(n_rev_steps, 1, 3), 'tor': torch.zeros(n_rev_steps, R)}
        else:
            noise = helpers.draw_noise(3 + ci, n_rev_steps, 1, R)
        log = []
        with torch.no_grad():
            restate.sample(sd, cfg, batch, tables, sched, noise, inference_steps=n_rev_steps, edge_log=log, **helpers.README_TEMPS)
        edges += log
    dt = time.perf_counter() - t0
    pose_steps = len(picks) * n_rev_steps
    return pose_steps / REV_STEPS / dt, dt, float(np.mean(edges))


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    m, sd, cfg = make_weights(args.workload)
    complexes = build_workload(0, args.complexes, args.workload)
    for _ in range(args.warmup):
        cpu_reference_run(args.workload, complexes, sd, cfg, [(0, 0)], 1, threads)
    vals, times, edges = [], [], []
    for k in range(args.steps):                                  # bench step k: sample k of the first two complexes, all 20 steps
        v, dt, e = cpu_reference_run(args.workload, complexes, sd, cfg, [(0, k % N_SAMPLES), (1 % args.complexes, k % N_SAMPLES)],
                                     REV_STEPS, threads)
        vals.append(v); times.append(dt); edges.append(e)
    value = float(np.mean(vals))
    sample = (f'2 poses (sample k of complexes 0 and 1 of the same workload) x all {REV_STEPS} reverse steps per bench step, '
              f'same start poses and weights as the GPU arm')
    line = {'impl': 'reference', 'metric': 'docked_poses_per_sec', 'value': value, 'unit': 'poses/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(np.mean(times) * 1000), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic, fresh-init weights (seeded)',
            'config': workload_config(args.workload, args.complexes),
            'work': {'edges_per_pose_step': float(np.mean(edges)), 'poses_timed_per_step': 2, 'reverse_steps_timed': REV_STEPS},
            'cpu_baseline': {'value': value, 'unit': 'poses/s', 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'poses/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU arm
def time_resident(args, eng, m, cfg, dev, world, rank, complexes, workload, t2s, sched, dist):
    """`value`: K timed steps of one ddk_sample over the rank's poses, all inputs resident.  Returns a dict of measurements."""
    from disco_diffdock_b200 import data as ddata
    from disco_diffdock_b200 import sampling as dsampling
    from tests import helpers
    n_poses = len(complexes) * N_SAMPLES
    flat = [g for c in complexes for g in c]
    big = ddata.Batch.from_data_list(flat)
    info = eng.set_batch(big)
    steps_tab = dsampling.build_step_tables(m, cfg, t2s, sched, sched, sched, REV_STEPS, n_poses, **helpers.README_TEMPS)
    if workload == 'dense':
        z = None                                              # --no_random: utils/sampling.py:146-165 draws zeros
    else:
        gen = torch.Generator(device=dev).manual_seed(77 + rank)
        z = {'tr': torch.randn(REV_STEPS, n_poses, 3, device=dev, generator=gen),
             'rot': torch.randn(REV_STEPS, n_poses, 3, device=dev, generator=gen),
             'tor': torch.randn(REV_STEPS, info.RB, device=dev, generator=gen)}
        for v in z.values():
            v[-1] = 0                                         # --no_final_step_noise (README.md:15)
    for k in ('semb', 'cutoff', 'tr_sigma', 'rot_scale', 'tor_scale'):
        setattr(steps_tab, k, getattr(steps_tab, k).to(dev).contiguous())
    pos0 = big['ligand'].pos.to(dev).contiguous()
    gathered = [torch.empty_like(pos0) for _ in range(world)] if world > 1 else None

    def resident_step():
        pos = pos0.clone()
        eng.sample(pos, steps_tab, z)
        if world > 1:
            dist.all_gather(gathered, pos)                    # gather of the final poses
        return pos

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        resident_step()
    barrier()
    eng.profile_read()
    launches0, edges0 = eng.kernel_launches(), eng.edge_total()
    ge0, gs0 = eng.group_totals()
    sampler = ClockSampler(dev.index)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        final = resident_step()
    e1.record()
    barrier()
    sampler.stop_flag = True
    ms_local = e0.elapsed_time(e1)
    prof = eng.profile_read()
    ge1, gs1 = eng.group_totals()
    edges = eng.edge_total() - edges0 + (info.EB + info.ER) * REV_STEPS * args.steps
    assert torch.isfinite(final).all()
    return dict(ms_local=ms_local, prof=prof, launches=eng.kernel_launches() - launches0, edges=edges,
                g_edges=(ge1 - ge0).astype(float), g_segs=(gs1 - gs0).astype(float), info=info, clocks=sampler.summary(),
                n_poses=n_poses, barrier=barrier)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--workload', default='dense', choices=['dense', 'sparse'])
    ap.add_argument('--complexes', type=int, default=N_COMPLEX)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-sparse', action='store_true', help='skip the second (other trajectory) line')
    ap.add_argument('--no-strict', action='store_true', help='skip the strict_fp32 line (the same job with DDK_TC=0)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from disco_diffdock_b200 import diffusion_utils as du
    from disco_diffdock_b200 import sampling as dsampling
    from disco_diffdock_b200 import build as ddk_build
    from tests import helpers
    ddk_build.build()
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    # anything libraries write to stdout (e.g. NCCL's version banner) goes to stderr: stdout carries exactly one JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def load_model(workload):
        # weights: rank 0 initialises, one NCCL broadcast replicates them (no other parameter traffic, ever)
        m, sd, cfg = make_weights(workload)
        m = m.to(dev)
        if world > 1:
            for p_ in list(m.parameters()) + list(m.buffers()):
                dist.broadcast(p_.data, src=0)
            m.invalidate()
        eng = m.engine(dev)
        eng.profile_enable(True)
        return m, sd, cfg, eng

    sched = du.get_t_schedule(REV_STEPS)
    n_complex = args.complexes
    m, sd, cfg, eng = load_model(args.workload)
    t2s = partial(du.t_to_sigma, args=cfg)
    complexes = build_workload(rank, n_complex, args.workload)
    R = time_resident(args, eng, m, cfg, dev, world, rank, complexes, args.workload, t2s, sched, dist)
    barrier, n_poses, info, prof = R['barrier'], R['n_poses'], R['info'], R['prof']

    def reduce_max(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_all(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world == 1:
            return [float(x)]
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    ms = reduce_max(R['ms_local'])
    value = n_poses * world * args.steps / (ms / 1000)
    edges_per_pose_step = R['edges'] / (n_poses * REV_STEPS * args.steps)
    per_rank = {'ms_per_step': [v / args.steps for v in gather_all(R['ms_local'])],
                'edges_per_pose_step': gather_all(edges_per_pose_step)}

    # ---------------------------------------------------------------- e2e: host buffers through sampling()
    # Two ways of driving the drop-in API, both with every buffer in host memory:
    #   batched      : ONE sampling() call over the 10 x 40 graphs of the workload ("batch 10 complexes", configs[1]); the
    #                  copies of each complex are recognised, shipped once and replicated on the device -- the e2e value
    #   per complex  : one sampling() call per complex (40 poses), the loop of evaluate.py:219-291
    dense = args.workload == 'dense'

    def e2e_batched(seed):
        g = torch.Generator().manual_seed(seed)
        dl = [x.shallow_copy() for data_list in complexes for x in data_list]     # sampling() rebinds ['ligand'].pos only
        out, _ = dsampling.sampling(dl, m, REV_STEPS, sched, sched, sched, dev, t2s, cfg, no_random=dense, batch_size=len(dl),
                                    no_final_step_noise=True, generator=g, host_buffers=True, **helpers.README_TEMPS)
        return sum(x['ligand'].pos.numel() * 4 for x in out)

    def e2e_per_complex(seed):
        out_bytes = 0
        g = torch.Generator().manual_seed(seed)
        for ci, data_list in enumerate(complexes):
            dl = [x.shallow_copy() for x in data_list]
            out, _ = dsampling.sampling(dl, m, REV_STEPS, sched, sched, sched, dev, t2s, cfg, no_random=dense, batch_size=N_SAMPLES,
                                        no_final_step_noise=True, generator=g, host_buffers=True, **helpers.README_TEMPS)
            out_bytes += sum(x['ligand'].pos.numel() * 4 for x in out)
        return out_bytes

    def time_e2e(fn):
        for w in range(max(1, args.warmup - 1)):
            fn(w)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            nbytes = fn(100 + k)
        barrier()
        return n_poses * world * args.steps / reduce_max(time.perf_counter() - t0), nbytes

    e2e_pc_value, _ = time_e2e(e2e_per_complex)
    e2e_value, d2h = time_e2e(e2e_batched)
    bi = eng.batch_info                                       # of the batched call: every complex shipped once
    h2d = bi.h2d_bytes + bi.NL * 3 * 4 + REV_STEPS * bi.B * (32 + 4) * 4 + (0 if dense else REV_STEPS * (2 * bi.B * 3 + bi.RB) * 4)

    # ---------------------------------------------------------------- roofline of the dominant kernels
    # The two 84-wide conv layers (basis level 3): one launch = one layer over every pose of the rank.  Work model (DESIGN.md
    # section 5): per listed edge 2*72*U FLOP of outer-product accumulation (U = 276), per non-empty segment 2*72*2736 FLOP of
    # contraction with the second radial-MLP layer (W = 1872 weight rows, vector classes used for 3 components); algorithmic
    # bytes per launch = node features in and out, list entry, harmonics and the 72 hidden units of every listed edge.
    hbm_peak, peak_src, peaks_raw = peaks()
    lv3_ms = prof['conv_accum_lv3'][0] + prof['conv_tc_lv3'][0]
    lv3_n = max(prof['conv_accum_lv3'][1], prof['conv_tc_lv3'][1], 1)
    total_prof_ms = sum(v[0] for v in prof.values())
    passes = REV_STEPS * args.steps                                      # reverse steps in the timed region
    g_edges, g_segs = R['g_edges'], R['g_segs']
    # the two level-3 launches of a step: layer 3 over the segments of ligand nodes and of the residues with a cross edge
    # (work lists 0, 1, 3 and 4 = receptor contacts of those residues), layer 4 (the last before the heads) over the segments
    # of ligand nodes only (lists 0, 1) -- the heads never read receptor features
    e_launch = (g_edges[[0, 1, 3, 4]].sum() + g_edges[:2].sum()) / 2 / passes   # listed edges per launch, average of the two
    s_launch = (g_segs[[0, 1, 3, 4]].sum() + g_segs[:2].sum()) / 2 / passes
    n_nodes = ((info.NL + info.NR) + info.NL) / 2
    bytes_launch = n_nodes * (84 + 84) * 4 + e_launch * (8 + 16 + 72 * 4) + s_launch * 16
    flop_launch = e_launch * 2 * 72 * U_LEVEL[3] + s_launch * 2 * 72 * C_TP[3]
    t_launch = lv3_ms / lv3_n / 1000
    clocks = R['clocks']
    sm_mhz = clocks.get('sm_mhz') or peaks_raw.get('sm_max_mhz', 1965.0)
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    tf32_peak = peaks_raw.get('bf16_tflops_sustained', 1373.3) / 2           # dense TF32 = half the measured bf16 rate
    conv_keys = [k for k in prof if k.startswith('conv_')]
    conv_ms = sum(prof[k][0] for k in conv_keys)
    tc_ms = sum(prof[k][0] for k in conv_keys if k.startswith('conv_tc'))
    tensor_bound = tc_ms > 0.5 * conv_ms                                      # which pipe carries the layer
    traffic, traffic_src = None, None
    tj = os.path.join(ROOT, 'profiles', 'conv_lv3_traffic.json')
    if os.path.exists(tj) and n_complex == N_COMPLEX:
        td = json.load(open(tj))
        if td.get('workload') == args.workload:
            traffic, traffic_src = td['dram_bytes_per_launch'], 'profiles/conv_lv3_traffic.json (' + td['source'] + ')'
    peak_tf = tf32_peak if tensor_bound else fp32_peak
    roofline = {'kernel': 'level-3 conv layer: ' + ' + '.join(k for k in ('conv_accum_lv3', 'conv_tc_lv3') if prof[k][0] > 0),
                'bound': 'tensor' if tensor_bound else 'fp32_fma',
                'achieved': flop_launch / t_launch / 1e12, 'peak': peak_tf, 'unit': 'TFLOP/s',
                'frac': flop_launch / t_launch / 1e12 / peak_tf,
                'peak_source': ('dense TF32 = MEASURED_PEAKS.json bf16_tflops_sustained / 2; a 3xTF32 product spends three tensor '
                                'passes per algorithmic FLOP, so frac <= 1/3 by construction') if tensor_bound else
                               '148 SMs x 128 FMA/clk x 2 at the SM clock sampled in this run',
                'traffic': traffic, 'traffic_unit': 'bytes per launch', 'traffic_source': traffic_src,
                'algorithmic_bytes_per_launch': bytes_launch, 'algorithmic_flop_per_launch': flop_launch,
                'hbm': {'achieved_gbs': bytes_launch / t_launch / 1e9, 'peak_gbs': hbm_peak, 'frac': bytes_launch / t_launch / 1e9 / hbm_peak,
                        'peak_source': peak_src, 'note': 'compute-bound layer (SURVEY 8d): a low HBM fraction is the design goal'},
                'fp32_fma': {'achieved_tflops': flop_launch / t_launch / 1e12, 'peak_tflops_at_measured_clock': fp32_peak,
                             'frac': flop_launch / t_launch / 1e12 / fp32_peak},
                'launch_ms': t_launch * 1000, 'launches': lv3_n, 'share_of_step': lv3_ms / max(total_prof_ms, 1e-9),
                'conv_share_of_step': conv_ms / max(total_prof_ms, 1e-9),
                'tensor_core_share_of_conv': tc_ms / max(conv_ms, 1e-9),
                'edges_per_launch': e_launch, 'segments_per_launch': s_launch,
                'kernel_ms': {k: round(v[0], 3) for k, v in prof.items()}}

    line = {'metric': 'docked_poses_per_sec', 'value': value, 'unit': 'poses/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic, fresh-init weights (seeded)',
            'arithmetic': {'2': 'fp32 in / out; the outer-product accumulations of the conv layers as 3xTF32 split products (hi*hi + hi*lo '
                                '+ lo*hi, fp32 accumulators in tensor memory, chains of at most 64 edges) on tcgen05, everything else fp32 '
                                'FMA (k_conv_tcr)',
                           '1': 'fp32 FMA (k_conv_fused); the long lig<-rec segments as 3xTF32 split products on tcgen05 (k_acc_tc)',
                           '0': 'fp32 FMA throughout (k_conv_fused)'}[os.environ.get('DDK_TC', '2') if os.environ.get('DDK_TC', '2') in '012' else '2'],
            'config': workload_config(args.workload, n_complex),
            'work': {'edges_per_pose_step': edges_per_pose_step,
                     'reference_formulation_equiv_tflops': R['edges'] * world * FLOP_PER_EDGE_REF / (ms / 1000) / 1e12,
                     'parallelism': f'pose-sharded x{world}', 'per_rank': per_rank},
            'e2e': {'value': e2e_value, 'unit': 'poses/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'call': 'one sampling() call over the 10 x 40 graphs, host buffers',
                    'per_complex_calls': {'value': e2e_pc_value, 'unit': 'poses/s',
                                          'call': 'one sampling() call per complex (40 poses), the evaluate.py loop'}},
            'gpu_launches': int(R['launches']), 'clocks': clocks, 'roofline': roofline}

    # ---------------------------------------------------------------- the other trajectory next to it (resident only)
    if not args.no_sparse:
        other = 'sparse' if dense else 'dense'
        m2, sd2, cfg2, eng2 = load_model(other)
        cx2 = build_workload(rank, n_complex, other)
        a2 = argparse.Namespace(steps=max(1, args.steps - 1), warmup=max(1, args.warmup - 2))
        R2 = time_resident(a2, eng2, m2, cfg2, dev, world, rank, cx2, other, partial(du.t_to_sigma, args=cfg2), sched, dist)
        ms2 = reduce_max(R2['ms_local'])
        line[other] = {'value': n_poses * world * a2.steps / (ms2 / 1000), 'unit': 'poses/s', 'steps': a2.steps,
                       'edges_per_pose_step': R2['edges'] / (n_poses * REV_STEPS * a2.steps),
                       'trajectory': workload_config(other, n_complex)['start_and_noise']}
    if not args.no_strict and os.environ.get('DDK_TC') is None:
        # the same job on the all-fp32-FMA conv kernels (DDK_TC=0): the path that holds 1e-3 A on ill-conditioned trajectories too
        from disco_diffdock_b200 import engine as dengine
        dengine.set_tensor_core_path(0)
        a0 = argparse.Namespace(steps=max(1, args.steps - 1), warmup=1)
        R0 = time_resident(a0, eng, m, cfg, dev, world, rank, complexes, args.workload, t2s, sched, dist)
        dengine.set_tensor_core_path(None)
        ms0 = reduce_max(R0['ms_local'])
        line['strict_fp32'] = {'value': n_poses * world * a0.steps / (ms0 / 1000), 'unit': 'poses/s', 'steps': a0.steps, 'conv_path': 'DDK_TC=0',
                               'note': 'conv layers on k_conv_fused (packed fp32 FMA) only; the headline runs them as 3xTF32 split '
                                       'products on tcgen05 (k_conv_tcr); both hold the 1e-3 A bound in the pretrained regime (DESIGN.md section 2)'}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_reference_run(args.workload, complexes, sd, cfg, [(0, 0)], 1, threads)
        v, dt, e = cpu_reference_run(args.workload, complexes, sd, cfg, [(0, 0), (1 % n_complex, 0)], REV_STEPS, threads)
        line['cpu_baseline'] = {'value': v, 'unit': 'poses/s', 'cores': threads, 'kind': 'port', 'edges_per_pose_step': e,
                                'sample': f'2 poses of this workload (sample 0 of complexes 0 and 1) x all {REV_STEPS} reverse steps ({dt:.1f} s)'}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
