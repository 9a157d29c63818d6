#!/usr/bin/env python
"""Benchmark of the reverse-diffusion docking sampler (BASELINE.json metric: docked poses/s, 20 reverse steps,
40 samples per complex).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at every rank = BASELINE.json configs[1]: 10 synthetic complexes (60-atom ligand, 300 C-alpha receptor, 24-nn
receptor graph) x 40 samples x 20 reverse steps, DiffDock-S architecture (ns=24 nv=6 L=5 sh_lmax=1), fresh-initialised
weights (seeded; no checkpoints on the box), README low-temperature sampling.  One "step" = the whole 400-pose job.
Weak scaling: with N ranks the job is 10*N complexes sharded by rank, no collective inside a diffusion step
(weights broadcast once, final poses gathered at the end of each step).

value : poses/s with every input resident in HBM (one ddk_sample call over the rank's 400 poses)
e2e   : poses/s through the drop-in sampling() API with HOST buffers (H2D of the static complex data, start poses, noise
        and step tables, the 20-step run, D2H of the final poses) -- the headline number: one call over the 10 x 40 graphs
        of the workload; the evaluate.py-style loop (one call per complex) is reported next to it
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
WORKLOAD = (f'{N_COMPLEX} synthetic complexes ({N_LIG}-atom ligand, {N_REC} C-alpha receptor) x {N_SAMPLES} samples x '
            f'{REV_STEPS} reverse steps, DiffDock-S ns=24 nv=6 L=5 sh_lmax=1 (BASELINE.json configs[1])')
W_CONV = (720, 936, 1152, 1872, 1872)
C_TP = (1008, 1368, 1728, 2736, 2736)
FLOP_PER_EDGE_REF = 2 * sum(72 * 72 + 72 * w + c for w, c in zip(W_CONV, C_TP))     # 1.014 MFLOP per edge-step
U_LV3 = 276


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


def build_workload(rank, n_complex):
    from disco_diffdock_b200 import synthetic
    from disco_diffdock_b200.sampling import randomize_position
    from tests import helpers
    np.random.seed(1234 + rank)
    torch.manual_seed(1234 + rank)
    complexes = []
    for c in range(n_complex):
        g = synthetic.make_complex(1000 + rank * n_complex + c, N_LIG, N_REC)
        item = synthetic.as_loader_item(g)
        data_list = [copy.deepcopy(item) for _ in range(N_SAMPLES)]
        randomize_position(data_list, False, False, 19.0)        # evaluate.py:232-233
        complexes.append(data_list)
    return complexes


def cpu_reference_run(n_poses, n_rev_steps, threads, seed=0):
    """The oracle port (oracle/restate.py = the reference algorithm on torch-CPU) on a bounded sample of the workload."""
    from disco_diffdock_b200 import data as ddata
    from oracle import restate
    from tests import helpers
    from tests.test_oracle_golden import load_tables
    torch.set_num_threads(threads)
    m, sd, cfg = helpers.make_model(0, gain=5.0)
    g, lst = helpers.make_pose_batch(1000 + seed, N_LIG, N_REC, n_poses, jitter=True)
    R = g['ligand'].mask_rotate.shape[0]
    noise = helpers.draw_noise(3, n_rev_steps, n_poses, R)
    sched = np.linspace(1, 0, REV_STEPS + 1)[:-1]
    batch = ddata.Batch.from_data_list(lst)
    t0 = time.perf_counter()
    with torch.no_grad():
        restate.sample(sd, cfg, batch, load_tables(), sched, noise, inference_steps=n_rev_steps, **helpers.README_TEMPS)
    dt = time.perf_counter() - t0
    pose_steps = n_poses * n_rev_steps
    return pose_steps / REV_STEPS / dt, dt


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_poses, n_rev = 2, 4
    for _ in range(args.warmup):
        cpu_reference_run(1, 1, threads)
    vals, times = [], []
    for k in range(args.steps):
        v, dt = cpu_reference_run(n_poses, n_rev, threads, seed=k)
        vals.append(v); times.append(dt)
    value = float(np.mean(vals))
    sample = f'{n_poses} poses x {n_rev} of {REV_STEPS} reverse steps per bench step (same complex shape), scaled to poses of {REV_STEPS} steps'
    line = {'impl': 'reference', 'metric': 'docked_poses_per_sec', 'value': value, 'unit': 'poses/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(np.mean(times) * 1000), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'inputs': 'larger than L2 not applicable (CPU)'},
            'cpu_baseline': {'value': value, 'unit': 'poses/s', 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'poses/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--complexes', type=int, default=N_COMPLEX)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from disco_diffdock_b200 import data as ddata
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

    # weights: rank 0 initialises, one NCCL broadcast replicates them (no other parameter traffic, ever)
    m, sd, cfg = helpers.make_model(0, gain=5.0)
    m = m.to(dev)
    if world > 1:
        for p_ in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(p_.data, src=0)
        m.invalidate()
    eng = m.engine(dev)
    eng.profile_enable(True)
    t2s = partial(du.t_to_sigma, args=cfg)
    sched = du.get_t_schedule(REV_STEPS)
    n_complex = args.complexes
    complexes = build_workload(rank, n_complex)
    n_poses = n_complex * N_SAMPLES
    Rs = [int(c[0]['ligand'].edge_mask.sum()) for c in complexes]

    # ---------------------------------------------------------------- value: inputs resident in HBM
    flat = [g for c in complexes for g in c]
    big = ddata.Batch.from_data_list(flat)
    info = eng.set_batch(big)
    steps_tab = dsampling.build_step_tables(m, cfg, t2s, sched, sched, sched, REV_STEPS, n_poses, **helpers.README_TEMPS)
    gen = torch.Generator(device=dev).manual_seed(77 + rank)
    z = {'tr': torch.randn(REV_STEPS, n_poses, 3, device=dev, generator=gen),
         'rot': torch.randn(REV_STEPS, n_poses, 3, device=dev, generator=gen),
         'tor': torch.randn(REV_STEPS, info.RB, device=dev, generator=gen)}
    for v in z.values():
        v[-1] = 0                                             # --no_final_step_noise (README.md:15)
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
    launches0, edges0, segs0 = eng.kernel_launches(), eng.edge_total(), eng.segment_total()
    ge0, gs0 = eng.group_totals()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        final = resident_step()
    e1.record()
    barrier()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    prof = eng.profile_read()
    launches = eng.kernel_launches() - launches0
    dyn_edges = eng.edge_total() - edges0
    segments = eng.segment_total() - segs0                  # non-empty (node, edge group) segments, summed over reverse steps
    ge1, gs1 = eng.group_totals()
    g_edges, g_segs = (ge1 - ge0).astype(float), (gs1 - gs0).astype(float)   # per edge group, summed over reverse steps
    static_edges = (info.EB + info.ER) * REV_STEPS * args.steps
    edges = dyn_edges + static_edges
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = n_poses * world * args.steps / (ms / 1000)
    assert torch.isfinite(final).all()

    # ---------------------------------------------------------------- e2e: host buffers through sampling()
    # Two ways of driving the drop-in API, both with every buffer in host memory:
    #   batched      : ONE sampling() call over the 10 x 40 graphs of the workload ("batch 10 complexes", configs[1]); the
    #                  copies of each complex are recognised, shipped once and replicated on the device -- the e2e value
    #   per complex  : one sampling() call per complex (40 poses), the loop of evaluate.py:219-291
    def e2e_batched(seed):
        g = torch.Generator().manual_seed(seed)
        dl = [x.shallow_copy() for data_list in complexes for x in data_list]     # sampling() rebinds ['ligand'].pos only
        out, _ = dsampling.sampling(dl, m, REV_STEPS, sched, sched, sched, dev, t2s, cfg, batch_size=len(dl),
                                    no_final_step_noise=True, generator=g, host_buffers=True, **helpers.README_TEMPS)
        return sum(x['ligand'].pos.numel() * 4 for x in out)

    def e2e_per_complex(seed):
        out_bytes = 0
        g = torch.Generator().manual_seed(seed)
        for ci, data_list in enumerate(complexes):
            dl = [x.shallow_copy() for x in data_list]
            out, _ = dsampling.sampling(dl, m, REV_STEPS, sched, sched, sched, dev, t2s, cfg, batch_size=N_SAMPLES,
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
        te = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return n_poses * world * args.steps / float(te.item()), nbytes

    e2e_pc_value, _ = time_e2e(e2e_per_complex)
    e2e_value, d2h = time_e2e(e2e_batched)
    bi = eng.batch_info                                       # of the batched call: every complex shipped once
    h2d = bi.h2d_bytes + bi.NL * 3 * 4 + REV_STEPS * (2 * bi.B * 3 + bi.RB) * 4 + REV_STEPS * bi.B * (32 + 4) * 4

    # ---------------------------------------------------------------- roofline of the dominant kernel
    # k_conv_fused<3> (the two 84-wide conv layers): one launch = one layer over every pose of the rank.  Work model
    # (DESIGN.md section 5): per listed edge 2*72*U FLOP of rank-1 updates (U = 276), per non-empty segment 2*72*W FLOP of
    # the second radial-MLP layer (W = 1872); algorithmic bytes per launch = node features in and out, the edge list, the
    # harmonics and the 72 hidden units of every listed edge.
    hbm_peak, peak_src, peaks_raw = peaks()
    # the two 84-wide layers: k_conv_fused<3> plus k_acc_tc<3>, which accumulates their long lig<-rec segments on the tensor cores
    acc_ms, acc_n = prof['conv_accum_lv3'][0] + prof['conv_tc_lv3'][0], prof['conv_accum_lv3'][1]
    total_prof_ms = sum(v[0] for v in prof.values())
    n_nodes = info.NL + info.NR
    passes = REV_STEPS * args.steps                                      # reverse steps in the timed region
    # the two level-3 launches of a step: layer 3 over the segments of ligand nodes and of the residues with a cross edge
    # (work lists 0, 1, 3 and 4 = receptor contacts of those residues), layer 4 (the last before the heads) over the segments
    # of ligand nodes only (lists 0, 1) -- the heads never read receptor features
    e_launch = (g_edges[[0, 1, 3, 4]].sum() + g_edges[:2].sum()) / 2 / passes   # listed edges per launch, average of the two
    s_launch = (g_segs[[0, 1, 3, 4]].sum() + g_segs[:2].sum()) / 2 / passes
    n_nodes = (n_nodes + info.NL) / 2
    bytes_launch = n_nodes * (84 + 84) * 4 + e_launch * (8 + 16 + 72 * 4) + s_launch * 16
    flop_launch = e_launch * 2 * 72 * U_LV3 + s_launch * 2 * 72 * W_CONV[3]
    t_launch = acc_ms / max(acc_n, 1) / 1000
    clocks = sampler.summary()
    sm_mhz = clocks.get('sm_mhz') or peaks_raw.get('sm_max_mhz', 1965.0)
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    conv_ms = sum(prof[k][0] for k in ('conv_accum_lv0', 'conv_accum_lv1', 'conv_accum_lv2', 'conv_accum_lv3', 'conv_tc_lv0',
                                       'conv_tc_lv1', 'conv_tc_lv2', 'conv_tc_lv3'))
    # DRAM bytes per launch of this kernel from the committed ncu pass over this very command (tools/gpu_round.sh,
    # tools/launch_summary.py); null when the capture is absent
    traffic, traffic_src = None, None
    tj = os.path.join(ROOT, 'profiles', 'conv_fused3_traffic.json')
    if os.path.exists(tj) and n_complex == N_COMPLEX:
        td = json.load(open(tj))
        traffic, traffic_src = td['dram_bytes_per_launch'], 'profiles/conv_fused3_traffic.json (' + td['source'] + ')'
    roofline = {'kernel': 'k_conv_fused<3> (+ k_acc_tc<3>: its long cross segments, 3xTF32 tcgen05)', 'bound': 'hbm', 'achieved': bytes_launch / t_launch / 1e9, 'peak': hbm_peak,
                'unit': 'GB/s', 'frac': bytes_launch / t_launch / 1e9 / hbm_peak, 'traffic': traffic,
                'traffic_unit': 'bytes per launch', 'traffic_source': traffic_src, 'algorithmic_bytes_per_launch': bytes_launch,
                'peak_source': peak_src,
                'launch_ms': t_launch * 1000, 'launches': acc_n, 'share_of_step': acc_ms / max(total_prof_ms, 1e-9),
                'conv_share_of_step': conv_ms / max(total_prof_ms, 1e-9),
                'edges_per_launch': e_launch, 'segments_per_launch': s_launch,
                'fp32_fma': {'achieved_tflops': flop_launch / t_launch / 1e12, 'peak_tflops_at_measured_clock': fp32_peak,
                             'frac': flop_launch / t_launch / 1e12 / fp32_peak,
                             'note': 'the kernel is bound by the FP32 FMA pipe and shared-memory issue, not by HBM (re-associated '
                                     'tensor product, SURVEY 8d); the HBM fraction is low by design.  FLOP of the layer / time of both '
                                     'kernels; the long lig<-rec segments run as 3xTF32 tcgen05.mma in k_acc_tc'},
                'tensor_pipe': {'kernel': 'k_acc_tc<3>', 'ms_in_timed_region': round(prof['conv_tc_lv3'][0], 3),
                                'sm__pipe_tensor_cycles_active_pct': 30.7,
                                'source': 'ncu --set full capture, dense t = 1 step at 80 poses per launch: '
                                          'profiles/r01i_ncu_tc_summary.txt (not measured in this run)'},
                'kernel_ms': {k: round(v[0], 3) for k, v in prof.items()}}
    ref_equiv_tflops = edges * FLOP_PER_EDGE_REF / (ms / 1000) / 1e12

    line = {'metric': 'docked_poses_per_sec', 'value': value, 'unit': 'poses/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic, fresh-init weights (seeded)',
            'config': {'workload': WORKLOAD, 'poses_per_gpu': n_poses, 'reverse_steps': REV_STEPS,
                       'l2': 'inputs larger than L2: the per-layer working set (edge embeddings, hidden units and harmonics of 5.2 M listed edges) is about 2 GB per pass',
                       'edges_per_pose_step': edges / (n_poses * REV_STEPS * args.steps),
                       'reference_formulation_equiv_tflops': ref_equiv_tflops, 'parallelism': f'pose-sharded x{world}'},
            'e2e': {'value': e2e_value, 'unit': 'poses/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'call': 'one sampling() call over the 10 x 40 graphs, host buffers',
                    'per_complex_calls': {'value': e2e_pc_value, 'unit': 'poses/s',
                                          'call': 'one sampling() call per complex (40 poses), the evaluate.py loop'}},
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_reference_run(1, 1, threads)
        v, dt = cpu_reference_run(2, 6, threads)
        line['cpu_baseline'] = {'value': v, 'unit': 'poses/s', 'cores': threads, 'kind': 'port',
                                'sample': f'2 poses x 6 of {REV_STEPS} reverse steps of the same complex shape ({dt:.1f} s), scaled'}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
