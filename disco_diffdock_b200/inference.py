"""The loop of ``evaluate.py`` around the sampler (``/root/reference/evaluate.py:219-400, 441-448``), SURVEY 8f-2 / f-4.

``evaluate.py`` walks the test set one complex at a time: deep-copy the complex ``samples_per_complex`` times,
``randomize_position``, ``sampling()``, geometry metrics of the poses, then ``np.save`` of the per-complex arrays.  On
a B200 one complex (40 poses) does not fill the GPU (profiles/README.md, "small batches"), and the host work of the
next complex (copies, random start poses) and of the previous one (metrics) sits between two sampler calls.
``run_inference`` keeps the per-complex semantics -- same calls, same outputs, same file names -- and

* samples ``complexes_per_call`` complexes in one ``sampling()`` call (their copies form one batch; poses are
  independent, so the result of a complex does not depend on what it is batched with), and
* prepares the next call's start poses on a worker thread while the GPU runs (``libddk`` releases the GIL).

Metrics that need rdkit / spyrmsd (symmetry-corrected RMSD, evaluate.py:309-310) are out of scope; the plain RMSD the
reference falls back to (evaluate.py:313) is used.  ``save_complex_pack`` / ``load_complex_pack`` store the tensors the
sampler reads in one flat ``.npz`` so a box without PyG / rdkit can run the sampler.
"""
from __future__ import annotations

import copy
import os
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import data as ddata
from .diffusion_utils import get_t_schedule
from .sampling import randomize_position, sampling


def _name(g):
    n = g['name'] if 'name' in g else None
    while isinstance(n, (list, tuple)):
        n = n[0]
    return str(n) if n is not None else ''


def _as_np(x):
    while isinstance(x, (list, tuple)):
        x = x[0]
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def pose_metrics(orig, data_list, receptor_pos: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """evaluate.py:301-337 for one complex: heavy-atom poses, plain RMSD to the crystal pose (:313), centroid distance
    (:315), minimum ligand self distance (:332-334) and minimum ligand-receptor distance (:330-331; C-alpha positions
    unless all-atom ``receptor_pos`` [n,3] is given, as evaluate.py reads them from the PDB file)."""
    filt = torch.not_equal(data_list[0]['ligand'].x[:, 0], 0).cpu().numpy()          # :299 (atomic-number class 0 = H)
    pos = np.asarray([g['ligand'].pos.detach().cpu().numpy()[filt] for g in data_list])
    out = {'ligand_pos': pos}
    lig = orig['ligand']
    if 'orig_pos' in lig:
        centre = _as_np(orig.original_center) if 'original_center' in orig else np.zeros((1, 3), np.float32)
        ref = (_as_np(lig.orig_pos)[filt] - centre)[None]
        out['rmsd'] = np.sqrt(((pos - ref) ** 2).sum(axis=2).mean(axis=1))
        out['centroid_distance'] = np.linalg.norm(pos.mean(axis=1) - ref.mean(axis=1), axis=1)
    sd = np.linalg.norm(pos[:, :, None, :] - pos[:, None, :, :], axis=-1)
    sd = np.where(np.eye(sd.shape[2]), np.inf, sd)
    out['min_self_distance'] = sd.min(axis=(1, 2))
    rp = receptor_pos if receptor_pos is not None else data_list[0]['receptor'].pos.detach().cpu().numpy()
    cd = np.linalg.norm(rp[None, :, None, :] - pos[:, None, :, :], axis=-1)
    out['min_cross_distance'] = cd.min(axis=(1, 2))
    return out


def run_inference(complexes: Sequence, model, model_args, device, t_to_sigma, *, samples_per_complex=40,
                  inference_steps=20, actual_steps=None, complexes_per_call=1, no_torsion=None, no_random=False,
                  no_final_step_noise=False, ode=False, temp_sampling=1.0, temp_psi=0.0, temp_sigma_data=0.5,
                  out_dir: Optional[str] = None, generator=None, noise_fn=None, host_buffers=True, prefetch=True,
                  metrics=True) -> Dict[str, object]:
    """Sample ``samples_per_complex`` poses for every complex.  Returns ``{'names', 'ligand_pos' (list of [N, heavy, 3]),
    'rmsds', 'centroid_distances', 'min_self_distances', 'min_cross_distances', 'run_times', 'data_lists'}`` and, with
    ``out_dir``, writes the arrays under the file names of evaluate.py:441-448.

    ``noise_fn(first_complex_index, data_list) -> {'tr','rot','tor'}`` injects pre-drawn noise (parity tests)."""
    N = samples_per_complex
    schedule = get_t_schedule(inference_steps=inference_steps)                         # evaluate.py:209-211
    steps = actual_steps if actual_steps is not None else inference_steps              # :269
    no_torsion = bool(getattr(model_args, 'no_torsion', False)) if no_torsion is None else no_torsion
    chunks = [list(range(i, min(i + complexes_per_call, len(complexes)))) for i in range(0, len(complexes), complexes_per_call)]

    def prepare(idx):
        lists = []
        for i in idx:
            dl = [copy.deepcopy(complexes[i]) for _ in range(N)]                       # :229
            randomize_position(dl, no_torsion, no_random, model_args.tr_sigma_max)     # :230
            lists.append(dl)
        return lists

    res = {k: [] for k in ('names', 'ligand_pos', 'rmsds', 'centroid_distances', 'min_self_distances',
                           'min_cross_distances', 'run_times', 'data_lists')}
    pool = ThreadPoolExecutor(1) if prefetch and len(chunks) > 1 else None
    if pool is not None and generator is None and noise_fn is None and not (no_random or ode):
        # The prefetch thread draws the start poses from the GLOBAL numpy / torch / scipy generators (randomize_position follows
        # the reference's RNG calls).  The noise of the reverse steps must not interleave with those draws, or a seeded run would
        # depend on thread timing: it gets its own generator, seeded from the global one before the worker starts.
        zdev = torch.device('cpu') if host_buffers else torch.device(device)
        generator = torch.Generator(device=zdev)
        generator.manual_seed(int(torch.randint(0, 2 ** 62, (1,)).item()))
    nxt = pool.submit(prepare, chunks[0]) if pool else None
    for ci, idx in enumerate(chunks):
        lists = nxt.result() if pool else prepare(idx)
        if pool and ci + 1 < len(chunks):
            nxt = pool.submit(prepare, chunks[ci + 1])
        flat = [g for dl in lists for g in dl]
        t0 = time.time()
        sampling(data_list=flat, model=model, inference_steps=steps, tr_schedule=schedule, rot_schedule=schedule,
                 tor_schedule=schedule, device=device, t_to_sigma=t_to_sigma, model_args=model_args, no_random=no_random,
                 ode=ode, batch_size=len(flat), no_final_step_noise=no_final_step_noise, temp_sampling=temp_sampling,
                 temp_psi=temp_psi, temp_sigma_data=temp_sigma_data, generator=generator,
                 noise=noise_fn(idx[0], flat) if noise_fn is not None else None, host_buffers=host_buffers)
        dt = time.time() - t0
        for k, i in enumerate(idx):
            dl = lists[k]
            res['names'].append(_name(complexes[i]))
            res['run_times'].append(dt / len(idx))                                     # evaluate.py:293 per complex
            res['data_lists'].append(dl)
            if metrics:
                pm = pose_metrics(complexes[i], dl)
                res['ligand_pos'].append(pm['ligand_pos'])
                res['min_self_distances'].append(pm['min_self_distance'])
                res['min_cross_distances'].append(pm['min_cross_distance'])
                if 'rmsd' in pm:
                    res['rmsds'].append(pm['rmsd'])
                    res['centroid_distances'].append(pm['centroid_distance'])
    if pool:
        pool.shutdown()
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        save = lambda name, v: np.save(os.path.join(out_dir, name), np.array(v))
        save('min_cross_distances.npy', res['min_cross_distances'])
        save('min_self_distances.npy', res['min_self_distances'])
        save('rmsds.npy', res['rmsds'])
        save('centroid_distances.npy', res['centroid_distances'])
        save('run_times.npy', res['run_times'])
        save('complex_names.npy', res['names'])
    return res


# ---------------------------------------------------------------------------------------------- complex pack (f-4)
_PACK_FIELDS = (('ligand', 'x', np.int64), ('ligand', 'pos', np.float32), ('ligand', 'edge_mask', np.bool_),
                ('receptor', 'x', np.float32), ('receptor', 'pos', np.float32))


def save_complex_pack(path: str, complexes: Sequence) -> None:
    """All tensors ``TensorProductScoreModel.forward`` / ``sampling()`` read from a complex, for a list
    of complexes, as one flat ``.npz``: ``<i>/<store>/<field>`` arrays plus the bond / contact edge lists, ``mask_rotate``,
    ``orig_pos``, ``original_center`` and the name."""
    out = {'n': np.asarray(len(complexes))}
    for i, g in enumerate(complexes):
        for store, field, dt in _PACK_FIELDS:
            out[f'{i}/{store}/{field}'] = g[store][field].detach().cpu().numpy().astype(dt)
        out[f'{i}/bonds/edge_index'] = g['ligand', 'ligand'].edge_index.cpu().numpy().astype(np.int64)
        out[f'{i}/bonds/edge_attr'] = g['ligand', 'ligand'].edge_attr.cpu().numpy().astype(np.float32)
        out[f'{i}/contacts/edge_index'] = g['receptor', 'receptor'].edge_index.cpu().numpy().astype(np.int64)
        out[f'{i}/ligand/mask_rotate'] = np.asarray(_as_np(g['ligand'].mask_rotate), dtype=np.bool_)
        if 'orig_pos' in g['ligand']:
            out[f'{i}/ligand/orig_pos'] = _as_np(g['ligand'].orig_pos).astype(np.float32)
        if 'original_center' in g:
            out[f'{i}/original_center'] = _as_np(g.original_center).astype(np.float32)
        out[f'{i}/name'] = np.asarray(_name(g))
    np.savez(path, **out)


def load_complex_pack(path: str) -> List[ddata.HeteroData]:
    z = np.load(path, allow_pickle=False)
    out = []
    for i in range(int(z['n'])):
        g = ddata.HeteroData()
        for store, field, _ in _PACK_FIELDS:
            setattr(g[store], field, torch.from_numpy(z[f'{i}/{store}/{field}']))
        g['ligand', 'ligand'].edge_index = torch.from_numpy(z[f'{i}/bonds/edge_index'])
        g['ligand', 'ligand'].edge_attr = torch.from_numpy(z[f'{i}/bonds/edge_attr'])
        g['receptor', 'receptor'].edge_index = torch.from_numpy(z[f'{i}/contacts/edge_index'])
        g['ligand'].mask_rotate = z[f'{i}/ligand/mask_rotate']
        if f'{i}/ligand/orig_pos' in z:
            g['ligand'].orig_pos = z[f'{i}/ligand/orig_pos']
        if f'{i}/original_center' in z:
            g.original_center = torch.from_numpy(z[f'{i}/original_center'])
        g.name = str(z[f'{i}/name'])
        out.append(g)
    return out


# ---------------------------------------------------------------------------------------------- pose-sharded driver
def pose_noise(seed: int, ci: int, k: int, steps: int, n_rot: int) -> Dict[str, torch.Tensor]:
    """The noise stream of pose (complex ci, sample k): keyed by the pose id, not by rank, batch or call order, so a sharded run
    consumes exactly the z of the unsharded one.  Draw order tr, rot, tor like utils/sampling.py:146-165."""
    g = torch.Generator().manual_seed(((int(seed) * 1000003 + int(ci)) * 8191 + int(k)) % (2 ** 63 - 1))
    return {'tr': torch.randn(steps, 3, generator=g), 'rot': torch.randn(steps, 3, generator=g),
            'tor': torch.randn(steps, n_rot, generator=g)}


def seeded_start_poses(complex_graph, ci: int, n: int, seed: int, no_torsion: bool, no_random: bool, tr_sigma_max: float):
    """evaluate.py:229-233 for one complex with the global numpy / torch generators seeded from (seed, ci) -- every rank that owns
    a sample of the complex derives the same ``n`` start poses -- and restored afterwards."""
    np_state, torch_state = np.random.get_state(), torch.random.get_rng_state()
    try:
        s = (int(seed) * 7919 + int(ci) * 104729 + 17) % (2 ** 32)
        np.random.seed(s)
        torch.manual_seed(s)
        dl = [copy.deepcopy(complex_graph) for _ in range(n)]
        randomize_position(dl, no_torsion, no_random, tr_sigma_max)
        return dl
    finally:
        np.random.set_state(np_state)
        torch.random.set_rng_state(torch_state)


def run_inference_sharded(complexes: Sequence, model, model_args, device, t_to_sigma, *, samples_per_complex=40,
                          inference_steps=20, actual_steps=None, seed=0, rank: Optional[int] = None, world: Optional[int] = None,
                          poses_per_call=400, no_torsion=None, no_random=False, no_final_step_noise=False, ode=False,
                          temp_sampling=1.0, temp_psi=0.0, temp_sigma_data=0.5, host_buffers=True, broadcast_weights=True,
                          gather=True, sampler=None, **sampling_kw) -> Dict[str, object]:
    """``samples_per_complex`` poses for every complex, the (complex, sample) poses sharded over the ranks of the process group
    (one process per GPU; evaluate.py:219-400 is the single-process loop this replaces).  One broadcast of the weights, then no
    communication until ONE all-gather of the final poses; ranks take contiguous pose ranges balanced by N_l * N_r
    (``dist.shard_poses``).  Start poses and noise are keyed by (seed, complex, sample), and a pose's trajectory does not depend
    on what it is batched with (tests), so the result is bit-identical for every world size.  Extra keyword arguments (``ar_model``,
    ``classifier_free_guidance_weight``, ``cfg_start`` ... of the DisCo path) are handed to ``sampling()``.

    Returns ``{'names', 'ligand_pos': [per complex: tensor [samples, N_l, 3]], 'shard': [(complex, first, stop)], 'run_time'}``;
    with ``gather=False`` (or on ranks > 0 when the backend cannot gather) only the rank's own poses are filled in."""
    from . import dist as ddist
    import torch.distributed as tdist
    if world is None:
        world = tdist.get_world_size() if tdist.is_initialized() else 1
    if rank is None:
        rank = tdist.get_rank() if tdist.is_initialized() else 0
    sampler = sampler or sampling
    N = samples_per_complex
    schedule = get_t_schedule(inference_steps=inference_steps)
    steps = actual_steps if actual_steps is not None else inference_steps
    no_torsion = bool(getattr(model_args, 'no_torsion', False)) if no_torsion is None else no_torsion
    if broadcast_weights and tdist.is_initialized() and world > 1:
        sm = model.score_model if hasattr(model, 'score_model') else model
        ddist.broadcast_module(sm, src=0)
        if hasattr(sm, 'invalidate'):
            sm.invalidate()
    cost = [float(g['ligand'].pos.shape[0]) * float(g['receptor'].pos.shape[0]) for g in complexes]
    shard = ddist.shard_poses([N] * len(complexes), cost, rank, world)
    n_rot = [int(g['ligand'].edge_mask.sum()) for g in complexes]
    local: Dict[tuple, torch.Tensor] = {}
    t0 = time.time()
    # calls of at most poses_per_call poses, whole shard entries per call (an entry = a run of copies of one complex)
    calls, cur, cur_n = [], [], 0
    for ci, a, b in shard:
        k = a
        while k < b:
            take = min(b - k, max(1, poses_per_call - cur_n))
            cur.append((ci, k, k + take)); cur_n += take; k += take
            if cur_n >= poses_per_call:
                calls.append(cur); cur, cur_n = [], 0
    if cur:
        calls.append(cur)
    start_cache: Dict[int, list] = {}
    for call in calls:
        flat, zs = [], []
        for ci, a, b in call:
            if ci not in start_cache:
                start_cache = {ci: seeded_start_poses(complexes[ci], ci, N, seed, no_torsion, no_random, model_args.tr_sigma_max)}
            flat += start_cache[ci][a:b]
            zs += [pose_noise(seed, ci, k, steps, n_rot[ci]) for k in range(a, b)]
        noise = None
        if not (no_random or ode):
            noise = {'tr': torch.stack([z['tr'] for z in zs], dim=1), 'rot': torch.stack([z['rot'] for z in zs], dim=1),
                     'tor': torch.cat([z['tor'] for z in zs], dim=1)}
        sampler(data_list=flat, model=model, inference_steps=steps, tr_schedule=schedule, rot_schedule=schedule,
                tor_schedule=schedule, device=device, t_to_sigma=t_to_sigma, model_args=model_args, no_random=no_random, ode=ode,
                batch_size=len(flat), no_final_step_noise=no_final_step_noise, temp_sampling=temp_sampling, temp_psi=temp_psi,
                temp_sigma_data=temp_sigma_data, noise=noise, host_buffers=host_buffers, **sampling_kw)
        i = 0
        for ci, a, b in call:
            for k in range(a, b):
                local[(ci, k)] = flat[i]['ligand'].pos.detach().to('cpu', torch.float32)
                i += 1
    run_time = time.time() - t0
    # ---- one gather of the final poses: [pose id, padded coordinates]
    max_nl = max(int(g['ligand'].pos.shape[0]) for g in complexes)
    ids = sorted(local.keys())
    buf = torch.zeros(len(ids), max_nl, 3)
    for r, key in enumerate(ids):
        buf[r, :local[key].shape[0]] = local[key]
    idt = torch.tensor(ids, dtype=torch.int64).reshape(-1, 2)
    if gather and tdist.is_initialized() and world > 1:
        counts = []
        for r in range(world):
            counts.append(sum(b - a for _, a, b in ddist.shard_poses([N] * len(complexes), cost, r, world)))
        gdev = torch.device(device) if tdist.get_backend() == 'nccl' else torch.device('cpu')
        buf = ddist.gather_poses(buf.to(gdev), counts).cpu()
        idt = ddist.gather_poses(idt.to(gdev), counts).cpu()
    out_pos = [torch.full((N, int(g['ligand'].pos.shape[0]), 3), float('nan')) for g in complexes]
    for r in range(idt.shape[0]):
        ci, k = int(idt[r, 0]), int(idt[r, 1])
        out_pos[ci][k] = buf[r, :out_pos[ci].shape[1]]
    return {'names': [_name(g) for g in complexes], 'ligand_pos': out_pos, 'shard': shard, 'run_time': run_time}
