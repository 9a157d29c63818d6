"""Drop-in for ``utils/sampling.py`` of the reference: ``randomize_position`` and ``sampling`` with the same
signatures (``/root/reference/utils/sampling.py:12, 49-54``; called from ``evaluate.py:233, 268-291``).

``sampling()`` keeps the reference's contract -- it mutates ``data_list[i]['ligand'].pos`` and returns
``(data_list, confidence)`` -- but the whole reverse-diffusion loop of a batch (``sampling.py:105-198``: set_time,
score model, perturbation, modify_conformer_batch) runs inside ``libddk`` as one stream-ordered sequence of kernels
with no host synchronisation (``ddk_sample``).  Everything that depends only on the schedule (sigma(t), the sinusoidal
embedding, the cross cut-off, SO(3)/torus score-norm look-ups and the a*score + b*z coefficients) is computed once on
the host for all steps.

Two keyword-only extensions: ``noise`` (pre-drawn z, used by the parity tests so the CPU oracle and the GPU consume
identical noise) and ``generator`` (seeded device RNG; the reference never seeds).
"""
from __future__ import annotations

import copy
from typing import Dict, Optional

import numpy as np
import torch

from .data import Batch, DataLoader
from .engine import StepTables, so3_score_norm, torus_score_norm


def is_iterable(arr):
    try:
        iter(arr)
        return True
    except TypeError:
        return False


def _three(x):
    return list(x) if is_iterable(x) else [x] * 3


def _score_model_of(model):
    m = model.module if hasattr(model, 'module') else model
    return m.score_model if hasattr(m, 'score_model') else m


def randomize_position(data_list, no_torsion, no_random, tr_sigma_max, unbatched=False, ar_args=None):
    """sampling.py:12-46: uniform torsions, uniform random rotation about the centroid, N(0, tr_sigma_max) shift.
    Host code (numpy / scipy RNG), once per complex -- it defines the start poses."""
    from scipy.spatial.transform import Rotation as R
    if not no_torsion:
        for g in data_list:
            lig = g['ligand']
            mask = lig.edge_mask.cpu().numpy().astype(bool)
            upd = np.random.uniform(low=-np.pi, high=np.pi, size=int(mask.sum()))
            mr = lig.mask_rotate if unbatched else lig.mask_rotate[0]
            edges = g['ligand', 'ligand'].edge_index.T[torch.from_numpy(mask)]
            lig.pos = torsion_update_host(lig.pos, edges, np.asarray(mr), upd)
    for g in data_list:
        lig = g['ligand']
        center = torch.mean(lig.pos, dim=0, keepdim=True)
        rot = torch.from_numpy(R.random().as_matrix()).float()
        lig.pos = (lig.pos - center) @ rot.T
        if not no_random:
            lig.pos = lig.pos + torch.normal(mean=0, std=tr_sigma_max, size=(1, 3))
    if ar_args is not None:
        for g in data_list:
            if getattr(ar_args, 'no_randomness', False):
                ar = torch.from_numpy(np.asarray(g['ligand'].orig_rdkit_pos[0])).float()
                ar = (ar - ar.mean(0, keepdim=True)) @ torch.from_numpy(R.random().as_matrix()).float().T
                g['ligand'].ar_pos = ar
            else:
                g['ligand'].ar_pos = copy.deepcopy(g['ligand'].pos)


def torsion_update_host(pos, edge_index, mask_rotate, torsion_updates):
    """utils/torsion.py:48-68 (numpy, sequential over rotatable bonds); used only by randomize_position."""
    from scipy.spatial.transform import Rotation as R
    p = pos.cpu().numpy().astype(np.float64).copy() if torch.is_tensor(pos) else np.array(pos, dtype=np.float64)
    for k, e in enumerate(edge_index.cpu().numpy()):
        if torsion_updates[k] == 0:
            continue
        u, v = int(e[0]), int(e[1])
        assert not mask_rotate[k, u] and mask_rotate[k, v]
        rv = p[u] - p[v]
        rv = rv * torsion_updates[k] / np.linalg.norm(rv)
        rm = R.from_rotvec(rv).as_matrix()
        p[mask_rotate[k]] = (p[mask_rotate[k]] - p[v]) @ rm.T + p[v]
    return torch.from_numpy(p.astype(np.float32))


def _same_complex(a, b):
    """Is graph ``b`` a copy of complex ``a`` (evaluate.py:229 deep-copies one complex N times)?  Shared storage answers
    without reading; otherwise every static tensor must have the same shape and the receptor positions, ligand atom
    features and bonds the same content (every reference caller passes copies of ONE complex; two different complexes
    with identical C-alpha coordinates, ligand atoms and bonds do not occur)."""
    la, lb, ra, rb = a['ligand'], b['ligand'], a['receptor'], b['receptor']
    if la.x.shape != lb.x.shape or ra.x.shape != rb.x.shape:
        return False
    ea, eb = a['ligand', 'ligand'].edge_index, b['ligand', 'ligand'].edge_index
    if ra.x.data_ptr() == rb.x.data_ptr() and la.x.data_ptr() == lb.x.data_ptr() and ea.data_ptr() == eb.data_ptr():
        return True
    if ea.shape != eb.shape or a['receptor', 'receptor'].edge_index.shape != b['receptor', 'receptor'].edge_index.shape:
        return False
    return bool(torch.equal(ra.pos, rb.pos) and torch.equal(la.x, lb.x) and torch.equal(ea, eb))


def group_copies(items):
    """Runs of consecutive copies of one complex: [(first graph of the run, run length), ...]."""
    groups = []
    for g in items:
        if groups and _same_complex(groups[-1][0], g):
            groups[-1][1] += 1
        else:
            groups.append([g, 1])
    return [(g, n) for g, n in groups]


_STEP_TABLE_CACHE = {}


def build_step_tables(score_model, model_args, t_to_sigma, tr_schedule, rot_schedule, tor_schedule, inference_steps, B,
                      temp_sampling, temp_psi, temp_sigma_data, ode=False) -> StepTables:
    """Host arithmetic of sampling.py:106-113, 137-192 and of the sigma-dependent parts of
    TensorProductScoreModel.forward (score_model.py:187, 203, 276, 284-286, 303-307) for every step.  The tables depend only
    on the schedule, the noise-level hyper-parameters, the temperatures and the batch size, so evaluate.py's loop over
    complexes reuses them (a small keyed cache)."""
    try:
        key = (id(score_model), inference_steps, B, bool(ode), tuple(np.asarray(tr_schedule, dtype=np.float64).tolist()),
               tuple(np.asarray(rot_schedule, dtype=np.float64).tolist()), tuple(np.asarray(tor_schedule, dtype=np.float64).tolist()),
               tuple(_three(temp_sampling)), tuple(_three(temp_psi)), tuple(_three(temp_sigma_data)),
               tuple(float(getattr(model_args, k)) for k in ('tr_sigma_min', 'tr_sigma_max', 'rot_sigma_min', 'rot_sigma_max',
                                                              'tor_sigma_min', 'tor_sigma_max')))
    except Exception:
        key = None
    if key is not None and key in _STEP_TABLE_CACHE:
        return _STEP_TABLE_CACHE[key]
    tab = _build_step_tables(score_model, model_args, t_to_sigma, tr_schedule, rot_schedule, tor_schedule, inference_steps, B,
                             temp_sampling, temp_psi, temp_sigma_data, ode)
    if key is not None:
        if len(_STEP_TABLE_CACHE) > 64:
            _STEP_TABLE_CACHE.clear()
        _STEP_TABLE_CACHE[key] = tab
    return tab


def _build_step_tables(score_model, model_args, t_to_sigma, tr_schedule, rot_schedule, tor_schedule, inference_steps, B,
                       temp_sampling, temp_psi, temp_sigma_data, ode=False) -> StepTables:
    ts, tp, td = _three(temp_sampling), _three(temp_psi), _three(temp_sigma_data)
    emb_fn = score_model.timestep_emb_func
    rng = [(model_args.tr_sigma_min, model_args.tr_sigma_max), (model_args.rot_sigma_min, model_args.rot_sigma_max),
           (model_args.tor_sigma_min, model_args.tor_sigma_max)]
    semb, cutoff, trs, rots, tors, coef = [], [], [], [], [], []
    for s in range(inference_steps):
        t = [tr_schedule[s], rot_schedule[s], tor_schedule[s]]
        last = s == inference_steps - 1
        dt = [t[0] if last else tr_schedule[s] - tr_schedule[s + 1], t[1] if last else rot_schedule[s] - rot_schedule[s + 1],
              t[2] if last else tor_schedule[s] - tor_schedule[s + 1]]
        sig = t_to_sigma(t[0], t[1], t[2])                                   # float64 scalars (sampling.py:111)
        # what the model derives from complex_t (float32 tensors, set_time at sampling.py:113)
        ct = [float(x) * torch.ones(B) for x in t]
        m_sig = [torch.as_tensor(x).float() for x in score_model.t_to_sigma(*ct)]
        semb.append(emb_fn(ct[0]))
        cutoff.append(m_sig[0] * 3 + 20 if score_model.dynamic_max_cross else torch.full((B,), score_model.cross_max_distance))
        if score_model.scale_by_sigma:
            trs.append(m_sig[0]); rots.append(so3_score_norm(m_sig[1])); tors.append(torch.sqrt(torus_score_norm(m_sig[2])))
        else:
            trs.append(torch.ones(B)); rots.append(torch.ones(B)); tors.append(torch.ones(B))
        row = []
        for i in range(3):
            g = sig[i] * np.sqrt(2 * np.log(rng[i][1] / rng[i][0]))        # sampling.py:137-140, 167-169
            if ode:
                a, b = 0.5 * g ** 2 * dt[i], 0.0                            # :142-144, 171
            elif ts[i] != 1.0:                                              # low-temperature sampling, :179-192
                sd = np.exp(td[i] * np.log(rng[i][1]) + (1 - td[i]) * np.log(rng[i][0]))
                lam = (sd + sig[i]) / (sd + sig[i] / ts[i])
                a = g ** 2 * dt[i] * (lam + ts[i] * tp[i] / 2)
                b = g * np.sqrt(dt[i] * (1 + tp[i]))
            else:
                a, b = g ** 2 * dt[i], g * np.sqrt(dt[i])                   # :149, 154, 175
            row += [np.float32(a), np.float32(b)]
        coef.append(row)
    return StepTables(inference_steps, torch.stack(semb), torch.stack(cutoff), torch.stack(trs), torch.stack(rots),
                      torch.stack(tors), coef)


def _sample_with_guidance(sm, eng, batch, start_pos, steps, z, tr_schedule, weight, cfg_start, cfg_end, device, copies=True):
    """Classifier-free guidance (utils/sampling.py:119-135): inside [cfg_end, cfg_start] every reverse step evaluates the score
    model twice -- with the latents, and with ``unconditional = 1`` and zeroed latents -- and extrapolates
    ``s + w (s - s_uncond)``.  The two branches live in two libddk contexts over the same pose buffer (the latents enter the
    step-invariant node embedding, so each branch keeps its own batch); the update runs once on the guided scores."""
    from . import engine as _engine
    if getattr(sm, '_engine_uncond', None) is None or sm._engine_uncond.device != eng.device:
        sm._engine_uncond = _engine.Engine(sm.hyper(), {k: v.detach() for k, v in sm.state_dict().items()}, eng.device)
    eng_u = sm._engine_uncond
    ub = batch.shallow_copy()
    for nt in ('ligand', 'receptor'):
        ub[nt].unconditional = torch.ones(batch[nt].num_nodes, 1)
        ub[nt].latent_h = torch.zeros_like(batch[nt].latent_h)
    eng_u.set_batch(ub, assume_copies=copies)
    pos = start_pos.to(device, torch.float32).contiguous().clone()
    dev = lambda t: None if t is None else t.to(device, torch.float32).contiguous()
    for i in range(steps.n_steps):
        args = (steps.semb[i], steps.cutoff[i], steps.tr_sigma[i], steps.rot_scale[i], steps.tor_scale[i])
        tr, rot, tor = eng.score(pos, *args)
        t_tr = float(tr_schedule[i])
        if cfg_end <= t_tr <= cfg_start:
            utr, urot, utor = eng_u.score(pos, *args)
            tr = tr + weight * (tr - utr)
            rot = rot + weight * (rot - urot)
            tor = tor + weight * (tor - utor)
        zi = (None, None, None) if z is None else (dev(z['tr'][i]), dev(z['rot'][i]), dev(z['tor'][i]) if z.get('tor') is not None else None)
        eng.update(pos, tr, rot, tor, zi[0], zi[1], zi[2], steps.coef[i])
    return pos


def sampling(data_list, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, device, t_to_sigma, model_args,
             no_random=False, ode=False, visualization_list=None, confidence_model=None, confidence_data_list=None,
             confidence_model_args=None, batch_size=32, no_final_step_noise=False, use_latent=True,
             gumbel_latent_temperature=0.01, ar_model=None, ar_args=None, temp_sampling=1.0, temp_psi=0.0, temp_sigma_data=0.5,
             classifier_free_guidance_weight=0.0, softmax_latent_temperature=1.0, cfg_start=1.0, cfg_end=0.0,
             compute_ar_accuracy=False, *, noise: Optional[Dict[str, torch.Tensor]] = None, generator=None,
             host_buffers=False):
    N = len(data_list)
    device = torch.device(device)
    sm = _score_model_of(model)
    confidence = [] if confidence_model is not None else None
    conf_loader = iter(DataLoader(confidence_data_list, batch_size=batch_size)) if confidence_data_list is not None else None
    latent = use_latent and getattr(model_args, 'latent_dim', 0) > 0
    cfg_on = classifier_free_guidance_weight != 0.0
    if cfg_on and not (latent and getattr(sm, 'latent_droprate', 0) > 0):
        raise ValueError('classifier-free guidance needs a latent-conditioned score model trained with latent drop-out')
    pose0 = 0
    with torch.no_grad():
        n_batches = (N + batch_size - 1) // batch_size
        fast = not latent and confidence_model is None      # runs of copies: no host-side PyG collation needed
        tor0 = 0
        for batch_id in range(n_batches):
            items = data_list[batch_id * batch_size:(batch_id + 1) * batch_size]
            b = len(items)
            eng = sm.engine(device)
            if fast:
                batch = None
                start_pos = torch.cat([x['ligand'].pos for x in items], dim=0)
                info = eng.set_batch_groups(group_copies(items))
            else:
                batch = Batch.from_data_list(items)
                if latent:
                    if ar_model is None:
                        raise NotImplementedError('oracle latent encoder (TPEncoder) is out of scope; pass ar_model')
                    from .latent import encode_ar_batch
                    pos_keep = batch['ligand'].pos
                    if 'ar_pos' in batch['ligand']:
                        batch['ligand'].pos = batch['ligand'].ar_pos
                    lat_l, lat_r = encode_ar_batch(ar_model, batch, softmax_latent_temperature, device, generator=generator)
                    batch['ligand'].pos = pos_keep
                    batch['ligand'].latent_h, batch['receptor'].latent_h = lat_l, lat_r
                if getattr(sm, 'latent_droprate', 0) > 0:
                    batch['ligand'].unconditional = torch.zeros(batch['ligand'].num_nodes, 1)
                    batch['receptor'].unconditional = torch.zeros(batch['receptor'].num_nodes, 1)
                start_pos = batch['ligand'].pos
                copies = len(group_copies(items)) == 1
                info = eng.set_batch(batch, assume_copies=copies)
            eng._batch_key = None
            steps = build_step_tables(sm, model_args, t_to_sigma, tr_schedule, rot_schedule, tor_schedule, inference_steps, b,
                                      temp_sampling, temp_psi, temp_sigma_data, ode)
            R = info.RB
            if no_random or ode:
                z = None
            elif noise is not None:
                z = {'tr': noise['tr'][:, pose0:pose0 + b], 'rot': noise['rot'][:, pose0:pose0 + b],
                     'tor': noise['tor'][:, tor0:tor0 + R] if R else None}
                if no_final_step_noise:                                     # utils/sampling.py:146-148 holds for supplied noise too
                    z = {k: (None if v is None else v.clone()) for k, v in z.items()}
                    for v in z.values():
                        if v is not None:
                            v[inference_steps - 1] = 0
            else:
                zdev = torch.device('cpu') if host_buffers else device
                z = {'tr': torch.randn(inference_steps, b, 3, device=zdev, generator=generator),
                     'rot': torch.randn(inference_steps, b, 3, device=zdev, generator=generator),
                     'tor': torch.randn(inference_steps, R, device=zdev, generator=generator) if R else None}
                if no_final_step_noise:
                    for v in z.values():
                        if v is not None:
                            v[-1] = 0
            if cfg_on:
                pos = _sample_with_guidance(sm, eng, batch, start_pos, steps, z, tr_schedule, classifier_free_guidance_weight,
                                            cfg_start, cfg_end, device, copies)
            elif host_buffers:
                pos = start_pos.detach().to('cpu', torch.float32).contiguous()
                eng.sample_host(pos, steps, z)
            else:
                pos = start_pos.to(device, torch.float32).contiguous().clone()
                eng.sample(pos, steps, z)
            if batch is not None:
                batch['ligand'].pos = pos
            lp = info.lig_ptr
            for i in range(b):
                data_list[batch_id * batch_size + i]['ligand'].pos = pos[int(lp[i]):int(lp[i + 1])]
                if latent:                                                      # utils/sampling.py:205-222
                    item = data_list[batch_id * batch_size + i]
                    rp = info.rec_ptr
                    lig_lat = batch['ligand'].latent_h[int(lp[i]):int(lp[i + 1])]
                    rec_lat = batch['receptor'].latent_h[int(rp[i]):int(rp[i + 1])]
                    item['ligand'].latent_h = lig_lat
                    centre = item.original_center.detach().cpu() if 'original_center' in item else torch.zeros(1, 3)
                    lat_str, lat_pos = '', []
                    for j in range(model_args.latent_dim):
                        if float(lig_lat[:, j].sum()) == 1:
                            idx = int(torch.argmax(lig_lat[:, j]))
                            lat_str += 'L' + str(idx)
                            lat_pos.append(item['ligand'].pos[idx:idx + 1].detach().cpu() + centre)
                        else:
                            idx = int(torch.argmax(rec_lat[:, j]))
                            lat_str += 'R' + str(idx)
                            lat_pos.append(item['receptor'].pos[idx:idx + 1].detach().cpu() + centre)
                    item.latent_str = lat_str
                    item.latent_pos = torch.cat(lat_pos, dim=0)
            pose0 += b
            tor0 += R
            if visualization_list is not None:
                for idx, vis in enumerate(visualization_list):
                    vis.add((data_list[idx]['ligand'].pos.detach().cpu() + data_list[idx].original_center.detach().cpu()),
                            part=1, order=2)
            if confidence_model is not None:
                from .diffusion_utils import set_time
                if conf_loader is not None:
                    cb = next(conf_loader)
                    cb['ligand'].pos = pos.cpu()
                    cb = cb.to(device)
                    set_time(cb, 0, 0, 0, b, confidence_model_args.all_atoms, device)
                    out = confidence_model(cb)
                else:
                    out = confidence_model(batch.to(device))
                confidence.append(out[0] if type(out) is tuple else out)
    if confidence_model is not None:
        confidence = torch.nan_to_num(torch.cat(confidence, dim=0), nan=-1000)
    return data_list, confidence
