"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the REFERENCE's own code.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
Each case = seeded synthetic inputs (tests/helpers.py) + seeded fresh-initialised weights pushed through the
reference's unmodified ``TensorProductScoreModel.forward`` (models/score_model.py:259) or ``sampling()``
(utils/sampling.py:49) behind the leaf-op shims, with ``torch.normal`` replaced by pre-drawn noise.  The
files hold the reference outputs plus fingerprints of the inputs / weights so a replay on another box can
tell "inputs drifted" from "math differs".
"""
from __future__ import annotations

import copy
import os
import sys
from functools import partial
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from disco_diffdock_b200 import data as ddata  # noqa: E402
from disco_diffdock_b200.synthetic import as_loader_item  # noqa: E402
from oracle import ref_loader, restate  # noqa: E402
from tests import helpers  # noqa: E402

CASES = {
    # name: (model seed, gain, complex seed, n_lig, n_rec, B, steps, temps, latent_dim)
    'forward_small': dict(mseed=0, gain=1.0, cseed=3, n_lig=20, n_rec=50, B=3, t=0.6, latent=0),
    'forward_latent': dict(mseed=0, gain=1.0, cseed=3, n_lig=20, n_rec=50, B=3, t=0.35, latent=2),
    'forward_cfg1': dict(mseed=2, gain=1.0, cseed=11, n_lig=60, n_rec=300, B=1, t=0.9, latent=0),
    'sample_small': dict(mseed=1, gain=5.0, cseed=4, n_lig=14, n_rec=40, B=2, steps=8, temps=True, latent=0),
    'sample_mid': dict(mseed=1, gain=5.0, cseed=6, n_lig=30, n_rec=80, B=3, steps=20, temps=True, latent=0),
    # cseed 12: a WELL-CONDITIONED trajectory.  The model is discontinuous in the positions (radius-graph edges switch on
    # and off), so a 20-step trajectory can sit next to an edge flip: with cseed 11 the oracle itself lands 3e-3 A away
    # when its start pose is perturbed by 1e-7 A (below fp32 resolution), with cseed 12 / 15 four 2e-6 A perturbations
    # all stay within 2e-4 A (tools/golden_conditioning.py).  A golden next to a flip tests rounding luck, not parity.
    'sample_cfg1': dict(mseed=2, gain=5.0, cseed=12, n_lig=60, n_rec=300, B=1, steps=20, temps=True, latent=0),
}


def fingerprint(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def forward_inputs(c):
    m, sd, cfg = helpers.make_model(c['mseed'], latent_dim=c['latent'], latent_droprate=0.1 if c['latent'] else 0.0,
                                    gain=c['gain'])
    _, lst = helpers.make_pose_batch(c['cseed'], c['n_lig'], c['n_rec'], c['B'])
    batch = ddata.Batch.from_data_list(lst)
    restate.set_time(batch, c['t'], c['t'], c['t'], c['B'])
    if c['latent']:
        gen = torch.Generator().manual_seed(5)
        for nt in ('ligand', 'receptor'):
            batch[nt].latent_h = (torch.rand(batch[nt].num_nodes, c['latent'], generator=gen) > 0.9).float()
            batch[nt].unconditional = (torch.rand(batch[nt].num_nodes, 1, generator=gen) > 0.5).float()
    return m, sd, cfg, batch


def sample_inputs(c):
    m, sd, cfg = helpers.make_model(c['mseed'], gain=c['gain'])
    g, lst = helpers.make_pose_batch(c['cseed'], c['n_lig'], c['n_rec'], c['B'])
    R = g['ligand'].mask_rotate.shape[0]
    noise = helpers.draw_noise(7, c['steps'], c['B'], R)
    sched = np.linspace(1, 0, c['steps'] + 1)[:-1]
    temps = helpers.README_TEMPS if c['temps'] else {}
    return m, sd, cfg, lst, noise, sched, temps


def main():
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    mods = ref_loader.modules()
    only = set(sys.argv[1:])
    for name, c in CASES.items():
        if only and name not in only:
            continue
        if name.startswith('forward'):
            m, sd, cfg, batch = forward_inputs(c)
            ref_model, _ = ref_loader.build_reference_model(cfg, sd)
            with torch.no_grad():
                tr, rot, tor = ref_model(copy.deepcopy(batch))
                lig_h, rec_h = ref_model.embed(copy.deepcopy(batch))[:2]
            np.savez_compressed(os.path.join(out_dir, name + '.npz'), tr=tr.numpy(), rot=rot.numpy(), tor=tor.numpy(),
                                lig_h=lig_h.numpy(), rec_h=rec_h.numpy(), weights_fp=fingerprint(sd),
                                pos_fp=float(batch['ligand'].pos.double().abs().sum()),
                                rec_fp=float(batch['receptor'].x.double().abs().sum()))
        else:
            m, sd, cfg, lst, noise, sched, temps = sample_inputs(c)
            ref_model, args = ref_loader.build_reference_model(cfg, sd)
            t2s = partial(mods.diffusion_utils.t_to_sigma, args=args)
            ref_list = [as_loader_item(x) for x in copy.deepcopy(lst)]
            with ref_loader.InjectedNormal(noise, c['steps']), torch.no_grad():
                out_list, _ = mods.sampling.sampling(ref_list, SimpleNamespace(score_model=ref_model), c['steps'],
                                                     sched, sched, sched, torch.device('cpu'), t2s, args,
                                                     batch_size=c['B'], no_final_step_noise=False, **temps)
            pos = torch.cat([x['ligand'].pos for x in out_list])
            start = torch.cat([x['ligand'].pos for x in lst])
            np.savez_compressed(os.path.join(out_dir, name + '.npz'), pos=pos.numpy(), weights_fp=fingerprint(sd),
                                pos_fp=float(start.double().abs().sum()),
                                rec_fp=float(lst[0]['receptor'].x.double().abs().sum()))
        print('wrote', name)
    if not only or 'sample_disco' in only:
        write_disco(out_dir, mods)
    for name in helpers.PRE_CASES:
        if not only or name in only:
            write_pretrained(out_dir, mods, name)
    if not only or 'host_fixtures' in only:
        write_host_fixtures(out_dir, mods)


def write_host_fixtures(out_dir, mods):
    """randomize_position (utils/sampling.py:12-46) and encode_ar with multinomial decoding (models/model_classes.py:9-49) run by
    the REFERENCE's functions on the seeded inputs of tests/test_host_parity.py."""
    import contextlib, io
    from tests import test_host_parity as T
    with contextlib.redirect_stdout(io.StringIO()):
        from models.pretrained_score_encoder import PretrainedScoreEncoder as RefEnc
    out = {}
    for key, nr in (('random', False), ('no_random', True)):
        b = T._rp_inputs(); T._seed(); mods.sampling.randomize_position(b, False, nr, 19.0)
        out[key] = torch.stack([x['ligand'].pos for x in b]).numpy()
    np.savez_compressed(os.path.join(out_dir, 'randomize_position.npz'), **out)
    out = {}
    for temp in (1.0, 3.0):
        lst, heads, B = T._ar_case()
        ref = RefEnc(pretrained_score_model=T._StubScore(), ns=16, latent_dim=1, latent_vocab=1, input_latent_dim=2)
        ref.load_state_dict(heads, strict=True)
        ref.eval()
        torch.manual_seed(123)
        with torch.no_grad():
            l, r = ref.encode_ar(ddata.Batch.from_data_list(lst), temp)
        out[f'l_{temp}'] = l.numpy(); out[f'r_{temp}'] = r.numpy()
    np.savez_compressed(os.path.join(out_dir, 'encode_ar_multinomial.npz'), **out)
    print('wrote host fixtures (reference functions)')


class _Recorder:
    """Stands where ``model.score_model`` is called in the reference's sampling() (utils/sampling.py:116-117): records the pose
    BEFORE every reverse step and the scores the reference's own forward returns for it."""

    def __init__(self, ref_model):
        self.ref_model, self.pos, self.scores = ref_model, [], []

    def __call__(self, batch):
        self.pos.append(batch['ligand'].pos.detach().clone())
        out = self.ref_model(batch)
        self.scores.append([o.detach().clone() for o in out[:3]])
        return out


def write_pretrained(out_dir, mods, name):
    """Goldens in the PRETRAINED-weight regime: the reference's unmodified forward / sampling() with the shipped checkpoint
    (evaluate.py:160-174 loads the same file) on a calibrated synthetic complex (tests/helpers.py PRE_CASES)."""
    c = helpers.PRE_CASES[name]
    sd, cfg = helpers.load_checkpoint(c['ckpt'])
    ref_model, args = ref_loader.build_reference_model(cfg, sd)
    tables = ref_loader.load_tables()
    out = dict(weights_fp=fingerprint({k: v.float() for k, v in sd.items()}))
    if 'ts' in c:
        for i, t in enumerate(c['ts']):
            batch = helpers.pre_forward_batch(c, t)
            with torch.no_grad():
                tr, rot, tor = ref_model(copy.deepcopy(batch))
                lig_h, rec_h = ref_model.embed(copy.deepcopy(batch))[:2]
            out.update({f'tr{i}': tr.numpy(), f'rot{i}': rot.numpy(), f'tor{i}': tor.numpy(), f'lig_h{i}': lig_h.numpy(),
                        f'rec_h{i}': rec_h.numpy(), f'pos{i}': batch['ligand'].pos.numpy()})
        out['rec_fp'] = float(batch['receptor'].x.double().abs().sum())
    else:
        g, lst, noise, sched, kw = helpers.pre_traj_inputs(c)
        t2s = partial(mods.diffusion_utils.t_to_sigma, args=args)
        ref_list = [as_loader_item(x) for x in copy.deepcopy(lst)]
        rec = _Recorder(ref_model)
        with ref_loader.InjectedNormal(noise, c['steps']), torch.no_grad():
            out_list, _ = mods.sampling.sampling(ref_list, SimpleNamespace(score_model=rec), c['steps'], sched, sched, sched,
                                                 torch.device('cpu'), t2s, args, batch_size=c['B'], no_final_step_noise=False, **kw)
        final = torch.cat([x['ligand'].pos for x in out_list])
        out.update(start=torch.cat([x['ligand'].pos for x in lst]).numpy(), final=final.numpy(),
                   pos_steps=torch.stack(rec.pos + [final]).numpy(),
                   tr=torch.stack([s[0] for s in rec.scores]).numpy(), rot=torch.stack([s[1] for s in rec.scores]).numpy(),
                   tor=torch.stack([s[2] for s in rec.scores]).numpy(),
                   rec_fp=float(lst[0]['receptor'].x.double().abs().sum()))
        # how well conditioned is this trajectory?  the oracle against itself from start poses perturbed by 2e-6 A
        def oracle_run(eps, seed):
            l2 = copy.deepcopy(lst)
            if eps:
                gg = torch.Generator().manual_seed(seed)
                for x in l2:
                    x['ligand'].pos = x['ligand'].pos + eps * torch.randn(x['ligand'].pos.shape, generator=gg)
            with torch.no_grad():
                return restate.sample(sd, cfg, ddata.Batch.from_data_list(l2), tables, sched, noise, inference_steps=c['steps'], **kw).clone()
        base = oracle_run(0, 0)
        out['oracle_vs_reference_rmsd'] = float(helpers.rmsd_per_pose(base, final, c['B']).max())
        out['oracle_spread_2e-6'] = np.array([float(helpers.rmsd_per_pose(base, oracle_run(2e-6, s), c['B']).max()) for s in range(3)])
        print(name, 'oracle vs reference', out['oracle_vs_reference_rmsd'], 'spread', out['oracle_spread_2e-6'])
    np.savez_compressed(os.path.join(out_dir, name + '.npz'), **out)
    print('wrote', name)


def write_disco(out_dir, mods):
    """DisCo path through the reference's own code: PretrainedScoreEncoder.encode_ar (models/model_classes.py:9-49,
    models/pretrained_score_encoder.py:46-89; arg-max decoding) feeding utils/sampling.sampling with classifier-free
    guidance (utils/sampling.py:119-135)."""
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        from models.pretrained_score_encoder import PretrainedScoreEncoder
    c = helpers.DISCO_CASE
    m, sd, cfg, lst, noise, sched, heads = helpers.disco_inputs(c)
    ref_model, args = ref_loader.build_reference_model(cfg, sd)
    ar = PretrainedScoreEncoder(pretrained_score_model=ref_model, ns=cfg.ns, latent_dim=1, latent_vocab=1, latent_no_batchnorm=False,
                                latent_dropout=0.0, latent_hidden_dim=128, input_latent_dim=cfg.latent_dim, apply_gumbel_softmax=True)
    missing = ar.load_state_dict(heads, strict=False)
    assert all(k.startswith('pretrained_score_model.') for k in missing.missing_keys) and not missing.unexpected_keys
    ar.eval()
    t2s = partial(mods.diffusion_utils.t_to_sigma, args=args)
    ref_list = [as_loader_item(x) for x in copy.deepcopy(lst)]
    for a, b in zip(ref_list, lst):
        a['ligand'].ar_pos = b['ligand'].ar_pos.clone()
    with ref_loader.InjectedNormal(noise, c['steps']), torch.no_grad():
        # the latents the AR model decodes for this batch (same call sampling() makes)
        probe = ddata.Batch.from_data_list(copy.deepcopy(ref_list))
        probe['ligand'].pos = probe['ligand'].ar_pos
        lat_l, lat_r = ar.encode_ar(probe, c['softmax_latent_temperature'])
        out_list, _ = mods.sampling.sampling(ref_list, SimpleNamespace(score_model=ref_model, encoder=None), c['steps'], sched, sched,
                                             sched, torch.device('cpu'), t2s, args, batch_size=c['B'], no_final_step_noise=False,
                                             ar_model=ar, classifier_free_guidance_weight=c['cfg_weight'], cfg_start=c['cfg_start'],
                                             cfg_end=c['cfg_end'], softmax_latent_temperature=c['softmax_latent_temperature'],
                                             **helpers.README_TEMPS)
    pos = torch.cat([x['ligand'].pos for x in out_list])
    np.savez_compressed(os.path.join(out_dir, 'sample_disco.npz'), pos=pos.numpy(), lat_l=lat_l.numpy(), lat_r=lat_r.numpy(),
                        latent_str=np.array([x.latent_str for x in out_list]), weights_fp=fingerprint(sd),
                        heads_fp=fingerprint({k: v.float() for k, v in heads.items()}))
    print('wrote sample_disco', [x.latent_str for x in out_list])


if __name__ == '__main__':
    main()
