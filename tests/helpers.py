"""Shared builders for the test-suite: seeded weights, synthetic batches, pre-drawn noise."""
from __future__ import annotations

import copy
import os

import numpy as np
import torch

from disco_diffdock_b200 import data as ddata
from disco_diffdock_b200 import synthetic
from disco_diffdock_b200.score_model import TensorProductScoreModel
from oracle import restate
from functools import partial
from disco_diffdock_b200 import diffusion_utils as du

README_TEMPS = dict(  # /root/reference/README.md:15 (DiffDock-S inference command)
    temp_sampling=(1.886430780895051, 5.659562317960644, 2.8888668488630156),
    temp_psi=(0.07085125444659945, 2.686505606141324, 4.089493860493927),
    temp_sigma_data=(0.3617563913086843, 0.7437588205919711, 0.08897393057297842))


_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKPOINTS = {   # the reference's shipped weights (evaluate.py:160-181 loads exactly these files)
    'diffdockS': ('diffdockS_score_model', 'best_ema_inference_epoch_model.pt', dict(latent_dim=0, latent_droprate=0.0)),
    'disco': ('disco_diffdockS_score_model', 'best_ema_inference_epoch_model.pt', dict(latent_dim=2, latent_droprate=0.1)),
    'disco_ar': ('disco_diffdockS_ar_model', 'best_model_loss.pt', dict(latent_dim=2, latent_droprate=0.1)),
}


def checkpoint_path(name):
    """baseline/_ref/workdir (git-ignored copy that travels to the GPU box, tools/fetch_ref.py) or the mounted reference."""
    d, f, _ = CHECKPOINTS[name]
    for root in (os.path.join(_ROOT, 'baseline', '_ref', 'workdir'), os.path.join(os.environ.get('DDK_REFERENCE', '/root/reference'), 'workdir')):
        p = os.path.join(root, d, f)
        if os.path.exists(p):
            return p
    return None


def load_checkpoint(name):
    """(state_dict as fp32 tensors, oracle config) of a shipped checkpoint, or None when it is not on this box."""
    p = checkpoint_path(name)
    if p is None:
        return None
    sd = torch.load(p, map_location='cpu', weights_only=True)
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}
    return sd, restate.default_config(**CHECKPOINTS[name][2])


def make_checkpoint_model(name, device='cpu'):
    """The drop-in TensorProductScoreModel with a shipped checkpoint loaded strict=True (utils/model_utils.py:24-68)."""
    sd, cfg = load_checkpoint(name)
    if name == 'disco_ar':
        sd = {k[len('pretrained_score_model.'):]: v for k, v in sd.items() if k.startswith('pretrained_score_model.')}
    m = TensorProductScoreModel(partial(du.t_to_sigma, args=cfg), device, du.get_timestep_embedding('sinusoidal', 32, 1000),
                                sh_lmax=1, ns=24, nv=6, num_conv_layers=5, lig_max_radius=5.0, cross_max_distance=80.0,
                                dynamic_max_cross=True, dropout=0.1, lm_embedding_type='esm', latent_dim=cfg.latent_dim,
                                latent_vocab=1, latent_droprate=cfg.latent_droprate)
    m.load_state_dict(sd, strict=True)
    return m, sd, cfg


def make_model(seed=0, latent_dim=0, latent_droprate=0.0, device='cpu', randomize_bn=True, gain=1.0, num_conv_layers=5):
    """Fresh default-initialised DiffDock-S architecture (model_parameters.yml of the shipped checkpoint)."""
    torch.manual_seed(seed)
    cfg = restate.default_config(latent_dim=latent_dim, latent_droprate=latent_droprate, num_conv_layers=num_conv_layers)
    m = TensorProductScoreModel(partial(du.t_to_sigma, args=cfg), device,
                                du.get_timestep_embedding('sinusoidal', 32, 1000), sh_lmax=1, ns=24, nv=6, num_conv_layers=num_conv_layers, lig_max_radius=5.0,
                                cross_max_distance=80.0, dynamic_max_cross=True, dropout=0.1, lm_embedding_type='esm',
                                latent_dim=latent_dim, latent_vocab=1, latent_droprate=latent_droprate)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    if randomize_bn:   # exercise the batch-norm affine instead of the identity it is at initialisation
        for k in sd:
            if 'batch_norm.running_var' in k:
                sd[k] = torch.rand(sd[k].shape, generator=g) * 1.5 + 0.5
            elif 'batch_norm.running_mean' in k or 'batch_norm.bias' in k:
                sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
            elif 'batch_norm.weight' in k:
                sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
    if latent_droprate > 0:
        for k in sd:
            if 'unconditional_embedding' in k:
                sd[k] = torch.randn(sd[k].shape, generator=g) * 0.3
    if gain != 1.0:    # larger scores -> the drift term matters in the 20-step parity runs
        for k in ('tr_final_layer.3.weight', 'rot_final_layer.3.weight', 'tor_final_layer.3.weight'):
            sd[k] = sd[k] * gain
    m.load_state_dict(sd)
    return m, sd, cfg


def make_pose_batch(seed, n_lig, n_rec, B, jitter=True):
    """B copies of one synthetic complex with perturbed ligand poses (as sampling() sees them)."""
    g = synthetic.make_complex(seed, n_lig, n_rec)
    gen = torch.Generator().manual_seed(seed + 100)
    lst = [copy.deepcopy(g) for _ in range(B)]
    if jitter:
        for x in lst:
            x['ligand'].pos = (x['ligand'].pos + torch.randn(1, 3, generator=gen) * 3
                               + torch.randn(n_lig, 3, generator=gen) * 0.2)
    return g, lst


def draw_noise(seed, steps, B, R, no_final_step_noise=True):
    g = torch.Generator().manual_seed(seed)
    z = {'tr': torch.randn(steps, B, 3, generator=g), 'rot': torch.randn(steps, B, 3, generator=g),
         'tor': torch.randn(steps, B * R, generator=g)}
    if no_final_step_noise:
        for k in z:
            z[k][-1] = 0
    return z


# ---- pretrained-regime cases (SURVEY finding 7, App. C): shipped checkpoint, synthetic complex with the ESM scale calibrated
# so that the activations stay in the trained range.  With these weights the reverse process is far more sensitive than with
# fresh weights (the oracle's own spread under a 2e-6 A start perturbation is stored next to each golden), so besides one
# well-conditioned free-running ODE trajectory the goldens hold the reference's pose BEFORE every step and its scores AT every
# step: the per-step ("teacher-forced") comparison is well conditioned whatever the trajectory does.
PRE_ESM_SCALE = 0.08
PRE_CASES = {
    'pre_forward': dict(ckpt='diffdockS', cseed=15, n_lig=60, n_rec=300, B=2, ts=(0.95, 0.5, 0.1)),
    'pre_forward_disco': dict(ckpt='disco', cseed=15, n_lig=60, n_rec=300, B=2, ts=(0.7, 0.2)),
    'pre_traj_ode': dict(ckpt='diffdockS', cseed=22, n_lig=60, n_rec=300, B=1, steps=20, ode=True, noise_scale=0.0),
    'pre_traj_temps': dict(ckpt='diffdockS', cseed=15, n_lig=60, n_rec=300, B=1, steps=20, ode=False, noise_scale=0.1),
}


def pre_complex(c):
    return synthetic.make_complex(c['cseed'], c['n_lig'], c['n_rec'], esm_scale=PRE_ESM_SCALE)


def pre_forward_batch(c, t):
    """B in-pocket jittered poses of the case's complex at time t (+ seeded latents for the DisCo checkpoint)."""
    g = pre_complex(c)
    gen = torch.Generator().manual_seed(c['cseed'] + 100)
    lst = [copy.deepcopy(g) for _ in range(c['B'])]
    for x in lst:
        x['ligand'].pos = (x['ligand'].pos + torch.randn(1, 3, generator=gen) * 3 + torch.randn(c['n_lig'], 3, generator=gen) * 0.2)
    batch = ddata.Batch.from_data_list(lst)
    restate.set_time(batch, t, t, t, c['B'])
    if CHECKPOINTS[c['ckpt']][2]['latent_dim']:
        # what sampling() feeds the DisCo model (utils/sampling.py:77-135): per graph and latent dimension exactly one node of
        # ligand + receptor carries a 1 (encode_ar); the LAST pose is the classifier-free-guidance branch instead
        # (unconditional = 1, latents zeroed)
        nl, nr, B = c['n_lig'], c['n_rec'], c['B']
        lat_l, lat_r = torch.zeros(B * nl, 2), torch.zeros(B * nr, 2)
        unc_l, unc_r = torch.zeros(B * nl, 1), torch.zeros(B * nr, 1)
        for b in range(B):
            if b == B - 1 and B > 1:
                unc_l[b * nl:(b + 1) * nl] = 1.0
                unc_r[b * nr:(b + 1) * nr] = 1.0
                continue
            for j in range(2):
                k = int(torch.randint(0, nl + nr, (1,), generator=gen))
                if k < nl:
                    lat_l[b * nl + k, j] = 1.0
                else:
                    lat_r[b * nr + k - nl, j] = 1.0
        batch['ligand'].latent_h, batch['receptor'].latent_h = lat_l, lat_r
        batch['ligand'].unconditional, batch['receptor'].unconditional = unc_l, unc_r
    return batch


def pre_traj_inputs(c):
    """Start poses (randomize_position(no_random=True): uniform torsions, random rotation, centred on the protein,
    utils/sampling.py:12-46), pre-drawn noise and schedule of a pretrained trajectory case."""
    import numpy as _np
    from disco_diffdock_b200 import sampling as dsampling
    g = pre_complex(c)
    lst = [copy.deepcopy(g) for _ in range(c['B'])]
    _np.random.seed(c['cseed']); torch.manual_seed(c['cseed'])
    dsampling.randomize_position(lst, False, True, 19.0, unbatched=True)
    R = g['ligand'].mask_rotate.shape[0]
    noise = draw_noise(7, c['steps'], c['B'], R)
    for k in noise:
        noise[k] = noise[k] * c['noise_scale']
    sched = np.linspace(1, 0, c['steps'] + 1)[:-1]
    kw = dict(ode=True) if c['ode'] else dict(README_TEMPS)
    return g, lst, noise, sched, kw


def rmsd_per_pose(a, b, B):
    a, b = a.reshape(B, -1, 3).double(), b.reshape(B, -1, 3).double()
    return ((a - b) ** 2).sum(-1).mean(-1).sqrt()


def make_ar_heads(seed, ns=24, hidden=128, latent_dim=1):
    """Seeded parameters of the two latent prediction heads of PretrainedScoreEncoder
    (pretrained_score_encoder.py:23-44), batch-norm statistics randomised so the eval-mode affine is exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for head in ('latent_s_predictor', 'latent_r_predictor'):
        dims = [(2 * ns, hidden), (hidden, hidden), (hidden, latent_dim)]
        for li, (i, o) in zip((0, 4, 8), dims):
            sd[f'{head}.{li}.weight'] = torch.randn(o, i, generator=g) / np.sqrt(i)
            sd[f'{head}.{li}.bias'] = torch.randn(o, generator=g) * 0.1
        for bi in (1, 5):
            sd[f'{head}.{bi}.weight'] = torch.rand(hidden, generator=g) + 0.5
            sd[f'{head}.{bi}.bias'] = torch.randn(hidden, generator=g) * 0.1
            sd[f'{head}.{bi}.running_mean'] = torch.randn(hidden, generator=g) * 0.1
            sd[f'{head}.{bi}.running_var'] = torch.rand(hidden, generator=g) + 0.5
            sd[f'{head}.{bi}.num_batches_tracked'] = torch.tensor(10)
    return sd


DISCO_CASE = dict(mseed=3, gain=5.0, cseed=9, n_lig=16, n_rec=40, B=3, steps=8, latent=2, ar_seed=21, cfg_weight=0.6,
                  cfg_start=1.0, cfg_end=0.3, softmax_latent_temperature=100.0)


def disco_inputs(c=DISCO_CASE):
    """Config 4 in miniature: latent-conditioned score model + AR latent sampler + classifier-free guidance."""
    m, sd, cfg = make_model(c['mseed'], latent_dim=c['latent'], latent_droprate=0.1, gain=c['gain'])
    g, lst = make_pose_batch(c['cseed'], c['n_lig'], c['n_rec'], c['B'])
    for x in lst:
        x['ligand'].ar_pos = g['ligand'].pos.clone()          # utils/sampling.py:78-80
    R = g['ligand'].mask_rotate.shape[0]
    noise = draw_noise(11, c['steps'], c['B'], R)
    sched = np.linspace(1, 0, c['steps'] + 1)[:-1]
    return m, sd, cfg, lst, noise, sched, make_ar_heads(c['ar_seed'])
