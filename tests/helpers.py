"""Shared builders for the test-suite: seeded weights, synthetic batches, pre-drawn noise."""
from __future__ import annotations

import copy

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
