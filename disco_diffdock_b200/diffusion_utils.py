"""Host-side schedule / noise-level helpers with the reference's names and semantics
(/root/reference/utils/diffusion_utils.py:12-16, 58-69, 87-117)."""
from __future__ import annotations

import math

import numpy as np
import torch


def t_to_sigma(t_tr, t_rot, t_tor, args):
    """diffusion_utils.py:12-16: geometric interpolation between sigma_min and sigma_max."""
    tr_sigma = args.tr_sigma_min ** (1 - t_tr) * args.tr_sigma_max ** t_tr
    rot_sigma = args.rot_sigma_min ** (1 - t_rot) * args.rot_sigma_max ** t_rot
    tor_sigma = args.tor_sigma_min ** (1 - t_tor) * args.tor_sigma_max ** t_tor
    return tr_sigma, rot_sigma, tor_sigma


def sinusoidal_embedding(timesteps, embedding_dim, max_positions=10000):
    """diffusion_utils.py:58-69."""
    assert timesteps.dim() == 1
    half = embedding_dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32, device=timesteps.device) * -(math.log(max_positions) / (half - 1)))
    arg = timesteps.float()[:, None] * freq[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
    if embedding_dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1))
    return emb


def get_timestep_embedding(embedding_type, embedding_dim, embedding_scale=10000):
    """diffusion_utils.py:87-94 (only the sinusoidal embedding the shipped checkpoints use)."""
    if embedding_type != 'sinusoidal':
        raise NotImplementedError(embedding_type)
    return lambda x: sinusoidal_embedding(embedding_scale * x, embedding_dim)


def get_t_schedule(inference_steps):
    """diffusion_utils.py:97-98."""
    return np.linspace(1, 0, inference_steps + 1)[:-1]


def set_time(complex_graphs, t_tr, t_rot, t_tor, batchsize, all_atoms, device):
    """diffusion_utils.py:101-117."""
    for nt in ('ligand', 'receptor') + (('atom',) if all_atoms else ()):
        n = complex_graphs[nt].num_nodes
        complex_graphs[nt].node_t = {'tr': t_tr * torch.ones(n).to(device), 'rot': t_rot * torch.ones(n).to(device),
                                     'tor': t_tor * torch.ones(n).to(device)}
    complex_graphs.complex_t = {'tr': t_tr * torch.ones(batchsize).to(device), 'rot': t_rot * torch.ones(batchsize).to(device),
                                'tor': t_tor * torch.ones(batchsize).to(device)}
