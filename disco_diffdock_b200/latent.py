"""DisCo latent machinery on top of the CUDA score model: the autoregressive latent sampler and the model wrapper.

Mirrors /root/reference/models/model_classes.py:9-49 (``GenericEncoder.encode_ar``), :53-95 (``ModelWrapper``) and
/root/reference/models/pretrained_score_encoder.py:8-89 (``PretrainedScoreEncoder``) for the equivariant latents
(``latent_vocab == 1``) the DisCo-DiffDock-S checkpoints use.  The expensive part -- two ``embed()`` passes of the pretrained
score model at t = 1 per complex -- runs through ``ddk_embed`` (libddk); the two small prediction heads
(Linear 48 -> 128 -> 128 -> 1 with BatchNorm1d) are plain torch modules with the reference's parameter names, so the
shipped AR checkpoints load with ``strict=True``.
"""
from __future__ import annotations

import copy

import torch
from torch import nn

from .diffusion_utils import set_time


def gumbel_softmax(logits, temperature, generator=None):
    """models/layers.py:152-181 (straight-through Gumbel soft-max; at inference only the arg-max matters)."""
    u = torch.rand(logits.shape, generator=generator, device=logits.device if generator is None else generator.device)
    g = -torch.log(-torch.log(u.to(logits.device) + 1e-20) + 1e-20)
    y = torch.softmax((logits + g) / temperature, dim=-1)
    ind = y.argmax(dim=-1, keepdim=True)
    return torch.zeros_like(y).scatter_(-1, ind, 1.0)


class GenericEncoder(nn.Module):
    """model_classes.py:5-49."""

    def encode_ar(self, data, sampling_temperature=1.0, generator=None):
        # assumes graphs of the same complex as input (model_classes.py:10)
        if self.latent_vocab > 1:
            raise NotImplementedError('categorical latents (latent_vocab > 1) are not used by the shipped checkpoints')
        B = data.num_graphs
        dev = data['ligand'].pos.device
        keep = self.apply_gumbel_softmax
        self.apply_gumbel_softmax = False
        n_l, n_r = data['ligand'].pos.shape[0], data['receptor'].pos.shape[0]
        len_lig, len_rec = n_l // B, n_r // B
        latent_l = torch.zeros(n_l, self.input_latent_dim, device=dev)
        latent_r = torch.zeros(n_r, self.input_latent_dim, device=dev)
        try:
            for decoding_idx in range(self.input_latent_dim):
                data['ligand'].input_latent, data['receptor'].input_latent = latent_l.clone(), latent_r.clone()
                data.decoding_idx = torch.zeros(B, dtype=torch.long, device=dev) + decoding_idx
                lat = self.forward(data.shallow_copy() if hasattr(data, 'shallow_copy') else copy.copy(data))
                lat = torch.cat(lat, dim=0)[:, 0, :] * sampling_temperature           # model_classes.py:32
                assert lat.shape == (B, len_lig + len_rec)
                if sampling_temperature >= 100:
                    choice = torch.argmax(lat, 1, keepdim=True)
                else:
                    p = torch.nan_to_num(torch.exp(lat))
                    if generator is not None and generator.device != p.device:
                        choice = torch.multinomial(p.to(generator.device), 1, generator=generator).to(dev)
                    else:
                        choice = torch.multinomial(p, 1, generator=generator)
                c = choice[:, 0]
                rows = torch.arange(B, device=dev)
                in_lig = c < len_lig
                latent_l[(rows * len_lig + c)[in_lig], decoding_idx] = 1
                latent_r[(rows * len_rec + c - len_lig)[~in_lig], decoding_idx] = 1
                self.last_logits = lat
        finally:
            self.apply_gumbel_softmax = keep
        return latent_l, latent_r


class PretrainedScoreEncoder(GenericEncoder):
    """pretrained_score_encoder.py:8-89: latent logits from the scalar node features of a pretrained score model."""

    def __init__(self, pretrained_score_model, ns, latent_dim, latent_vocab, latent_no_batchnorm=False, latent_dropout=0.0,
                 latent_hidden_dim=128, input_latent_dim=0, apply_gumbel_softmax=True):
        super().__init__()
        assert input_latent_dim > 0
        self.ns, self.latent_dim, self.latent_vocab = ns, latent_dim, latent_vocab
        self.latent_temperature = 1.0
        self.input_latent_dim = input_latent_dim
        self.apply_gumbel_softmax = apply_gumbel_softmax
        self.pretrained_score_model = pretrained_score_model
        width = 2 * ns if pretrained_score_model.num_conv_layers >= 3 else ns

        def head():
            bn = (lambda: nn.Identity()) if latent_no_batchnorm else (lambda: nn.BatchNorm1d(latent_hidden_dim))
            return nn.Sequential(nn.Linear(width, latent_hidden_dim), bn(), nn.ReLU(), nn.Dropout(latent_dropout),
                                 nn.Linear(latent_hidden_dim, latent_hidden_dim), bn(), nn.ReLU(), nn.Dropout(latent_dropout),
                                 nn.Linear(latent_hidden_dim, latent_dim))
        self.latent_s_predictor = head()
        self.latent_r_predictor = head()
        self.eval()

    def forward(self, data, generator=None):
        assert self.latent_vocab == 1
        lig, rec = data['ligand'], data['receptor']
        lig.latent_h, rec.latent_h = lig.input_latent, rec.input_latent
        assert torch.all(data.decoding_idx >= 0)
        dev = lig.pos.device
        B = data.num_graphs
        set_time(data, 1, 1, 1, B, False, dev)
        lig.unconditional = torch.ones(lig.pos.shape[0], 1, device=dev)
        rec.unconditional = torch.ones(rec.pos.shape[0], 1, device=dev)
        lig_h, rec_h = self.pretrained_score_model.embed(data)[:2]
        ns = self.ns
        if self.pretrained_score_model.num_conv_layers >= 3:
            s_l = torch.cat([lig_h[:, :ns], lig_h[:, -ns:]], dim=1)
            s_r = torch.cat([rec_h[:, :ns], rec_h[:, -ns:]], dim=1)
        else:
            s_l, s_r = lig_h[:, :ns], rec_h[:, :ns]
        hd = next(self.latent_s_predictor.parameters()).device
        s_l = self.latent_s_predictor(s_l.to(hd))
        s_r = self.latent_r_predictor(s_r.to(hd))
        n_l, n_r = s_l.shape[0] // B, s_r.shape[0] // B
        # per graph: [1, latent_dim, n_l + n_r] (ligand atoms first); graphs are copies of one complex, so equal sizes
        lat = torch.cat([s_l.view(B, n_l, -1), s_r.view(B, n_r, -1)], dim=1).transpose(1, 2)
        if not self.apply_gumbel_softmax:
            return [lat[i:i + 1] for i in range(B)]            # consumed by encode_ar
        hard = gumbel_softmax(lat, self.latent_temperature, generator)
        latent_l = hard[:, :, :n_l].transpose(1, 2).reshape(B * n_l, -1)
        latent_r = hard[:, :, n_l:].transpose(1, 2).reshape(B * n_r, -1)
        return latent_l, latent_r


class ModelWrapper(nn.Module):
    """model_classes.py:53-95, inference side: exposes ``.encoder`` and ``.score_model`` the way utils/sampling.py:63-117
    reaches for them; ``forward`` encodes the latents (no drop-out of latents at inference) and scores."""

    def __init__(self, encoder, score_model, training_latent_temperature=1.0, device=None, latent_droprate=0.0):
        super().__init__()
        self.encoder, self.score_model = encoder, score_model
        self.training_latent_temperature = training_latent_temperature
        self.device = device
        self.latent_droprate = latent_droprate
        self.eval()

    def forward(self, data):
        if self.encoder is not None:
            self.encoder.latent_temperature = self.training_latent_temperature
            latent_h = self.encoder(data)
            if not isinstance(latent_h, tuple):
                raise NotImplementedError('categorical latents (latent_vocab > 1)')
            data['ligand'].latent_h, data['receptor'].latent_h = latent_h
        return self.score_model(data)


def encode_ar_batch(ar_model, batch, temperature, device, generator=None):
    """utils/sampling.py:77-81: latents of a batch of poses of one complex from the AR model."""
    b = batch.to(device) if hasattr(batch, 'to') else batch
    return ar_model.encode_ar(b, temperature, generator=generator)
