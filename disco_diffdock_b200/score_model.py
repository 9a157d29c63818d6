"""Drop-in for ``models/score_model.py:TensorProductScoreModel`` of the reference, computed on B200 by ``libddk``.

Same constructor arguments, ``state_dict`` keys and call surface as the reference
(``/root/reference/models/score_model.py:14-22, 169, 259``; instantiated by ``utils/model_utils.py:24-68``):

    tr_pred[B,3], rot_pred[B,3], tor_pred[sum R] = model(batch)
    lig_node_attr, rec_node_attr, tr_sigma, rot_sigma, tor_sigma = model.embed(batch)

but there is no torch math here: ``forward`` hands raw device pointers to the C ABI in ``include/ddk.h``
(hand-written sm_100a kernels, ``csrc/``) and fails loudly when the CUDA library is missing.  Supported
architecture = what the reference's coarse-grained checkpoints use: ``sh_lmax=1``, no second-order
representation, ``ns=24, nv=6``, new atom encoder, ESM receptor embeddings, optional equivariant latents
(``latent_vocab==1``).  Anything else raises ``NotImplementedError`` at construction.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
from torch import nn

from . import engine
from .params import (AtomEncoderParams, ConvLayerParams, HeadConvParams, LIG_FEATURE_DIMS, REC_FEATURE_DIMS,
                     Smearing, TensorProductBuffers, edge_mlp, irrep_level_dims, tp_weight_numel)


class TensorProductScoreModel(nn.Module):
    def __init__(self, t_to_sigma, device, timestep_emb_func, in_lig_edge_features=4, sigma_embed_dim=32, sh_lmax=2,
                 ns=16, nv=4, num_conv_layers=2, lig_max_radius=5, rec_max_radius=30, cross_max_distance=250,
                 center_max_distance=30, distance_embed_dim=32, cross_distance_embed_dim=32, no_torsion=False,
                 scale_by_sigma=True, use_second_order_repr=False, batch_norm=True,
                 dynamic_max_cross=False, dropout=0.0, lm_embedding_type=None, confidence_mode=False,
                 confidence_dropout=0, confidence_no_batchnorm=False, num_confidence_outputs=1,
                 use_old_atom_encoder=False, latent_dim=0, latent_vocab=32, latent_cross_attention=False,
                 new_cross_attention=False, cross_attention_heads=2, cross_attention_dim=16, latent_droprate=0.0):
        super().__init__()
        unsupported = []
        if sh_lmax != 1 or use_second_order_repr:
            unsupported.append('sh_lmax must be 1 without second-order representation')
        if (ns, nv) != (24, 6):
            unsupported.append('kernels are compiled for ns=24, nv=6')
        if sigma_embed_dim != 32 or distance_embed_dim != 32 or cross_distance_embed_dim != 32:
            unsupported.append('embedding widths must be 32')
        if not batch_norm or confidence_mode or use_old_atom_encoder or latent_cross_attention:
            unsupported.append('batch_norm=True, score mode, new atom encoder, no latent cross attention')
        if lm_embedding_type != 'esm':
            unsupported.append("lm_embedding_type must be 'esm'")
        if latent_dim > 0 and latent_vocab != 1:
            unsupported.append('only equivariant latents (latent_vocab == 1)')
        if num_conv_layers < 3:
            unsupported.append('num_conv_layers >= 3 (the score heads consume the full 0e+1o+1e+0o representation)')
        if in_lig_edge_features != 4:
            unsupported.append('in_lig_edge_features must be 4')
        if unsupported:
            raise NotImplementedError('disco_diffdock_b200: ' + '; '.join(unsupported))

        self.t_to_sigma = t_to_sigma
        self.timestep_emb_func = timestep_emb_func
        self.device = device
        self.ns, self.nv = ns, nv
        self.num_conv_layers = num_conv_layers
        self.no_torsion = no_torsion
        self.scale_by_sigma = scale_by_sigma
        self.dynamic_max_cross = dynamic_max_cross
        self.lig_max_radius = float(lig_max_radius)
        self.rec_max_radius = float(rec_max_radius)
        self.cross_max_distance = float(cross_max_distance)
        self.center_max_distance = float(center_max_distance)
        self.sigma_embed_dim = sigma_embed_dim
        self.latent_dim = latent_dim
        self.latent_vocab = latent_vocab
        self.latent_droprate = latent_droprate
        self.confidence_mode = False
        lat_node = latent_dim * latent_vocab
        lat_edge = latent_dim * max(latent_vocab, 2)

        self.lig_node_embedding = AtomEncoderParams(ns, LIG_FEATURE_DIMS, sigma_embed_dim + lat_node)
        self.lig_edge_embedding = edge_mlp(in_lig_edge_features + sigma_embed_dim + distance_embed_dim + lat_edge, ns, dropout)
        self.rec_node_embedding = AtomEncoderParams(ns, REC_FEATURE_DIMS, sigma_embed_dim + 1280 + lat_node)
        self.rec_edge_embedding = edge_mlp(sigma_embed_dim + distance_embed_dim + lat_edge, ns, dropout)
        self.cross_edge_embedding = edge_mlp(sigma_embed_dim + cross_distance_embed_dim + lat_edge, ns, dropout)
        if latent_droprate > 0:
            for name in ('lig_node', 'rec_node', 'lig_edge', 'rec_edge', 'cross_edge'):
                setattr(self, f'{name}_unconditional_embedding', nn.Parameter(torch.zeros(1, ns)))
        self.lig_distance_expansion = Smearing(self.lig_max_radius, distance_embed_dim)
        self.rec_distance_expansion = Smearing(self.rec_max_radius, distance_embed_dim)
        self.cross_distance_expansion = Smearing(self.cross_max_distance, cross_distance_embed_dim)
        self.conv_layers = nn.ModuleList([ConvLayerParams(ns, nv, l, dropout) for l in range(num_conv_layers)])

        self.center_distance_expansion = Smearing(self.center_max_distance, distance_embed_dim)
        self.center_edge_embedding = edge_mlp(distance_embed_dim + sigma_embed_dim, ns, dropout)
        top = irrep_level_dims(ns, nv, num_conv_layers)
        fc_numel = 2 * (top['0e'] + top['1o'] + top['1o'] + top['1e'] + top['1e'] + top['0o'])
        self.final_conv = HeadConvParams(2 * ns, fc_numel, 12, 4, 0, dropout, w3j=('_w3j_1_1_1',))
        self.tr_final_layer = nn.Sequential(nn.Linear(1 + sigma_embed_dim, ns), nn.Dropout(dropout), nn.ReLU(), nn.Linear(ns, 1))
        self.rot_final_layer = nn.Sequential(nn.Linear(1 + sigma_embed_dim, ns), nn.Dropout(dropout), nn.ReLU(), nn.Linear(ns, 1))
        if not no_torsion:
            self.final_edge_embedding = edge_mlp(distance_embed_dim, ns, dropout)
            self.final_tp_tor = TensorProductBuffers(20, w3j=('_w3j_0_2_2', '_w3j_1_2_1', '_w3j_1_2_2', '_w3j_1_2_3'))
            self.tor_bond_conv = HeadConvParams(3 * ns, (top['1o'] + top['1e']) * ns, 2 * ns, 2 * ns, ns, dropout)
            self.tor_final_layer = nn.Sequential(nn.Linear(2 * ns, ns, bias=False), nn.Tanh(), nn.Dropout(dropout),
                                                 nn.Linear(ns, 1, bias=False))
        self._engines = {}            # one libddk context per device
        self._engines_key = None
        self._engine_uncond = None
        self.eval()

    # ------------------------------------------------------------------------------------------ engine
    def hyper(self):
        return SimpleNamespace(ns=self.ns, nv=self.nv, num_conv_layers=self.num_conv_layers,
                               lig_max_radius=self.lig_max_radius, rec_max_radius=self.rec_max_radius,
                               cross_max_distance=self.cross_max_distance, center_max_distance=self.center_max_distance,
                               dynamic_max_cross=self.dynamic_max_cross, scale_by_sigma=self.scale_by_sigma,
                               no_torsion=self.no_torsion, latent_dim=self.latent_dim,
                               latent_droprate=self.latent_droprate)

    def _weights_key(self):
        """Changes whenever any parameter / buffer is replaced or written in place (load_state_dict of this module or of a
        parent wrapper, ``.to()``, dist.broadcast, optimiser-style in-place updates)."""
        return tuple((v.data_ptr(), v._version) for v in self.state_dict(keep_vars=True).values())

    def engine(self, device=None) -> 'engine.Engine':
        """The CUDA context of ``device`` holding the packed weights: one per GPU, rebuilt when the weights changed."""
        dev = torch.device(device if device is not None else self.device)
        if dev.type == 'cuda' and dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        key = self._weights_key()
        if key != self._engines_key:
            self._engines.clear()
            self._engine_uncond = None
            self._engines_key = key
        eng = self._engines.get(str(dev))
        if eng is None:
            eng = engine.Engine(self.hyper(), {k: v.detach() for k, v in self.state_dict().items()}, dev)
            self._engines[str(dev)] = eng
        return eng

    def invalidate(self):
        """Drop the packed-weight contexts (called after anything that changes the parameters)."""
        self._engines.clear()
        self._engines_key = None
        self._engine_uncond = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        # tolerate DataParallel / ModelWrapper prefixes (utils/model_utils.py:16-21, evaluate.py:167-174)
        cleaned = {}
        for k, v in state_dict.items():
            k = k.replace('module.', '')
            if k.startswith('score_model.'):
                k = k[len('score_model.'):]
            cleaned[k] = v
        out = super().load_state_dict(cleaned, strict=strict, **kw)
        self.invalidate()
        return out

    def to(self, *a, **kw):
        out = super().to(*a, **kw)
        for x in a:
            if isinstance(x, (str, torch.device)):
                self.device = torch.device(x)
        if 'device' in kw:
            self.device = torch.device(kw['device'])
        self.invalidate()
        return out

    def train(self, mode=True):
        if mode:
            raise NotImplementedError('disco_diffdock_b200 implements inference (eval mode) only')
        return super().train(False)

    # ------------------------------------------------------------------------------------------ API
    def forward(self, data):
        """score_model.py:259-308.  Returns (tr_pred [B,3], rot_pred [B,3], tor_pred [sum R])."""
        return engine.forward_batch(self, data)

    def embed(self, data):
        """score_model.py:169-257.  Returns (lig_node_attr, rec_node_attr, tr_sigma, rot_sigma, tor_sigma)."""
        return engine.embed_batch(self, data)
