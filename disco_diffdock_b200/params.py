"""Parameter tree of the coarse-grained score model with the reference's ``state_dict`` key names.

The CUDA path consumes one flat fp32 blob (``weights.py``); this module only owns the *named* parameters so
that the reference's shipped checkpoints load with ``strict=True`` and a fresh model gets the reference's
default initialisation (nn.Linear defaults, xavier-uniform embeddings, identity batch-norm):
``/root/reference/models/score_model.py:49-167``, ``models/layers.py:15-22, 118-149``,
``models/tensor_layers.py:119-145``.  No module here has a ``forward`` -- the math lives in ``csrc/``.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch import nn

LIG_FEATURE_DIMS = [119, 4, 12, 12, 8, 10, 6, 6, 2, 8, 2, 2, 2, 2, 2, 2]   # process_mols.py:62-79
REC_FEATURE_DIMS = [38]                                                    # process_mols.py:88-90
_TABLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tables')


def _w3j(name):
    z = np.load(os.path.join(_TABLES, 'w3j_constants.npz'))
    return torch.from_numpy(z[name]).float()


class AtomEncoderParams(nn.Module):
    def __init__(self, emb_dim, feature_dims, extra_dim):
        super().__init__()
        self.atom_embedding_list = nn.ModuleList()
        for dim in feature_dims:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)
        self.additional_features_embedder = nn.Linear(extra_dim + emb_dim, emb_dim)


def edge_mlp(n_in, ns, dropout):
    return nn.Sequential(nn.Linear(n_in, ns), nn.ReLU(), nn.Dropout(dropout), nn.Linear(ns, ns))


def fc_block(n_in, hidden, n_out, dropout):
    return nn.Sequential(nn.Linear(n_in, hidden), nn.Identity(), nn.ReLU(), nn.Dropout(dropout), nn.Linear(hidden, n_out))


class Smearing(nn.Module):
    def __init__(self, stop, n):
        super().__init__()
        self.register_buffer('offset', torch.linspace(0.0, stop, n))


class IrrepBatchNormParams(nn.Module):
    def __init__(self, n_fields, n_scalar_even):
        super().__init__()
        self.register_buffer('running_mean', torch.zeros(n_scalar_even))
        self.register_buffer('running_var', torch.ones(n_fields))
        self.weight = nn.Parameter(torch.ones(n_fields))
        self.bias = nn.Parameter(torch.zeros(n_scalar_even))


class _Buffers(nn.Module):
    pass


class TensorProductBuffers(nn.Module):
    """The (unused-by-math) buffers e3nn registers, kept so shipped checkpoints load strictly."""

    def __init__(self, out_dim, w3j=()):
        super().__init__()
        self.register_buffer('weight', torch.zeros(0))
        self.register_buffer('output_mask', torch.ones(out_dim))
        self._compiled_main_left_right = _Buffers()
        for name in w3j:
            self._compiled_main_left_right.register_buffer(name, _w3j(name))


def irrep_level_dims(ns, nv, level):
    level = min(level, 3)
    m = {'0e': ns, '1o': nv if level >= 1 else 0, '1e': nv if level >= 2 else 0, '0o': ns if level >= 3 else 0}
    return m


def tp_weight_numel(mi, mo):
    return ((mi['0e'] + mi['1o']) * mo['0e'] + (mi['0e'] + mi['1o'] + mi['1e']) * mo['1o'] +
            (mi['1o'] + mi['1e'] + mi['0o']) * mo['1e'] + (mi['1e'] + mi['0o']) * mo['0o'])


class ConvLayerParams(nn.Module):
    def __init__(self, ns, nv, layer, dropout, faster=True):
        super().__init__()
        mi, mo = irrep_level_dims(ns, nv, layer), irrep_level_dims(ns, nv, layer + 1)
        if not faster:
            self.tp = TensorProductBuffers(0)
        self.fc = nn.ModuleList([fc_block(3 * ns, 3 * ns, tp_weight_numel(mi, mo), dropout) for _ in range(4)])
        self.batch_norm = IrrepBatchNormParams(sum(mo.values()), mo['0e'])


class HeadConvParams(nn.Module):
    def __init__(self, n_edge_features, weight_numel, out_dim, n_fields, n_scalar_even, dropout, w3j=()):
        super().__init__()
        self.tp = TensorProductBuffers(out_dim, w3j)
        self.fc = fc_block(n_edge_features, n_edge_features, weight_numel, dropout)
        self.batch_norm = IrrepBatchNormParams(n_fields, n_scalar_even)
