"""Repack a reference ``state_dict`` into the flat fp32 blob + offset table ``libddk`` consumes.

One-time host cost (evaluate.py:160-174 loads the checkpoint once).  The only non-trivial layout is the second
layer of each per-edge radial MLP (``conv_layers.L.fc.G.4``), which the kernels consume *re-associated*:

    sum_e TP(x_dst, sh_e; W2 h_e + b2)  =  W2p (*) (sum_e basis_e (x) h_e)  +  b2p (*) sum_e basis_e

(valid because the tensor product is linear in its weights and aggregation is a mean).
``basis_e`` is the list of ``FasterTensorProduct`` basis functions of the edge
(/root/reference/models/tensor_layers.py:71-84) *without* their constant factors, in the kernel order

    u = [ 0e: F0e | 1o: comp-major 3 x F1o | 1e: 3 x F1e | 0o: F0o ]

and for every irrep class k the packed weight is ``W2p_k[(u_k, j), o] = W2[off_k + u_k*O_k + o, j] * s(k, u_k)``
with ``s`` = 1/sqrt(fan_in_k) (tensor_layers.py:90-91) times the basis constant (1/sqrt3 for dot-product
entries, 1/sqrt2 for cross-product entries, tensor_layers.py:75-81).
"""
from __future__ import annotations

import math
import os
import re
from typing import Dict

import numpy as np
import torch

from .params import irrep_level_dims

NS, NV, H = 24, 6, 72
_HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'ddk.h')


def parse_header_enums(path=_HDR) -> Dict[str, int]:
    """Read the enum constants out of include/ddk.h so Python and C share one definition."""
    txt = open(path).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    out = {}
    for body in re.findall(r'enum[^{]*\{(.*?)\}', txt, flags=re.S):
        val = -1
        for item in body.split(','):
            item = item.strip()
            if not item:
                continue
            if '=' in item:
                name, expr = [x.strip() for x in item.split('=')]
                val = int(eval(expr, {}, out))
            else:
                name, val = item, val + 1
            out[name] = val
    for name, v in re.findall(r'#define\s+(DDK_\w+)\s+(\d+)', txt):
        out[name] = int(v)
    return out


ENUMS = parse_header_enums()


def class_table(layer: int):
    """Per irrep class (0e,1o,1e,0o): fan-in F, mul_out O, ncomp, basis offset, per-u_k scale, reference row offset."""
    mi, mo = irrep_level_dims(NS, NV, layer), irrep_level_dims(NS, NV, layer + 1)
    r3, r2 = 1 / math.sqrt(3.0), 1 / math.sqrt(2.0)
    spec = {
        '0e': [(mi['0e'], 1.0), (mi['1o'], r3)],
        '1o': [(mi['0e'], 1.0), (mi['1o'], 1.0), (mi['1e'], r2)],
        '1e': [(mi['1o'], r2), (mi['1e'], 1.0), (mi['0o'], 1.0)],
        '0o': [(mi['1e'], r3), (mi['0o'], 1.0)],
    }
    table, uoff, roff = [], 0, 0
    for k in ('0e', '1o', '1e', '0o'):
        F = sum(n for n, _ in spec[k])
        O = mo[k]
        ncomp = 3 if k[0] == '1' else 1
        scale = np.concatenate([np.full(n, c) for n, c in spec[k]]) / math.sqrt(F) if F else np.zeros(0)
        table.append(dict(key=k, F=F, O=O, ncomp=ncomp, uoff=uoff, scale=scale, roff=roff))
        if F and O:
            uoff += ncomp * F
        roff += F * O
    return table, uoff


def pack_conv_second_layer(w: np.ndarray, b: np.ndarray, layer: int):
    table, U = class_table(layer)
    wp, bp = [], []
    for c in table:
        F, O = c['F'], c['O']
        if F == 0 or O == 0:
            continue
        blk = w[c['roff']:c['roff'] + F * O].reshape(F, O, H) * c['scale'][:, None, None]      # [u_k, o, j]
        wp.append(np.ascontiguousarray(blk.transpose(0, 2, 1)).reshape(F * H, O).ravel())      # [(u_k, j), o]
        bp.append((b[c['roff']:c['roff'] + F * O].reshape(F, O) * c['scale'][:, None]).ravel())
    return np.concatenate(wp), np.concatenate(bp)


def bn_affine_84(sd, key, layer, eps=1e-5):
    mo = irrep_level_dims(NS, NV, layer + 1)
    w, rv = sd[key + '.weight'].double().numpy(), sd[key + '.running_var'].double().numpy()
    rm, b = sd[key + '.running_mean'].double().numpy(), sd[key + '.bias'].double().numpy()
    s = w / np.sqrt(rv + eps)
    scale, shift = np.zeros(84), np.zeros(84)
    scale[0:NS] = s[0:NS]
    shift[0:NS] = b - rm * s[0:NS]
    f = NS
    for name, base in (('1o', 24), ('1e', 42)):
        if mo[name]:
            for i in range(NV):
                scale[base + 3 * i: base + 3 * i + 3] = s[f + i]
            f += NV
    if mo['0o']:
        scale[60:84] = s[f:f + NS]
    return scale.astype(np.float32), shift.astype(np.float32)


def pack_weights(sd: Dict[str, torch.Tensor], hyper):
    """Returns (blob float32 [n], offsets int64 [n_offsets])."""
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    L = hyper.num_conv_layers
    n_off = ENUMS['DDK_W_CONV_BASE'] + L * ENUMS['DDK_W_CONV_STRIDE']
    offsets = np.full(n_off, -1, dtype=np.int64)
    chunks, pos = [], 0

    def put(idx, arr):
        nonlocal pos
        arr = np.ascontiguousarray(np.asarray(arr, dtype=np.float32)).ravel()
        pad = (-pos) % 4                      # keep every tensor 16-byte aligned
        if pad:
            chunks.append(np.zeros(pad, np.float32))
            pos += pad
        offsets[idx] = pos
        chunks.append(arr)
        pos += arr.size

    def t(key):
        return sd[key].detach().cpu().float().numpy()

    E = ENUMS
    put(E['DDK_W_LIG_EMB_TABLES'], np.concatenate([t(f'lig_node_embedding.atom_embedding_list.{i}.weight') for i in range(16)]))
    put(E['DDK_W_LIG_NODE_W'], t('lig_node_embedding.additional_features_embedder.weight'))
    put(E['DDK_W_LIG_NODE_B'], t('lig_node_embedding.additional_features_embedder.bias'))
    put(E['DDK_W_REC_EMB_TABLE'], t('rec_node_embedding.atom_embedding_list.0.weight'))
    put(E['DDK_W_REC_NODE_W'], t('rec_node_embedding.additional_features_embedder.weight'))
    put(E['DDK_W_REC_NODE_B'], t('rec_node_embedding.additional_features_embedder.bias'))
    for name, key in (('LIG_EDGE', 'lig_edge_embedding'), ('REC_EDGE', 'rec_edge_embedding'),
                      ('CROSS_EDGE', 'cross_edge_embedding'), ('CENTER_EDGE', 'center_edge_embedding'),
                      ('FINAL_EDGE', 'final_edge_embedding')):
        if key + '.0.weight' not in sd:        # no_torsion models have no final_edge_embedding
            continue
        put(E[f'DDK_W_{name}_W1'], t(key + '.0.weight'))
        put(E[f'DDK_W_{name}_B1'], t(key + '.0.bias'))
        put(E[f'DDK_W_{name}_W2'], t(key + '.3.weight'))
        put(E[f'DDK_W_{name}_B2'], t(key + '.3.bias'))
    unc = np.zeros((5, NS), np.float32)
    for i, nm in enumerate(('lig_node', 'rec_node', 'lig_edge', 'rec_edge', 'cross_edge')):
        k = f'{nm}_unconditional_embedding'
        if k in sd:
            unc[i] = t(k).reshape(-1)
    put(E['DDK_W_UNCOND'], unc)
    put(E['DDK_W_FINAL_CONV_W1'], t('final_conv.fc.0.weight'))
    put(E['DDK_W_FINAL_CONV_B1'], t('final_conv.fc.0.bias'))
    put(E['DDK_W_FINAL_CONV_W2'], t('final_conv.fc.4.weight'))
    put(E['DDK_W_FINAL_CONV_B2'], t('final_conv.fc.4.bias'))
    put(E['DDK_W_FINAL_CONV_BN'], (sd['final_conv.batch_norm.weight'].double() /
                                   torch.sqrt(sd['final_conv.batch_norm.running_var'].double() + 1e-5)).float().numpy())
    for nm, key in (('TR', 'tr_final_layer'), ('ROT', 'rot_final_layer')):
        put(E[f'DDK_W_{nm}_FINAL_W1'], t(key + '.0.weight'))
        put(E[f'DDK_W_{nm}_FINAL_B1'], t(key + '.0.bias'))
        put(E[f'DDK_W_{nm}_FINAL_W2'], t(key + '.3.weight'))
        put(E[f'DDK_W_{nm}_FINAL_B2'], t(key + '.3.bias'))
    if 'tor_bond_conv.fc.0.weight' in sd:
        put(E['DDK_W_TOR_CONV_W1'], t('tor_bond_conv.fc.0.weight'))
        put(E['DDK_W_TOR_CONV_B1'], t('tor_bond_conv.fc.0.bias'))
        put(E['DDK_W_TOR_CONV_W2'], t('tor_bond_conv.fc.4.weight'))
        put(E['DDK_W_TOR_CONV_B2'], t('tor_bond_conv.fc.4.bias'))
        w = sd['tor_bond_conv.batch_norm.weight'].double().numpy()
        rv = sd['tor_bond_conv.batch_norm.running_var'].double().numpy()
        rm = sd['tor_bond_conv.batch_norm.running_mean'].double().numpy()
        bb = sd['tor_bond_conv.batch_norm.bias'].double().numpy()
        s = w / np.sqrt(rv + 1e-5)                       # fields: 24 x 0o (scale only) then 24 x 0e
        shift = np.zeros(2 * NS)
        shift[NS:] = bb - rm * s[NS:]
        put(E['DDK_W_TOR_CONV_BN_SCALE'], s)
        put(E['DDK_W_TOR_CONV_BN_SHIFT'], shift)
        put(E['DDK_W_TOR_FINAL_W1'], t('tor_final_layer.0.weight'))
        put(E['DDK_W_TOR_FINAL_W2'], t('tor_final_layer.3.weight'))
    sm = np.zeros((4, 33), np.float32)
    for i, key in enumerate(('lig_distance_expansion', 'rec_distance_expansion', 'cross_distance_expansion',
                             'center_distance_expansion')):
        offs = sd[key + '.offset'].detach().cpu().float()
        sm[i, :32] = offs.numpy()
        sm[i, 32] = -0.5 / (offs[1] - offs[0]).item() ** 2          # tensor_layers.py:176
    put(E['DDK_W_SMEAR'], sm)
    base, stride = E['DDK_W_CONV_BASE'], E['DDK_W_CONV_STRIDE']
    for l in range(L):
        o = base + l * stride
        for g in range(4):
            put(o + E['DDK_WL_W1'] + g, t(f'conv_layers.{l}.fc.{g}.0.weight'))
            put(o + E['DDK_WL_B1'] + g, t(f'conv_layers.{l}.fc.{g}.0.bias'))
            wp, bp = pack_conv_second_layer(t(f'conv_layers.{l}.fc.{g}.4.weight').astype(np.float64),
                                            t(f'conv_layers.{l}.fc.{g}.4.bias').astype(np.float64), l)
            put(o + E['DDK_WL_W2P'] + g, wp)
            put(o + E['DDK_WL_B2P'] + g, bp)
        sc, sh = bn_affine_84(sd, f'conv_layers.{l}.batch_norm', l)
        put(o + E['DDK_WL_BN_SCALE'], sc)
        put(o + E['DDK_WL_BN_SHIFT'], sh)
    blob = np.concatenate(chunks)
    return blob, offsets
