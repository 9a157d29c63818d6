"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's reverse-diffusion docking hot path.

This module is the *oracle*: a plain functional torch-CPU restatement of
``TensorProductScoreModel.forward`` and of the ``sampling()`` loop body of gcorso/disco-diffdock, written
from the reference's algorithm (each function cites the ``/root/reference`` file:line it follows).  It is
the checker for the CUDA path and the ``cpu_baseline`` of ``bench.py``; nothing in the product package may
import it (``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` only).

Parity pinning: the reference ships no tests or golden vectors.  The oracle is pinned
instead against the reference's *own code* executed in the build container behind leaf-op shims
(``oracle/ref_shims.py`` / ``oracle/ref_loader.py``): ``tests/test_oracle_vs_reference.py`` runs both on the
same inputs when ``/root/reference`` is present, and ``oracle/make_golden.py`` stores reference-generated
vectors in ``tests/golden/`` that ``tests/test_oracle_golden.py`` replays anywhere.  The third-party leaf
ops (e3nn, torch_cluster, torch_scatter; unpinned versions) are restated from their published semantics;
that residual risk is listed in DESIGN.md ("parity pinned to reference code, leaf ops restated").

All arithmetic is fp32 like the reference.  ``params`` is a flat ``{state_dict key: tensor}`` mapping with
the reference's key names (``workdir/*/best_ema_inference_epoch_model.pt``).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

LIG_FEATURE_DIMS = [119, 4, 12, 12, 8, 10, 6, 6, 2, 8, 2, 2, 2, 2, 2, 2]   # process_mols.py:62-79
REC_FEATURE_DIMS = [38]                                                    # process_mols.py:88-90


def default_config(**over):
    """Hyper-parameters of the shipped DiffDock-S model (workdir/diffdockS_score_model/model_parameters.yml)
    plus the constructor defaults ``get_model`` never overrides (score_model.py:15-17; model_utils.py:39-68)."""
    cfg = dict(ns=24, nv=6, num_conv_layers=5, sh_lmax=1, sigma_embed_dim=32, distance_embed_dim=32,
               cross_distance_embed_dim=32, lig_max_radius=5.0, rec_max_radius=30.0, cross_max_distance=80.0,
               center_max_distance=30.0, dynamic_max_cross=True, scale_by_sigma=True, no_torsion=False,
               embedding_scale=1000.0, in_lig_edge_features=4, lm_embedding_dim=1280,
               latent_dim=0, latent_vocab=1, latent_droprate=0.0,
               tr_sigma_min=0.1, tr_sigma_max=19.0, rot_sigma_min=0.03, rot_sigma_max=1.55,
               tor_sigma_min=0.03, tor_sigma_max=3.14)
    cfg.update(over)
    return SimpleNamespace(**cfg)


# ------------------------------------------------------------------------------------------------ leaf ops
def t_to_sigma(cfg, t_tr, t_rot, t_tor):
    """diffusion_utils.py:12-16."""
    return (cfg.tr_sigma_min ** (1 - t_tr) * cfg.tr_sigma_max ** t_tr,
            cfg.rot_sigma_min ** (1 - t_rot) * cfg.rot_sigma_max ** t_rot,
            cfg.tor_sigma_min ** (1 - t_tor) * cfg.tor_sigma_max ** t_tor)


def timestep_embedding(t, dim, scale):
    """diffusion_utils.py:58-69 through get_timestep_embedding (:87-94): sinusoidal, max_positions 1e4."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    arg = (scale * t).float()[:, None] * freq[None, :]
    return torch.cat([torch.sin(arg), torch.cos(arg)], 1)


def smear(dist, stop, n):
    """GaussianSmearing, tensor_layers.py:171-181 (start=0)."""
    mu = torch.linspace(0.0, stop, n)
    coeff = -0.5 / (mu[1] - mu[0]).item() ** 2
    return torch.exp(coeff * (dist.reshape(-1, 1) - mu.reshape(1, -1)) ** 2)


def sh_l01(vec):
    """e3nn spherical_harmonics('1x0e+1x1o', normalize=True, 'component'): [1, sqrt3 * unit(vec)]."""
    n = vec / vec.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    return torch.cat([torch.ones_like(n[:, :1]), math.sqrt(3.0) * n], 1)


def sh_l2(vec):
    """e3nn spherical_harmonics('2e', normalize=True, 'component') (score_model.py:295)."""
    n = vec / vec.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    x, y, z = n[:, 0], n[:, 1], n[:, 2]
    s3 = math.sqrt(3.0)
    return math.sqrt(5.0) * torch.stack([s3 * x * z, s3 * x * y, y * y - 0.5 * (x * x + z * z), s3 * y * z,
                                         s3 / 2 * (z * z - x * x)], 1)


def dist2_unfused(q, c):
    """fp32 ((dx^2 + dy^2) + dz^2), each op rounded -- the CUDA path uses __fmul_rn/__fadd_rn to match."""
    d = c.unsqueeze(0) - q.unsqueeze(1)
    sq = d * d
    return (sq[..., 0] + sq[..., 1]) + sq[..., 2]


def radius_pairs(x, y, r, batch_x, batch_y, max_num_neighbors):
    """torch_cluster.radius semantics (CUDA kernel): for each query y_i the first ``max_num_neighbors``
    candidates x_j (ascending j, same graph) with |x_j-y_i|^2 < r^2; returns (query idx, candidate idx)."""
    rows, cols = [], []
    nb = int(max(int(batch_x.max()), int(batch_y.max()))) + 1 if len(batch_x) and len(batch_y) else 0
    r2 = torch.tensor(float(r) * float(r), dtype=torch.float32)
    for b in range(nb):
        iy = torch.nonzero(batch_y == b).flatten()
        ix = torch.nonzero(batch_x == b).flatten()
        if len(iy) == 0 or len(ix) == 0:
            continue
        hit = dist2_unfused(y[iy], x[ix]) < r2
        hit &= (torch.cumsum(hit.long(), 1) - 1) < max_num_neighbors
        q, c = torch.nonzero(hit, as_tuple=True)
        rows.append(iy[q])
        cols.append(ix[c])
    if not rows:
        z = torch.zeros(0, dtype=torch.long)
        return z, z
    return torch.cat(rows), torch.cat(cols)


def segment_mean(src, index, n):
    """torch_scatter.scatter(..., reduce='mean'): sum / max(count, 1)."""
    out = src.new_zeros(n, src.shape[1])
    out.index_add_(0, index, src)
    cnt = torch.bincount(index, minlength=n).clamp(min=1).to(src.dtype)
    return out / cnt[:, None]


def linear(p, key, x):
    b = p.get(key + '.bias')
    return F.linear(x, p[key + '.weight'], b)


def mlp_relu(p, key, x, second=3):
    """nn.Sequential(Linear, ReLU, Dropout, Linear) in eval mode (score_model.py:51-56, 125-130, 146-151)."""
    return linear(p, f'{key}.{second}', torch.relu(linear(p, f'{key}.0', x)))


def atom_encoder(p, key, x, n_cat):
    """AtomEncoder.forward, layers.py:140-149: sum of categorical embeddings, then Linear over [emb | rest]."""
    emb = 0
    for i in range(n_cat):
        emb = emb + p[f'{key}.atom_embedding_list.{i}.weight'][x[:, i].long()]
    return linear(p, f'{key}.additional_features_embedder', torch.cat([emb, x[:, n_cat:].float()], 1))


def irrep_muls(cfg, level):
    """get_irrep_seq (tensor_layers.py:21-26) as {irrep: multiplicity}; order 0e,1o,1e,0o."""
    level = min(level, 3)
    m = {'0e': cfg.ns, '1o': 0, '1e': 0, '0o': 0}
    if level >= 1:
        m['1o'] = cfg.nv
    if level >= 2:
        m['1e'] = cfg.nv
    if level >= 3:
        m['0o'] = cfg.ns
    return m


def feat_dim(m):
    return m['0e'] + 3 * m['1o'] + 3 * m['1e'] + m['0o']


def split_irreps(x, m):
    o = 0
    s0e = x[:, o:o + m['0e']]; o += m['0e']
    v1o = x[:, o:o + 3 * m['1o']].reshape(len(x), m['1o'], 3); o += 3 * m['1o']
    v1e = x[:, o:o + 3 * m['1e']].reshape(len(x), m['1e'], 3); o += 3 * m['1e']
    s0o = x[:, o:o + m['0o']]
    return s0e, v1o, v1e, s0o


def faster_tp(x, sh, w, mi, mo):
    """FasterTensorProduct.forward (tensor_layers.py:65-116) for lmax=1 filters.

    Basis per output irrep (concatenation order matters because the weight rows follow it):
      0e <- [x0e*sh0 ; (x1o . s)/sqrt3]         1o <- [x0e (x) s ; x1o*sh0 ; (x1e x s)/sqrt2]
      1e <- [(x1o x s)/sqrt2 ; x1e*sh0 ; x0o (x) s]   0o <- [(x1e . s)/sqrt3 ; x0o*sh0]
    weights: blocks in order 0e,1o,1e,0o, each (fan_in, mul_out) row-major, divided by sqrt(fan_in).
    """
    s0e, v1o, v1e, s0o = split_irreps(x, mi)
    sh0, s = sh[:, 0:1], sh[:, 1:4]
    sv = s[:, None, :]
    b0e = [s0e * sh0]
    b1o = [s0e[:, :, None] * sv]
    b1e, b0o = [], []
    if mi['1o']:
        b0e.append((v1o * sv).sum(-1) / math.sqrt(3))
        b1o.append(v1o * sh0[:, :, None])
        b1e.append(torch.linalg.cross(v1o, sv.expand_as(v1o), dim=-1) / math.sqrt(2))
    if mi['1e']:
        b1o.append(torch.linalg.cross(v1e, sv.expand_as(v1e), dim=-1) / math.sqrt(2))
        b1e.append(v1e * sh0[:, :, None])
        b0o.append((v1e * sv).sum(-1) / math.sqrt(3))
    if mi['0o']:
        b1e.append(s0o[:, :, None] * sv)
        b0o.append(s0o * sh0)
    fan = {'0e': mi['0e'] + mi['1o'], '1o': mi['0e'] + mi['1o'] + mi['1e'],
           '1e': mi['1o'] + mi['1e'] + mi['0o'], '0o': mi['1e'] + mi['0o']}
    basis = {'0e': b0e, '1o': b1o, '1e': b1e, '0o': b0o}
    outs, off = [], 0
    for k in ('0e', '1o', '1e', '0o'):
        n = fan[k] * mo[k]
        wk = w[:, off:off + n].reshape(len(x), fan[k], mo[k]) / math.sqrt(fan[k]) if n else None
        off += n
        if mo[k] == 0:
            continue
        if k in ('0e', '0o'):
            bk = torch.cat(basis[k], 1)                                  # [E, fan]
            outs.append(torch.einsum('eu,euw->ew', bk, wk))
        else:
            bk = torch.cat(basis[k], 1)                                  # [E, fan, 3]
            outs.append(torch.einsum('euc,euw->ewc', bk, wk).reshape(len(x), -1))
    return torch.cat(outs, 1)


def tp_weight_numel(mi, mo):
    return ((mi['0e'] + mi['1o']) * mo['0e'] + (mi['0e'] + mi['1o'] + mi['1e']) * mo['1o'] +
            (mi['1o'] + mi['1e'] + mi['0o']) * mo['1e'] + (mi['1e'] + mi['0o']) * mo['0o'])


def bn_eval(p, key, x, blocks, eps=1e-5):
    """e3nn.nn.BatchNorm eval mode; ``blocks`` = [(mul, dim, is_0e)] in feature order.  Only 0e channels
    subtract running_mean and add bias; one scale per channel shared by a vector's 3 components."""
    w, rv = p[key + '.weight'], p[key + '.running_var']
    rm, b = p[key + '.running_mean'], p[key + '.bias']
    out, ix, iw, ib = [], 0, 0, 0
    for mul, d, is0e in blocks:
        f = x[:, ix:ix + mul * d].reshape(-1, mul, d)
        ix += mul * d
        if is0e:
            f = f - rm[ib:ib + mul].reshape(1, mul, 1)
        f = f * (rv[iw:iw + mul] + eps).pow(-0.5).reshape(1, mul, 1)
        f = f * w[iw:iw + mul].reshape(1, mul, 1)
        if is0e:
            f = f + b[ib:ib + mul].reshape(1, mul, 1)
            ib += mul
        iw += mul
        out.append(f.reshape(-1, mul * d))
    return torch.cat(out, 1)


def muls_blocks(m):
    return [(m[k], 3 if k[0] == '1' else 1, k == '0e') for k in ('0e', '1o', '1e', '0o') if m[k]]


# ------------------------------------------------------------------------------------------------ model
def _latents(cfg, data):
    if cfg.latent_dim > 0:
        assert cfg.latent_vocab == 1, 'only the equivariant (vocab==1) latents of DisCo-DiffDock-S are restated'
        return data['ligand'].latent_h.float(), data['receptor'].latent_h.float()
    return None


def embed(p, cfg, data, trace: Optional[dict] = None):
    """TensorProductScoreModel.embed (score_model.py:169-257)."""
    ns = cfg.ns
    lig, rec = data['ligand'], data['receptor']
    lat = _latents(cfg, data)
    tr_sigma, rot_sigma, tor_sigma = t_to_sigma(cfg, *[data.complex_t[k] for k in ('tr', 'rot', 'tor')])
    lig_semb = timestep_embedding(lig.node_t['tr'], cfg.sigma_embed_dim, cfg.embedding_scale)
    rec_semb = timestep_embedding(rec.node_t['tr'], cfg.sigma_embed_dim, cfg.embedding_scale)
    lig.node_sigma_emb, rec.node_sigma_emb = lig_semb, rec_semb          # side effects kept (:312, :348)

    # ligand graph: bonds + radius graph (score_model.py:310-344)
    pos_l, pos_r = lig.pos.float(), rec.pos.float()
    bond_ei = data['ligand', 'ligand'].edge_index.long()
    centre, neigh = radius_pairs(pos_l, pos_l, cfg.lig_max_radius, lig.batch, lig.batch, 33)
    keep = centre != neigh
    ll_src = torch.cat([bond_ei[0], neigh[keep]])
    ll_dst = torch.cat([bond_ei[1], centre[keep]])
    ll_attr = torch.cat([data['ligand', 'ligand'].edge_attr.float(),
                         torch.zeros(int(keep.sum()), cfg.in_lig_edge_features)], 0)
    ll_vec = pos_l[ll_dst] - pos_l[ll_src]
    ll_feats = [ll_attr, lig_semb[ll_src], smear(ll_vec.norm(dim=-1), cfg.lig_max_radius, cfg.distance_embed_dim)]
    lig_x = [lig.x.float(), lig_semb]
    if lat is not None:
        ll_feats.append(torch.cat([lat[0][ll_src], lat[0][ll_dst]], 1))
        lig_x.append(lat[0])
    ll_sh = sh_l01(ll_vec)
    lig_h = atom_encoder(p, 'lig_node_embedding', torch.cat(lig_x, 1), len(LIG_FEATURE_DIMS))
    ll_ea = mlp_relu(p, 'lig_edge_embedding', torch.cat(ll_feats, 1))

    # receptor graph (score_model.py:346-373)
    rr_ei = data['receptor', 'receptor'].edge_index.long()
    rr_vec = pos_r[rr_ei[1]] - pos_r[rr_ei[0]]
    rr_feats = [rec_semb[rr_ei[0]], smear(rr_vec.norm(dim=-1), cfg.rec_max_radius, cfg.distance_embed_dim)]
    rec_x = [rec.x.float(), rec_semb]
    if lat is not None:
        rr_feats.append(torch.cat([lat[1][rr_ei[0]], lat[1][rr_ei[1]]], 1))
        rec_x.append(lat[1])
    rr_sh = sh_l01(rr_vec)
    rec_h = atom_encoder(p, 'rec_node_embedding', torch.cat(rec_x, 1), len(REC_FEATURE_DIMS))
    rr_ea = mlp_relu(p, 'rec_edge_embedding', torch.cat(rr_feats, 1))

    # cross graph (score_model.py:375-408, cutoff :202-205)
    if cfg.dynamic_max_cross:
        cut = (tr_sigma * 3 + 20).float().unsqueeze(1)
        lr_src, lr_dst = radius_pairs(pos_r / cut[rec.batch], pos_l / cut[lig.batch], 1.0, rec.batch, lig.batch, 10000)
    else:
        lr_src, lr_dst = radius_pairs(pos_r, pos_l, cfg.cross_max_distance, rec.batch, lig.batch, 10000)
    lr_vec = pos_r[lr_dst] - pos_l[lr_src]
    lr_feats = [lig_semb[lr_src], smear(lr_vec.norm(dim=-1), cfg.cross_max_distance, cfg.cross_distance_embed_dim)]
    if lat is not None:
        lr_feats.append(torch.zeros(len(lr_src), 2 * cfg.latent_dim))       # zeroed latents (:401)
    lr_sh = sh_l01(lr_vec)
    lr_ea = mlp_relu(p, 'cross_edge_embedding', torch.cat(lr_feats, 1))

    if cfg.latent_droprate > 0:                                              # score_model.py:209-215
        ul, ur = lig.unconditional.float(), rec.unconditional.float()
        lig_h = lig_h + ul * p['lig_node_unconditional_embedding']
        rec_h = rec_h + ur * p['rec_node_unconditional_embedding']
        ll_ea = ll_ea + ul[ll_src] * p['lig_edge_unconditional_embedding']
        rr_ea = rr_ea + ur[rr_ei[0]] * p['rec_edge_unconditional_embedding']
        lr_ea = lr_ea + ul[lr_src] * p['cross_edge_unconditional_embedding']

    # combined graph (score_model.py:218-225): groups [ll | l->r | rr | r->l]; reversed cross edges reuse
    # the forward attributes and the *un-negated* harmonics.
    nl = len(lig_h)
    x = torch.cat([lig_h, rec_h], 0)
    src = torch.cat([ll_src, lr_src, rr_ei[0] + nl, lr_dst + nl])
    dst = torch.cat([ll_dst, lr_dst + nl, rr_ei[1] + nl, lr_src])
    ea = torch.cat([ll_ea, lr_ea, rr_ea, lr_ea], 0)
    sh = torch.cat([ll_sh, lr_sh, rr_sh, lr_sh], 0)
    bounds = np.cumsum([0, len(ll_src), len(lr_src), rr_ei.shape[1], len(lr_src)])
    if trace is not None:
        trace.update(lig_h0=lig_h, rec_h0=rec_h, ll_src=ll_src, ll_dst=ll_dst, ll_ea=ll_ea, ll_sh=ll_sh,
                     lr_src=lr_src, lr_dst=lr_dst, lr_ea=lr_ea, lr_sh=lr_sh, rr_ea=rr_ea, rr_sh=rr_sh,
                     n_edges=int(bounds[-1]))

    for l in range(cfg.num_conv_layers):                                     # score_model.py:227-230
        mi, mo = irrep_muls(cfg, l), irrep_muls(cfg, l + 1)
        feats = torch.cat([ea, x[src, :ns], x[dst, :ns]], 1)
        w = torch.cat([mlp_relu(p, f'conv_layers.{l}.fc.{g}', feats[bounds[g]:bounds[g + 1]], second=4)
                       for g in range(4)], 0)                                # tensor_layers.py:154-155
        msg = faster_tp(x[dst], sh, w, mi, mo)                               # :156
        agg = segment_mean(msg, src, len(x))                                 # :159
        agg = bn_eval(p, f'conv_layers.{l}.batch_norm', agg, muls_blocks(mo))  # :162
        x = agg + F.pad(x, (0, agg.shape[1] - x.shape[1]))                   # :165-166
        if trace is not None:
            trace[f'x{l + 1}'] = x
    return x[:nl], x[nl:], tr_sigma, rot_sigma, tor_sigma


def fctp_final_conv(x, sh, w, cfg):
    """e3nn FullyConnectedTensorProduct(84-irreps (x) (0e+1o) -> 2x1o + 2x1e), score_model.py:132-140.
    Weight blocks (in1-major, in2, out): [0e.1o->1o 2ns | 1o.0e->1o 2nv | 1o.1o->1e 2nv | 1e.0e->1e 2nv |
    1e.1o->1o 2nv | 0o.1o->1e 2ns]; both outputs have fan-in ns+2nv -> path weight sqrt(3/(ns+2nv));
    C(0,1,1)=C(1,0,1)=delta/sqrt3, C(1,1,1)=eps/sqrt6."""
    ns, nv = cfg.ns, cfg.nv
    m = irrep_muls(cfg, 3)
    s0e, v1o, v1e, s0o = split_irreps(x, m)
    sh0, s = sh[:, 0:1], sh[:, 1:4]
    sv = s[:, None, :]
    o = 0
    def blk(n_in):
        nonlocal o
        b = w[:, o:o + n_in * 2].reshape(-1, n_in, 2)
        o += n_in * 2
        return b
    w1, w2, w3, w4, w5, w6 = blk(ns), blk(nv), blk(nv), blk(nv), blk(nv), blk(ns)
    pw = math.sqrt(3.0 / (ns + 2 * nv))
    r3, r6 = 1 / math.sqrt(3.0), 1 / math.sqrt(6.0)
    out1o = (torch.einsum('eu,ec,euw->ewc', s0e, s, w1) * r3
             + torch.einsum('euc,euw->ewc', v1o * sh0[:, :, None], w2) * r3
             + torch.einsum('euc,euw->ewc', torch.linalg.cross(v1e, sv.expand_as(v1e), dim=-1), w5) * r6)
    out1e = (torch.einsum('euc,euw->ewc', torch.linalg.cross(v1o, sv.expand_as(v1o), dim=-1), w3) * r6
             + torch.einsum('euc,euw->ewc', v1e * sh0[:, :, None], w4) * r3
             + torch.einsum('eu,ec,euw->ewc', s0o, s, w6) * r3)
    return pw * torch.cat([out1o.reshape(len(x), 6), out1e.reshape(len(x), 6)], 1)


_C121 = None


def c121():
    """e3nn's real-basis Wigner 3j (1,2,1) as embedded in the reference checkpoints
    (final_tp_tor._compiled_main_left_right._w3j_1_2_1)."""
    global _C121
    if _C121 is None:
        a, b = 1 / math.sqrt(10.0), 1 / math.sqrt(30.0)
        c = torch.zeros(3, 5, 3)
        for (i, j, k) in [(0, 0, 2), (0, 1, 1), (1, 1, 0), (1, 3, 2), (2, 0, 0), (2, 3, 1), (2, 4, 2)]:
            c[i, j, k] = a
        c[0, 4, 0] = -a
        c[1, 2, 1] = 2 * b
        c[0, 2, 0] = -b
        c[2, 2, 2] = -b
        _C121 = c
    return _C121


def forward(p, cfg, data, tables, trace: Optional[dict] = None):
    """TensorProductScoreModel.forward (score_model.py:259-308) -> (tr_pred, rot_pred, tor_pred)."""
    ns, nv = cfg.ns, cfg.nv
    lig_h, rec_h, tr_sigma, rot_sigma, tor_sigma = embed(p, cfg, data, trace)
    lig = data['ligand']
    pos = lig.pos.float()
    B = int(data.num_graphs)

    # translation / rotation head (score_model.py:269-286, build_center_conv_graph :410-423)
    center = torch.zeros(B, 3).index_add_(0, lig.batch, pos) / torch.bincount(lig.batch, minlength=B).unsqueeze(1)
    cvec = pos - center[lig.batch]
    cattr = torch.cat([smear(cvec.norm(dim=-1), cfg.center_max_distance, cfg.distance_embed_dim), lig.node_sigma_emb], 1)
    csh = sh_l01(cvec)
    cfeat = torch.cat([mlp_relu(p, 'center_edge_embedding', cattr), lig_h[:, :ns]], 1)
    cw = mlp_relu(p, 'final_conv.fc', cfeat, second=4)
    gp = segment_mean(fctp_final_conv(lig_h, csh, cw, cfg), lig.batch, B)
    gp = bn_eval(p, 'final_conv.batch_norm', gp, [(2, 3, False), (2, 3, False)])
    tr_pred = gp[:, 0:3] + gp[:, 6:9]
    rot_pred = gp[:, 3:6] + gp[:, 9:12]
    gemb = timestep_embedding(data.complex_t['tr'], cfg.sigma_embed_dim, cfg.embedding_scale)
    data.graph_sigma_emb = gemb
    tr_norm = torch.linalg.vector_norm(tr_pred, dim=1).unsqueeze(1)
    tr_pred = tr_pred / tr_norm * mlp_relu(p, 'tr_final_layer', torch.cat([tr_norm, gemb], 1))
    rot_norm = torch.linalg.vector_norm(rot_pred, dim=1).unsqueeze(1)
    rot_pred = rot_pred / rot_norm * mlp_relu(p, 'rot_final_layer', torch.cat([rot_norm, gemb], 1))
    if cfg.scale_by_sigma:
        tr_pred = tr_pred / tr_sigma.float().unsqueeze(1)
        rot_pred = rot_pred * so3_score_norm(tables, rot_sigma).unsqueeze(1)
    if trace is not None:
        trace.update(lig_h=lig_h, rec_h=rec_h, global_pred=gp)

    mask = lig.edge_mask
    if cfg.no_torsion or int(mask.sum()) == 0:
        return tr_pred, rot_pred, torch.empty(0)

    # torsion head (score_model.py:291-307, build_bond_conv_graph :425-438)
    bonds = data['ligand', 'ligand'].edge_index[:, mask].long()
    bpos = (pos[bonds[0]] + pos[bonds[1]]) / 2
    bbatch = lig.batch[bonds[0]]
    t_bond, t_atom = radius_pairs(pos, bpos, cfg.lig_max_radius, lig.batch, bbatch, 32)
    tvec = pos[t_atom] - bpos[t_bond]
    tattr = mlp_relu(p, 'final_edge_embedding', smear(tvec.norm(dim=-1), cfg.lig_max_radius, cfg.distance_embed_dim))
    tsh = sh_l01(tvec)
    y2 = sh_l2(pos[bonds[1]] - pos[bonds[0]])
    # FullTensorProduct(0e+1o, 2e): only its 1o block can reach l=0 outputs: sqrt3 * C121[i,j,k] s_i Y2_j
    filt = math.sqrt(3.0) * torch.einsum('ijk,ei,ej->ek', c121(), tsh[:, 1:4], y2[t_bond])
    battr = lig_h[bonds[0]] + lig_h[bonds[1]]
    tfeat = torch.cat([tattr, lig_h[t_atom, :ns], battr[t_bond, :ns]], 1)
    tw = mlp_relu(p, 'tor_bond_conv.fc', tfeat, second=4)
    m = irrep_muls(cfg, 3)
    _, v1o, v1e, _ = split_irreps(lig_h[t_atom], m)
    # FCTP -> 'ns x0o + ns x0e': blocks [1o.1o->0e | 1e.1o->0o], path weight sqrt(1/nv), C(1,1,0)=delta/sqrt3
    w_e = tw[:, :nv * ns].reshape(-1, nv, ns)
    w_o = tw[:, nv * ns:].reshape(-1, nv, ns)
    pw = math.sqrt(1.0 / nv) / math.sqrt(3.0)
    out0e = pw * torch.einsum('eu,euw->ew', (v1o * filt[:, None, :]).sum(-1), w_e)
    out0o = pw * torch.einsum('eu,euw->ew', (v1e * filt[:, None, :]).sum(-1), w_o)
    nb = bonds.shape[1]
    tor = segment_mean(torch.cat([out0o, out0e], 1), t_bond, nb)
    tor = bn_eval(p, 'tor_bond_conv.batch_norm', tor, [(ns, 1, False), (ns, 1, True)])
    if trace is not None:
        trace.update(tor_feat=tor, t_bond=t_bond, t_atom=t_atom)
    tor_pred = F.linear(torch.tanh(F.linear(tor, p['tor_final_layer.0.weight'])), p['tor_final_layer.3.weight']).squeeze(1)
    if cfg.scale_by_sigma:
        edge_sigma = tor_sigma[lig.batch][data['ligand', 'ligand'].edge_index[0]][mask]
        tor_pred = tor_pred * torch.sqrt(torus_score_norm(tables, edge_sigma))
    return tr_pred, rot_pred, tor_pred


# ------------------------------------------------------------------------------------------------ LUTs
def so3_score_norm(tables, eps):
    """so3.score_norm (so3.py:91-95).  ``eps.numpy()`` is float32, so log10 is evaluated in float32 and the
    rest in float64 (the np.float64 constants promote), exactly as the reference module does under this numpy."""
    e = eps.float().numpy() if torch.is_tensor(eps) else np.asarray(eps, dtype=np.float32)
    idx = (np.log10(e) - np.log10(0.01)) / (np.log10(2) - np.log10(0.01)) * 1000
    idx = np.clip(np.around(idx).astype(int), a_min=0, a_max=999)
    return torch.from_numpy(tables['so3_exp_score_norms'][idx]).float()


def torus_score_norm(tables, sigma):
    """torus.score_norm (torus.py:79-83) on the float32 array score_model.py:306 passes in."""
    s = sigma.float().numpy() if torch.is_tensor(sigma) else np.asarray(sigma, dtype=np.float32)
    s = np.log(s / np.pi)
    s = (s - np.log(3e-3)) / (np.log(2) - np.log(3e-3)) * 5000
    idx = np.round(np.clip(s, 0, 5000)).astype(int)
    return torch.from_numpy(tables['torus_score_norm'][idx]).float()


# ------------------------------------------------------------------------------------------------ update
def axis_angle_to_matrix(aa):
    """geometry.py:38-85 (pytorch3d axis-angle -> quaternion -> matrix)."""
    ang = aa.norm(dim=-1, keepdim=True)
    half = 0.5 * ang
    small = ang.abs() < 1e-6
    k = torch.where(small, 0.5 - ang * ang / 48, torch.sin(half) / torch.where(small, torch.ones_like(ang), ang))
    q = torch.cat([torch.cos(half), aa * k], -1)
    r, i, j, kk = q.unbind(-1)
    two_s = 2.0 / (q * q).sum(-1)
    m = torch.stack([1 - two_s * (j * j + kk * kk), two_s * (i * j - kk * r), two_s * (i * kk + j * r),
                     two_s * (i * j + kk * r), 1 - two_s * (i * i + kk * kk), two_s * (j * kk - i * r),
                     two_s * (i * kk - j * r), two_s * (j * kk + i * r), 1 - two_s * (i * i + j * j)], -1)
    return m.reshape(aa.shape[:-1] + (3, 3))


def kabsch_batch(A, Bt):
    """rigid_transform_Kabsch_3D_torch_batch (geometry.py:126-156): R, t with R A + t ~ B."""
    A, Bt = A.permute(0, 2, 1), Bt.permute(0, 2, 1)
    ca, cb = A.mean(2, keepdim=True), Bt.mean(2, keepdim=True)
    H = torch.bmm(A - ca, (Bt - cb).transpose(1, 2))
    U, S, Vt = torch.linalg.svd(H)
    R = torch.bmm(Vt.transpose(1, 2), U.transpose(1, 2))
    SS = torch.diag(torch.tensor([1., 1., -1.]))
    Rm = torch.bmm(Vt.transpose(1, 2) @ SS, U.transpose(1, 2))
    R = torch.where(torch.linalg.det(R)[:, None, None] < 0, Rm, R)
    t = torch.bmm(-R, ca) + cb
    return R, t


def modify_conformer_batch(pos, B, rot_bonds, mask_rotate, tr_update, rot_update, tor_update):
    """modify_conformer_batch (diffusion_utils.py:37-55) with the sequential torsion loop of
    modify_conformer_torsion_angles_batch (torsion.py:71-86).  ``rot_bonds`` [R,2] (u,v) are local atom
    indices of one copy, ``mask_rotate`` bool [R,N]; every graph in the batch is a copy of one complex."""
    N = pos.shape[0] // B
    x = pos.reshape(B, N, 3) + 0
    c = x.mean(1, keepdim=True)
    Rm = axis_angle_to_matrix(rot_update)
    rigid = torch.bmm(x - c, Rm.permute(0, 2, 1)) + tr_update.unsqueeze(1) + c
    if tor_update is None:
        return rigid.reshape(-1, 3)
    tor_update = tor_update.reshape(B, -1)
    flex = rigid + 0
    for k in range(rot_bonds.shape[0]):
        u, v = int(rot_bonds[k, 0]), int(rot_bonds[k, 1])
        axis = flex[:, u] - flex[:, v]
        rm = axis_angle_to_matrix(axis / axis.norm(dim=-1, keepdim=True) * tor_update[:, k:k + 1])
        m = mask_rotate[k]
        flex[:, m] = torch.bmm(flex[:, m] - flex[:, v:v + 1], rm.transpose(1, 2)) + flex[:, v:v + 1]
    R, t = kabsch_batch(flex, rigid)
    return (torch.bmm(flex, R.transpose(1, 2)) + t.transpose(1, 2)).reshape(-1, 3)


def set_time(data, t_tr, t_rot, t_tor, B):
    """diffusion_utils.py:101-117."""
    for nt in ('ligand', 'receptor'):
        n = data[nt].num_nodes
        data[nt].node_t = {'tr': t_tr * torch.ones(n), 'rot': t_rot * torch.ones(n), 'tor': t_tor * torch.ones(n)}
    data.complex_t = {'tr': t_tr * torch.ones(B), 'rot': t_rot * torch.ones(B), 'tor': t_tor * torch.ones(B)}


def step_coefficients(cfg, t, dt, temp_sampling, temp_psi, temp_sigma_data, ode=False):
    """Scalar (float64) coefficients of one reverse step, sampling.py:111-192:
    perturb = a * score + b * z per component (tr, rot, tor)."""
    sig = t_to_sigma(cfg, t[0], t[1], t[2])
    rng = [(cfg.tr_sigma_min, cfg.tr_sigma_max), (cfg.rot_sigma_min, cfg.rot_sigma_max),
           (cfg.tor_sigma_min, cfg.tor_sigma_max)]
    a, b = [], []
    for i in range(3):
        g = sig[i] * np.sqrt(2 * np.log(rng[i][1] / rng[i][0]))
        if ode:
            a.append(0.5 * g ** 2 * dt[i]); b.append(0.0)
        elif temp_sampling[i] != 1.0:
            sd = np.exp(temp_sigma_data[i] * np.log(rng[i][1]) + (1 - temp_sigma_data[i]) * np.log(rng[i][0]))
            lam = (sd + sig[i]) / (sd + sig[i] / temp_sampling[i])
            a.append(g ** 2 * dt[i] * (lam + temp_sampling[i] * temp_psi[i] / 2))
            b.append(g * np.sqrt(dt[i] * (1 + temp_psi[i])))
        else:
            a.append(g ** 2 * dt[i]); b.append(g * np.sqrt(dt[i]))
    return a, b


def sample(p, cfg, batch, tables, schedule, noise, inference_steps=None, temp_sampling=(1.0, 1.0, 1.0),
           temp_psi=(0.0, 0.0, 0.0), temp_sigma_data=(0.5, 0.5, 0.5), ode=False, trajectory=None,
           step_callback=None, edge_log=None):
    """The reverse-diffusion loop of ``sampling()`` (sampling.py:105-198) for one batch of B copies of one
    complex, with *pre-drawn* noise so that the CPU oracle and the CUDA path consume identical z:
    ``noise = {'tr': [steps,B,3], 'rot': [steps,B,3], 'tor': [steps,B*R]}`` (the reference draws tr, rot, tor
    in that order with torch.normal on the device, :146-165; a zero row reproduces no_final_step_noise).
    Mutates and returns ``batch['ligand'].pos``."""
    steps = inference_steps or len(schedule)
    B = int(batch.num_graphs)
    lig = batch['ligand']
    M = batch['ligand', 'ligand'].edge_index.shape[1] // B
    mask0 = lig.edge_mask[:M]
    rot_bonds = batch['ligand', 'ligand'].edge_index[:, :M].T[mask0]
    mr = lig.mask_rotate
    while isinstance(mr, (list, tuple)):
        mr = mr[0]
    mask_rotate = torch.as_tensor(np.asarray(mr)).bool()
    for s in range(steps):
        t = schedule[s]
        dt = schedule[s] - schedule[s + 1] if s < steps - 1 else schedule[s]
        set_time(batch, t, t, t, B)
        if cfg.latent_droprate > 0:
            lig.unconditional = torch.zeros(lig.num_nodes, 1)
            batch['receptor'].unconditional = torch.zeros(batch['receptor'].num_nodes, 1)
        trace = {} if edge_log is not None else None
        tr_s, rot_s, tor_s = forward(p, cfg, batch, tables, trace)
        if edge_log is not None:
            edge_log.append(trace['n_edges'])                   # edges of the combined graph of this step (all poses)
        a, b = step_coefficients(cfg, (t, t, t), (dt, dt, dt), temp_sampling, temp_psi, temp_sigma_data, ode)
        tr_p = float(a[0]) * tr_s + float(b[0]) * noise['tr'][s]
        rot_p = float(a[1]) * rot_s + float(b[1]) * noise['rot'][s]
        tor_p = None
        if not cfg.no_torsion and tor_s.numel():
            tor_p = float(a[2]) * tor_s + float(b[2]) * noise['tor'][s]
        if step_callback is not None:
            step_callback(s, tr_s, rot_s, tor_s)
        lig.pos = modify_conformer_batch(lig.pos.float(), B, rot_bonds, mask_rotate, tr_p, rot_p, tor_p)
        if trajectory is not None:
            trajectory.append(lig.pos.clone())
    return lig.pos
