"""CPU emulation of the re-associated conv layer exactly as csrc/ computes it (basis layout, packed second-layer
weights, folded batch-norm), checked against the oracle's per-edge formulation.  Validates weights.py and the
algebra without a GPU."""
import numpy as np
import torch

from disco_diffdock_b200 import data as ddata
from disco_diffdock_b200 import weights
from oracle import restate
from tests import helpers
from tests.test_oracle_golden import load_tables


def raw_basis(x, sh, layer):
    """basis functions in kernel order [0e | 1o comp-major | 1e comp-major | 0o], no constant factors."""
    lv = min(layer, 3)
    x0e, x1o, x1e, x0o = x[:, 0:24], x[:, 24:42].reshape(-1, 6, 3), x[:, 42:60].reshape(-1, 6, 3), x[:, 60:84]
    sh0, s = sh[:, 0:1], sh[:, 1:4]

    def cross(a):
        return torch.linalg.cross(a, s[:, None, :].expand_as(a), dim=-1)

    def dot(a):
        return (a * s[:, None, :]).sum(-1)
    c0e = [x0e * sh0] + ([dot(x1o)] if lv >= 1 else [])
    c1o = [x0e[:, :, None] * s[:, None, :]] + ([x1o * sh0[:, :, None]] if lv >= 1 else []) + ([cross(x1e)] if lv >= 2 else [])
    c1e = ([cross(x1o)] if lv >= 1 else []) + ([x1e * sh0[:, :, None]] if lv >= 2 else []) + \
          ([x0o[:, :, None] * s[:, None, :]] if lv >= 3 else [])
    c0o = ([dot(x1e)] if lv >= 2 else []) + ([x0o * sh0] if lv >= 3 else [])
    out = [torch.cat(c0e, 1)]
    for cl in (c1o, c1e):
        if cl:
            v = torch.cat(cl, 1)                       # [E, F, 3]
            out.append(v.permute(0, 2, 1).reshape(len(x), -1))   # comp-major
    if c0o:
        out.append(torch.cat(c0o, 1))
    return torch.cat(out, 1)


def test_reassociated_layer_matches_oracle():
    m, sd, cfg = helpers.make_model(0)
    blob, off = weights.pack_weights(sd, m.hyper())
    blob = torch.from_numpy(blob)
    E = weights.ENUMS
    _, lst = helpers.make_pose_batch(3, 12, 30, 2)
    batch = ddata.Batch.from_data_list(lst)
    restate.set_time(batch, 0.4, 0.4, 0.4, 2)
    tr = {}
    with torch.no_grad():
        restate.forward(sd, cfg, batch, load_tables(), tr)
    nl = len(tr['lig_h0'])
    x = torch.zeros(nl + len(tr['rec_h0']), 84)
    x[:nl, :24], x[nl:, :24] = tr['lig_h0'], tr['rec_h0']
    rr = batch['receptor', 'receptor'].edge_index
    src = torch.cat([tr['ll_src'], tr['lr_src'], rr[0] + nl, tr['lr_dst'] + nl])
    dst = torch.cat([tr['ll_dst'], tr['lr_dst'] + nl, rr[1] + nl, tr['lr_src']])
    ea = torch.cat([tr['ll_ea'], tr['lr_ea'], tr['rr_ea'], tr['lr_ea']])
    sh = torch.cat([tr['ll_sh'], tr['lr_sh'], tr['rr_sh'], tr['lr_sh']])
    grp = torch.cat([torch.full((len(a),), g) for g, a in enumerate([tr['ll_src'], tr['lr_src'], rr[0], tr['lr_dst']])])
    for layer in range(cfg.num_conv_layers):
        table, U = weights.class_table(layer)
        base = E['DDK_W_CONV_BASE'] + layer * E['DDK_W_CONV_STRIDE']
        out_sum = torch.zeros(len(x), 84, dtype=torch.float64)
        for g in range(4):
            sel = grp == g
            W1 = blob[off[base + E['DDK_WL_W1'] + g]:][:72 * 72].reshape(72, 72)
            b1 = blob[off[base + E['DDK_WL_B1'] + g]:][:72]
            h = torch.relu(ea[sel] @ W1[:, :24].T + (x[src[sel], :24] @ W1[:, 24:48].T + b1) + x[dst[sel], :24] @ W1[:, 48:].T)
            Bm = raw_basis(x[dst[sel]], sh[sel], layer)
            assert Bm.shape[1] == U
            A = torch.zeros(len(x), U, 72, dtype=torch.float64).index_add_(0, src[sel], (Bm[:, :, None] * h[:, None, :]).double())
            Bs = torch.zeros(len(x), U, dtype=torch.float64).index_add_(0, src[sel], Bm.double())
            wp = blob[off[base + E['DDK_WL_W2P'] + g]:].double()
            bp = blob[off[base + E['DDK_WL_B2P'] + g]:].double()
            wo = bo = 0
            for c in table:
                F, O, nc, uo = c['F'], c['O'], c['ncomp'], c['uoff']
                if F == 0 or O == 0:
                    continue
                Wk = wp[wo:wo + F * 72 * O].reshape(F * 72, O)
                bk = bp[bo:bo + F * O].reshape(F, O)
                wo += F * 72 * O
                bo += F * O
                col0 = {'0e': 0, '1o': 24, '1e': 42, '0o': 60}[c['key']]
                for comp in range(nc):
                    Ak = A[:, uo + comp * F: uo + (comp + 1) * F, :].reshape(len(x), F * 72)
                    res = Ak @ Wk + Bs[:, uo + comp * F: uo + (comp + 1) * F] @ bk
                    cols = col0 + (torch.arange(O) * nc + comp if nc == 3 else torch.arange(O))
                    out_sum[:, cols] += res
        cnt = torch.bincount(src, minlength=len(x)).clamp(min=1).double()
        sc = blob[off[base + E['DDK_WL_BN_SCALE']]:][:84].double()
        sf = blob[off[base + E['DDK_WL_BN_SHIFT']]:][:84].double()
        x = ((out_sum / cnt[:, None]) * sc + sf + x.double()).float()
        ref = tr[f'x{layer + 1}']
        assert float((x[:, :ref.shape[1]] - ref).abs().max()) < 2e-5, layer
        if ref.shape[1] < 84:
            assert float(x[:, ref.shape[1]:].abs().max()) == 0.0
