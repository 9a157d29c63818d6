"""TEST INFRASTRUCTURE ONLY -- pure-torch shims that let the reference's own model files import here.

The reference (``/root/reference``) is 100 % Python but its leaf ops live in third-party wheels that are
not installed in this image and cannot be installed offline: ``e3nn``, ``torch_cluster``,
``torch_scatter``, ``torch_geometric`` (no version is pinned anywhere in the reference; its README.md:9
defers to upstream DiffDock).  ``install()`` registers restatements of exactly the leaf ops the hot path
calls, after which ``models/score_model.py``, ``models/tensor_layers.py``, ``utils/diffusion_utils.py``,
``utils/torsion.py`` ... import *unmodified* from ``/root/reference`` (see ``oracle/ref_loader.py``).

Used only (a) in this container to validate ``oracle/restate.py`` against the reference's own control
flow and (b) by ``oracle/make_golden.py`` to generate ``tests/golden/*.npz``.  ``/root/reference`` does not
exist on the GPU box, so nothing at run time depends on this file.

Published semantics restated (each cites the reference call site it serves):
  * ``e3nn.o3.spherical_harmonics(..., normalize=True, normalization='component')``
    (score_model.py:295, 342, 371, 406, 422, 436): l=0: 1; l=1: sqrt3*(x,y,z); l=2: sqrt5*(sqrt3 xz,
    sqrt3 xy, y^2-(x^2+z^2)/2, sqrt3 yz, sqrt3/2 (z^2-x^2)).
  * ``e3nn.o3.FullyConnectedTensorProduct(shared_weights=False)`` (tensor_layers.py:137): 'uvw' paths
    enumerated in1-major, in2, out; per path a ``[mul1, mul2, mul_out]`` weight block; component
    irrep normalisation and 'element' path normalisation: sqrt(dim_out / sum_paths mul1*mul2).
  * ``e3nn.o3.FullTensorProduct`` (score_model.py:152): 'uvuv' paths, outputs sorted, path weight
    sqrt(2 l_out + 1).
  * Wigner 3j: C(0,l,l)=C(l,0,l)=C(l,l,0)=delta/sqrt(2l+1); all others are the constants e3nn embedded as
    buffers in the reference's shipped checkpoints (``oracle/w3j_constants.npz``, extracted from
    ``workdir/diffdockS_score_model/best_ema_inference_epoch_model.pt``).
  * ``e3nn.nn.BatchNorm`` eval mode (tensor_layers.py:145, 162).
  * ``torch_cluster.radius`` / ``radius_graph`` (score_model.py:315, 379-384, 430): CUDA-kernel order
    (query-major, candidates ascending, strict <, stop at max_num_neighbors).
  * ``torch_scatter.scatter(reduce='mean')`` (tensor_layers.py:159).
"""
from __future__ import annotations

import math
import os
import re
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import torch
from torch import nn

_W3J = None


def _w3j_table():
    global _W3J
    if _W3J is None:
        z = np.load(os.path.join(os.path.dirname(__file__), 'w3j_constants.npz'))
        _W3J = {k: torch.from_numpy(z[k]).float() for k in z.files}
    return _W3J


def wigner3j(l1, l2, l3):
    if l1 == 0 and l2 == l3:
        return (torch.eye(2 * l2 + 1) / math.sqrt(2 * l2 + 1)).reshape(1, 2 * l2 + 1, 2 * l2 + 1)
    if l2 == 0 and l1 == l3:
        return (torch.eye(2 * l1 + 1) / math.sqrt(2 * l1 + 1)).reshape(2 * l1 + 1, 1, 2 * l1 + 1)
    if l3 == 0 and l1 == l2:
        return (torch.eye(2 * l1 + 1) / math.sqrt(2 * l1 + 1)).reshape(2 * l1 + 1, 2 * l1 + 1, 1)
    return _w3j_table()[f'_w3j_{l1}_{l2}_{l3}'].clone()


class Irrep(tuple):
    def __new__(cls, l, p=None):
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, str):
                s = l.strip()
                l, p = int(s[:-1]), {'e': 1, 'o': -1}[s[-1]]
            else:
                l, p = l
        return tuple.__new__(cls, (int(l), int(p)))

    l = property(lambda self: self[0])
    p = property(lambda self: self[1])
    dim = property(lambda self: 2 * self[0] + 1)

    def __str__(self):
        return f"{self[0]}{'e' if self[1] == 1 else 'o'}"

    __repr__ = __str__

    def sort_key(self):
        return (self[0], -self[1] * (-1) ** self[0])


class _MulIr(tuple):
    def __new__(cls, mul, ir):
        return tuple.__new__(cls, (int(mul), Irrep(ir)))

    mul = property(lambda self: self[0])
    ir = property(lambda self: self[1])
    dim = property(lambda self: self[0] * self[1].dim)


class Irreps(tuple):
    def __new__(cls, irreps):
        if isinstance(irreps, Irreps):
            return irreps
        out = []
        if isinstance(irreps, str):
            for part in irreps.split('+'):
                part = part.strip()
                if not part:
                    continue
                if 'x' in part:
                    m, ir = part.split('x')
                    out.append(_MulIr(int(m), Irrep(ir)))
                else:
                    out.append(_MulIr(1, Irrep(part)))
        else:
            for mul, ir in irreps:
                out.append(_MulIr(mul, Irrep(ir)))
        return tuple.__new__(cls, out)

    @staticmethod
    def spherical_harmonics(lmax, p=-1):
        return Irreps([(1, (l, p ** l)) for l in range(lmax + 1)])

    def slices(self):
        s, i = [], 0
        for mul, ir in self:
            s.append(slice(i, i + mul * ir.dim))
            i += mul * ir.dim
        return s

    @property
    def dim(self):
        return sum(mul * ir.dim for mul, ir in self)

    def __str__(self):
        return '+'.join(f'{mul}x{ir}' for mul, ir in self)

    __repr__ = __str__


def spherical_harmonics(l, x, normalize, normalization='integral'):
    assert normalize and normalization == 'component'
    if isinstance(l, (str, Irreps)):
        ls = [ir.l for _, ir in Irreps(l)]
    elif isinstance(l, int):
        ls = [l]
    else:
        ls = list(l)
    x = torch.nn.functional.normalize(x, dim=-1)
    X, Y, Z = x[..., 0], x[..., 1], x[..., 2]
    out = []
    for li in ls:
        if li == 0:
            out.append(torch.ones_like(X).unsqueeze(-1))
        elif li == 1:
            out.append(math.sqrt(3) * torch.stack([X, Y, Z], -1))
        elif li == 2:
            s3 = math.sqrt(3)
            out.append(math.sqrt(5) * torch.stack(
                [s3 * X * Z, s3 * X * Y, Y * Y - 0.5 * (X * X + Z * Z), s3 * Y * Z, s3 / 2 * (Z * Z - X * X)], -1))
        else:
            raise NotImplementedError(li)
    return torch.cat(out, -1)


class _Compiled(nn.Module):
    pass


class _TPBase(nn.Module):
    def _finish(self, instr, specialised=True):
        self.instructions = instr
        self.register_buffer('weight', torch.zeros(0))
        self.register_buffer('output_mask', torch.ones(self.irreps_out.dim))
        self._compiled_main_left_right = _Compiled()
        for (i1, i2, io) in instr:
            l1, l2, l3 = self.irreps_in1[i1].ir.l, self.irreps_in2[i2].ir.l, self.irreps_out[io].ir.l
            if specialised and 0 in (l1, l2, l3):
                continue    # e3nn's 'uvw' codegen special-cases l=0 paths (no buffer in the checkpoints)
            name = f'_w3j_{l1}_{l2}_{l3}'
            if not hasattr(self._compiled_main_left_right, name):
                self._compiled_main_left_right.register_buffer(name, wigner3j(l1, l2, l3))

    def _c(self, l1, l2, l3):
        name = f'_w3j_{l1}_{l2}_{l3}'
        if hasattr(self._compiled_main_left_right, name):
            return getattr(self._compiled_main_left_right, name)
        return wigner3j(l1, l2, l3)


class FullyConnectedTensorProduct(_TPBase):
    def __init__(self, irreps_in1, irreps_in2, irreps_out, shared_weights=None, internal_weights=None, **kw):
        super().__init__()
        assert shared_weights is False
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        instr = []
        for i1, (m1, ir1) in enumerate(self.irreps_in1):
            for i2, (m2, ir2) in enumerate(self.irreps_in2):
                for io, (mo, iro) in enumerate(self.irreps_out):
                    if abs(ir1.l - ir2.l) <= iro.l <= ir1.l + ir2.l and iro.p == ir1.p * ir2.p:
                        instr.append((i1, i2, io))
        self.weight_numel = sum(self.irreps_in1[a].mul * self.irreps_in2[b].mul * self.irreps_out[c].mul for a, b, c in instr)
        self._finish(instr)

    def forward(self, x1, x2, weight):
        s1, s2, so = self.irreps_in1.slices(), self.irreps_in2.slices(), self.irreps_out.slices()
        Z = x1.shape[0]
        out = [x1.new_zeros(Z, mo, iro.dim) for mo, iro in self.irreps_out]
        fan = [0] * len(self.irreps_out)
        for (i1, i2, io) in self.instructions:
            fan[io] += self.irreps_in1[i1].mul * self.irreps_in2[i2].mul
        off = 0
        for (i1, i2, io) in self.instructions:
            (m1, ir1), (m2, ir2), (mo, iro) = self.irreps_in1[i1], self.irreps_in2[i2], self.irreps_out[io]
            n = m1 * m2 * mo
            w = weight[:, off:off + n].reshape(Z, m1, m2, mo)
            off += n
            a = x1[:, s1[i1]].reshape(Z, m1, ir1.dim)
            b = x2[:, s2[i2]].reshape(Z, m2, ir2.dim)
            C = self._c(ir1.l, ir2.l, iro.l).to(x1.dtype)
            pw = math.sqrt(iro.dim / fan[io])
            out[io] = out[io] + pw * torch.einsum('zuvw,ijk,zui,zvj->zwk', w, C, a, b)
        return torch.cat([o.reshape(Z, -1) for o in out], -1)


class FullTensorProduct(_TPBase):
    def __init__(self, irreps_in1, irreps_in2, **kw):
        super().__init__()
        self.irreps_in1, self.irreps_in2 = Irreps(irreps_in1), Irreps(irreps_in2)
        outs = []
        for i1, (m1, ir1) in enumerate(self.irreps_in1):
            for i2, (m2, ir2) in enumerate(self.irreps_in2):
                for l in range(abs(ir1.l - ir2.l), ir1.l + ir2.l + 1):
                    outs.append((m1 * m2, Irrep(l, ir1.p * ir2.p), i1, i2))
        order = sorted(range(len(outs)), key=lambda i: outs[i][1].sort_key())
        self.irreps_out = Irreps([(outs[i][0], outs[i][1]) for i in order])
        instr = [(outs[i][2], outs[i][3], pos) for pos, i in enumerate(order)]
        self._finish(instr, specialised=False)

    def forward(self, x1, x2):
        s1, s2 = self.irreps_in1.slices(), self.irreps_in2.slices()
        Z = x1.shape[0]
        res = [None] * len(self.irreps_out)
        for (i1, i2, io) in self.instructions:
            (m1, ir1), (m2, ir2), (mo, iro) = self.irreps_in1[i1], self.irreps_in2[i2], self.irreps_out[io]
            a = x1[:, s1[i1]].reshape(Z, m1, ir1.dim)
            b = x2[:, s2[i2]].reshape(Z, m2, ir2.dim)
            C = self._c(ir1.l, ir2.l, iro.l).to(x1.dtype)
            res[io] = math.sqrt(iro.dim) * torch.einsum('ijk,zui,zvj->zuvk', C, a, b).reshape(Z, -1)
        return torch.cat(res, -1)


class BatchNorm(nn.Module):
    """e3nn.nn.BatchNorm, eval-mode arithmetic only (affine=True, eps=1e-5, normalization='component')."""

    def __init__(self, irreps, eps=1e-5, **kw):
        super().__init__()
        self.irreps = Irreps(irreps)
        self.eps = eps
        ns = sum(mul for mul, ir in self.irreps if ir.l == 0 and ir.p == 1)
        nf = sum(mul for mul, ir in self.irreps)
        self.register_buffer('running_mean', torch.zeros(ns))
        self.register_buffer('running_var', torch.ones(nf))
        self.weight = nn.Parameter(torch.ones(nf))
        self.bias = nn.Parameter(torch.zeros(ns))

    def forward(self, x):
        assert not self.training, 'oracle shim implements eval mode only'
        out, ix, iw, ib = [], 0, 0, 0
        for mul, ir in self.irreps:
            d = ir.dim
            f = x[:, ix:ix + mul * d].reshape(-1, mul, d)
            ix += mul * d
            if ir.l == 0 and ir.p == 1:
                f = f - self.running_mean[ib:ib + mul].reshape(1, mul, 1)
            f = f * (self.running_var[iw:iw + mul] + self.eps).pow(-0.5).reshape(1, mul, 1)
            f = f * self.weight[iw:iw + mul].reshape(1, mul, 1)
            if ir.l == 0 and ir.p == 1:
                f = f + self.bias[ib:ib + mul].reshape(1, mul, 1)
                ib += mul
            iw += mul
            out.append(f.reshape(-1, mul * d))
        return torch.cat(out, -1)


def scatter(src, index, dim=0, dim_size=None, reduce='sum', out=None):
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max()) + 1
    if torch.is_tensor(dim_size):
        dim_size = int(dim_size)
    res = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    res.index_add_(0, index, src)
    if reduce == 'mean':
        cnt = torch.bincount(index, minlength=dim_size).clamp(min=1).to(src.dtype)
        res = res / cnt.reshape(-1, *([1] * (src.dim() - 1)))
    elif reduce not in ('sum', 'add'):
        raise NotImplementedError(reduce)
    return res


def scatter_mean(src, index, dim=0, dim_size=None, out=None):
    return scatter(src, index, dim, dim_size, 'mean')


def pair_dist2(xq, xc):
    """Unfused fp32 ((dx*dx + dy*dy) + dz*dz) -- the product kernels round identically (csrc/ddk_geom.cuh)."""
    d = xc.unsqueeze(0) - xq.unsqueeze(1)
    sq = d * d
    return (sq[..., 0] + sq[..., 1]) + sq[..., 2]


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, num_workers=1):
    """row0 = query (y) index, row1 = candidate (x) index; CUDA-kernel order and truncation."""
    if batch_x is None:
        batch_x = torch.zeros(len(x), dtype=torch.long, device=x.device)
    if batch_y is None:
        batch_y = torch.zeros(len(y), dtype=torch.long, device=y.device)
    rows, cols = [], []
    nb = int(max(batch_x.max(), batch_y.max())) + 1 if len(batch_x) and len(batch_y) else 0
    r2 = torch.tensor(float(r) * float(r), dtype=x.dtype)
    for b in range(nb):
        iy = torch.nonzero(batch_y == b).flatten()
        ix = torch.nonzero(batch_x == b).flatten()
        if len(iy) == 0 or len(ix) == 0:
            continue
        within = pair_dist2(y[iy], x[ix]) < r2
        rank = torch.cumsum(within.long(), 1) - 1
        keep = within & (rank < max_num_neighbors)
        q, c = torch.nonzero(keep, as_tuple=True)
        rows.append(iy[q])
        cols.append(ix[c])
    if not rows:
        return torch.zeros(2, 0, dtype=torch.long, device=x.device)
    return torch.stack([torch.cat(rows), torch.cat(cols)])


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow='source_to_target', num_workers=1):
    ei = radius(x, x, r, batch, batch, max_num_neighbors if loop else max_num_neighbors + 1)
    if flow == 'source_to_target':
        row, col = ei[1], ei[0]
    else:
        row, col = ei[0], ei[1]
    if not loop:
        m = row != col
        row, col = row[m], col[m]
    return torch.stack([row, col])


_INSTALLED = False


def install():
    """Register the shim modules in ``sys.modules`` (idempotent)."""
    global _INSTALLED
    if _INSTALLED:
        return
    _INSTALLED = True
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    from disco_diffdock_b200 import data as ddata

    for name in ['rdkit', 'rdkit.Chem', 'rdkit.Chem.rdchem', 'rdkit.Chem.AllChem', 'rdkit.Chem.rdMolTransforms',
                 'rdkit.Chem.rdmolfiles', 'rdkit.Geometry', 'rdkit.RDLogger', 'Bio', 'Bio.PDB', 'Bio.PDB.PDBExceptions',
                 'spyrmsd', 'biopandas', 'biopandas.pdb', 'esm', 'wandb', 'torch_geometric.transforms',
                 'torch_geometric.nn', 'torch_geometric.nn.data_parallel', 'torch_geometric.utils']:
        if name not in sys.modules:
            sys.modules[name] = MagicMock(name=name)

    e3nn = types.ModuleType('e3nn')
    o3 = types.ModuleType('e3nn.o3')
    o3.Irrep, o3.Irreps = Irrep, Irreps
    o3.spherical_harmonics = spherical_harmonics
    o3.FullyConnectedTensorProduct = FullyConnectedTensorProduct
    o3.FullTensorProduct = FullTensorProduct
    e3nn_nn = types.ModuleType('e3nn.nn')
    e3nn_nn.BatchNorm = BatchNorm
    e3nn.o3, e3nn.nn = o3, e3nn_nn
    sys.modules.update({'e3nn': e3nn, 'e3nn.o3': o3, 'e3nn.nn': e3nn_nn})

    ts = types.ModuleType('torch_scatter')
    ts.scatter, ts.scatter_mean = scatter, scatter_mean
    tc = types.ModuleType('torch_cluster')
    tc.radius, tc.radius_graph = radius, radius_graph
    sys.modules.update({'torch_scatter': ts, 'torch_cluster': tc})

    tg = types.ModuleType('torch_geometric')
    tgd = types.ModuleType('torch_geometric.data')
    tgd.HeteroData, tgd.Batch, tgd.Data = ddata.HeteroData, ddata.Batch, ddata.HeteroData
    tgd.Dataset = object
    tgl = types.ModuleType('torch_geometric.loader')
    tgl.DataLoader = ddata.DataLoader
    tgl.DataListLoader = ddata.DataLoader
    tg.data, tg.loader = tgd, tgl
    sys.modules.update({'torch_geometric': tg, 'torch_geometric.data': tgd, 'torch_geometric.loader': tgl})
