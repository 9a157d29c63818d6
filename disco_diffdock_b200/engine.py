"""ctypes binding of ``libddk`` (include/ddk.h) and the host-side glue between PyG-style batches and the C ABI.

PyTorch is used here for device memory, streams and index bookkeeping only; every number on the hot path is
produced by the CUDA kernels in ``csrc/``.  There is no CPU fallback: if the shared library or a CUDA device is
missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from types import SimpleNamespace
from typing import Dict, List, Optional

import numpy as np
import torch

from . import weights as W

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libddk.so')
_lib = None


class DdkConfig(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('ns', C.c_int32), ('nv', C.c_int32), ('num_conv_layers', C.c_int32),
                ('latent_dim', C.c_int32), ('has_unconditional', C.c_int32), ('dynamic_max_cross', C.c_int32),
                ('scale_by_sigma', C.c_int32), ('no_torsion', C.c_int32),
                ('lig_max_radius', C.c_float), ('rec_max_radius', C.c_float), ('cross_max_distance', C.c_float),
                ('center_max_distance', C.c_float), ('scratch_bytes', C.c_int64)]


class DdkBatch(C.Structure):
    _fields_ = [('B', C.c_int32), ('NL', C.c_int32), ('NR', C.c_int32), ('EB', C.c_int32), ('ER', C.c_int32),
                ('RB', C.c_int32),
                ('lig_ptr_h', C.c_void_p), ('rec_ptr_h', C.c_void_p), ('bond_index_h', C.c_void_p),
                ('bond_ptr_h', C.c_void_p), ('edge_mask_h', C.c_void_p), ('rec_index_h', C.c_void_p),
                ('rec_edge_ptr_h', C.c_void_p), ('mask_rotate_off_h', C.c_void_p),
                ('lig_x', C.c_void_p), ('bond_attr', C.c_void_p), ('rec_x', C.c_void_p), ('rec_pos', C.c_void_p),
                ('mask_rotate', C.c_void_p), ('lig_latent', C.c_void_p), ('rec_latent', C.c_void_p),
                ('lig_uncond', C.c_void_p), ('rec_uncond', C.c_void_p)]


class DdkStepInputs(C.Structure):
    _fields_ = [('sigma_emb', C.c_void_p), ('cross_cutoff', C.c_void_p), ('tr_sigma', C.c_void_p),
                ('rot_scale', C.c_void_p), ('tor_scale', C.c_void_p)]


class DdkStepCoef(C.Structure):
    _fields_ = [('a_tr', C.c_float), ('b_tr', C.c_float), ('a_rot', C.c_float), ('b_rot', C.c_float),
                ('a_tor', C.c_float), ('b_tor', C.c_float)]


EXPORTS = ['ddk_abi_version', 'ddk_create', 'ddk_destroy', 'ddk_last_error', 'ddk_set_batch', 'ddk_score', 'ddk_embed',
           'ddk_get_node_features', 'ddk_update', 'ddk_sample', 'ddk_sample_host', 'ddk_kernel_launches',
           'ddk_last_edge_count', 'ddk_debug_read', 'ddk_host_kabsch', 'ddk_host_axis_angle_to_matrix', 'ddk_host_lane_tables_check', 'ddk_host_tc_rows_eval', 'ddk_host_tc_split',
           'ddk_profile_enable', 'ddk_profile_read', 'ddk_debug_set_tc', 'ddk_host_tcr_roles_check', 'ddk_host_tc_split_rn', 'ddk_edge_total', 'ddk_segment_total', 'ddk_group_totals']


def load_library(path: Optional[str] = None):
    """Load libddk.so (built in-tree by ``build.py``).  Raises if it is missing -- there is no fallback path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or _LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f'{p} not found: build it with `python -m disco_diffdock_b200.build` '
                           '(the CUDA extension is mandatory; there is no CPU fallback)')
    lib = C.CDLL(p)
    lib.ddk_last_error.restype = C.c_char_p
    lib.ddk_last_error.argtypes = [C.c_void_p]
    lib.ddk_create.argtypes = [C.POINTER(DdkConfig), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int32, C.c_int32,
                               C.POINTER(C.c_void_p)]
    lib.ddk_destroy.argtypes = [C.c_void_p]
    lib.ddk_set_batch.argtypes = [C.c_void_p, C.POINTER(DdkBatch), C.c_void_p]
    lib.ddk_score.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(DdkStepInputs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ddk_embed.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(DdkStepInputs), C.c_void_p]
    lib.ddk_get_node_features.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ddk_update.argtypes = [C.c_void_p] + [C.c_void_p] * 7 + [C.POINTER(DdkStepCoef), C.c_void_p]
    lib.ddk_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(DdkStepInputs), C.c_void_p, C.c_void_p,
                               C.c_void_p, C.POINTER(DdkStepCoef), C.c_void_p]
    lib.ddk_sample_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(DdkStepInputs), C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.POINTER(DdkStepCoef)]
    lib.ddk_kernel_launches.restype = C.c_int64
    lib.ddk_kernel_launches.argtypes = [C.c_void_p]
    lib.ddk_last_edge_count.restype = C.c_int64
    lib.ddk_last_edge_count.argtypes = [C.c_void_p]
    lib.ddk_edge_total.restype = C.c_int64
    lib.ddk_edge_total.argtypes = [C.c_void_p]
    lib.ddk_segment_total.restype = C.c_int64
    lib.ddk_segment_total.argtypes = [C.c_void_p]
    lib.ddk_group_totals.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ddk_debug_read.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.ddk_debug_set_tc.argtypes = [C.c_int32]
    lib.ddk_profile_enable.argtypes = [C.c_void_p, C.c_int32]
    lib.ddk_profile_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ddk_host_kabsch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.ddk_host_axis_angle_to_matrix.argtypes = [C.c_void_p, C.c_void_p]
    if path is None:
        _lib = lib
    return lib


def set_tensor_core_path(mode) -> int:
    """Run-time choice of the conv kernels (None: follow DDK_TC; 0 FFMA2 only, 1 FFMA2 + k_acc_tc, 2 k_conv_tcr); effective
    from the next ``set_batch``."""
    return int(load_library().ddk_debug_set_tc(-1 if mode is None else int(mode)))


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _unwrap_mask(m):
    while isinstance(m, (list, tuple)):
        m = m[0]
    return np.ascontiguousarray(np.asarray(m), dtype=np.uint8)


def batch_index_arrays(data, no_torsion: bool):
    """Host index arrays of a PyG-style batch (utils/sampling.py:56-67): per-graph node / edge offsets, int32 edge lists,
    the rotatable-bond mask and the flattened ``mask_rotate`` blocks (identical blocks stored once)."""
    lig, rec = data['ligand'], data['receptor']
    B = int(data.num_graphs)
    lb = lig.batch.cpu().long() if 'batch' in lig else torch.zeros(lig.num_nodes, dtype=torch.long)
    rb = rec.batch.cpu().long() if 'batch' in rec else torch.zeros(rec.num_nodes, dtype=torch.long)
    lig_ptr = np.concatenate([[0], np.cumsum(np.bincount(lb.numpy(), minlength=B))]).astype(np.int32)
    rec_ptr = np.concatenate([[0], np.cumsum(np.bincount(rb.numpy(), minlength=B))]).astype(np.int32)
    bei = data['ligand', 'ligand'].edge_index.cpu().long()
    rei = data['receptor', 'receptor'].edge_index.cpu().long()
    bond_index = np.ascontiguousarray(bei.numpy().astype(np.int32))
    rec_index = np.ascontiguousarray(rei.numpy().astype(np.int32))
    bond_ptr = np.concatenate([[0], np.cumsum(np.bincount(lb[bei[0]].numpy(), minlength=B))]).astype(np.int32)
    rec_eptr = np.concatenate([[0], np.cumsum(np.bincount(rb[rei[0]].numpy(), minlength=B))]).astype(np.int32)
    edge_mask = np.ascontiguousarray(lig.edge_mask.cpu().numpy().astype(np.uint8))
    RB = int(edge_mask.sum())
    # mask_rotate: list (one entry per graph, possibly nested) of [R, N] arrays; identical arrays are stored once
    mr_off = np.zeros(B, dtype=np.int64)
    mr_flat = None
    if RB > 0 and not no_torsion:
        mr = lig.mask_rotate if 'mask_rotate' in lig else None
        if mr is None:
            raise RuntimeError('batch has rotatable bonds but no mask_rotate')
        per_graph = [mr[g] for g in range(B)] if isinstance(mr, (list, tuple)) and len(mr) == B and B > 1 else [mr] * B
        chunks, off = [], 0
        for g in range(B):
            m = _unwrap_mask(per_graph[g])
            nl = int(lig_ptr[g + 1] - lig_ptr[g])
            rg = int(edge_mask[bond_ptr[g]:bond_ptr[g + 1]].sum())
            if m.shape != (rg, nl):
                raise RuntimeError(f'mask_rotate of graph {g} has shape {m.shape}, expected {(rg, nl)}')
            found = None
            for (o, mm) in chunks:      # dedupe equal content (deep copies of one complex)
                if mm.shape == m.shape and np.array_equal(mm, m):
                    found = o
                    break
            if found is None:
                chunks.append((off, m))
                found = off
                off += m.size
            mr_off[g] = found
        mr_flat = np.concatenate([m.ravel() for _, m in chunks]) if chunks else np.zeros(1, np.uint8)
    host = SimpleNamespace(lig_ptr=lig_ptr, rec_ptr=rec_ptr, bond_index=bond_index, bond_ptr=bond_ptr,
                           edge_mask=edge_mask, rec_index=rec_index, rec_eptr=rec_eptr, mr_off=mr_off)
    return host, mr_flat, RB


def group_index_arrays(groups, no_torsion: bool):
    """The same arrays as ``batch_index_arrays`` for a batch given as runs of copies ``[(complex, n_copies), ...]``, built
    from one graph per run (no PyG collation of the copies); the big edge lists are written once, as int32."""
    protos = []
    for proto, n in groups:
        lig, rec = proto['ligand'], proto['receptor']
        bei = proto['ligand', 'ligand'].edge_index.cpu().numpy().astype(np.int32)
        rei = proto['receptor', 'receptor'].edge_index.cpu().numpy().astype(np.int32)
        em1 = lig.edge_mask.cpu().numpy().astype(np.uint8)
        protos.append((int(n), int(lig.num_nodes), int(rec.num_nodes), bei, rei, em1, lig))
    B = sum(p[0] for p in protos)
    EB = sum(p[0] * p[3].shape[1] for p in protos)
    ER = sum(p[0] * p[4].shape[1] for p in protos)
    if max(sum(p[0] * p[1] for p in protos), sum(p[0] * p[2] for p in protos)) >= 2 ** 31:
        raise RuntimeError('batch too large for 32-bit node indices')
    bond_index, rec_index = np.empty((2, EB), np.int32), np.empty((2, ER), np.int32)
    edge_mask, mr_off = np.empty(EB, np.uint8), np.zeros(B, np.int64)
    counts = np.empty((4, B), np.int64)                       # ligand atoms, residues, bonds, contacts per graph
    mr_chunks = []
    lo = ro = mo = RB = g0 = e0 = r0 = 0
    for n, nl, nr, bei, rei, em1, lig in protos:
        eb, er = bei.shape[1], rei.shape[1]
        counts[:, g0:g0 + n] = np.array([[nl], [nr], [eb], [er]])
        ar = np.arange(n, dtype=np.int32)
        np.add(bei[:, None, :], (lo + ar * nl)[None, :, None], out=bond_index[:, e0:e0 + n * eb].reshape(2, n, eb))
        np.add(rei[:, None, :], (ro + ar * nr)[None, :, None], out=rec_index[:, r0:r0 + n * er].reshape(2, n, er))
        edge_mask[e0:e0 + n * eb].reshape(n, eb)[:] = em1[None, :]
        r1 = int(em1.sum())
        RB += r1 * n
        if r1 > 0 and not no_torsion:
            m = _unwrap_mask(lig.mask_rotate).ravel()
            if m.size != r1 * nl:
                raise RuntimeError('mask_rotate does not match edge_mask / ligand size')
            mr_chunks.append(m)
            mr_off[g0:g0 + n] = mo
            mo += m.size
        lo += n * nl; ro += n * nr; g0 += n; e0 += n * eb; r0 += n * er
    cum = lambda c: np.concatenate([[0], np.cumsum(c)]).astype(np.int32)
    host = SimpleNamespace(lig_ptr=cum(counts[0]), rec_ptr=cum(counts[1]), bond_ptr=cum(counts[2]), rec_eptr=cum(counts[3]),
                           bond_index=bond_index, rec_index=rec_index, edge_mask=edge_mask, mr_off=mr_off)
    mr_flat = np.concatenate(mr_chunks) if mr_chunks else None
    return host, mr_flat, RB


class Engine:
    """One ``DdkCtx``: packed weights on one GPU plus the workspace of the current batch."""

    def __init__(self, hyper: SimpleNamespace, state_dict: Dict[str, torch.Tensor], device: torch.device,
                 scratch_bytes: int = 0):
        self.lib = load_library()
        device = torch.device(device)
        if device.type != 'cuda' or not torch.cuda.is_available():
            raise RuntimeError('disco_diffdock_b200 needs a CUDA device (sm_100a); there is no CPU execution path')
        self.device = torch.device('cuda', device.index if device.index is not None else torch.cuda.current_device())
        self.hyper = hyper
        blob, offsets = W.pack_weights(state_dict, hyper)
        cfg = DdkConfig(abi_version=W.ENUMS['DDK_ABI_VERSION'], ns=hyper.ns, nv=hyper.nv,
                        num_conv_layers=hyper.num_conv_layers, latent_dim=hyper.latent_dim,
                        has_unconditional=int(hyper.latent_droprate > 0), dynamic_max_cross=int(hyper.dynamic_max_cross),
                        scale_by_sigma=int(hyper.scale_by_sigma), no_torsion=int(hyper.no_torsion),
                        lig_max_radius=hyper.lig_max_radius, rec_max_radius=hyper.rec_max_radius,
                        cross_max_distance=hyper.cross_max_distance, center_max_distance=hyper.center_max_distance,
                        scratch_bytes=int(scratch_bytes))
        ctx = C.c_void_p()
        rc = self.lib.ddk_create(C.byref(cfg), _np_ptr(blob), blob.size, _np_ptr(offsets), len(offsets),
                                 self.device.index, C.byref(ctx))
        if rc != 0:
            raise RuntimeError(f'ddk_create failed ({rc}): {self.lib.ddk_last_error(None).decode()}')
        self.ctx = ctx
        self.weight_bytes = blob.nbytes
        self._keep: List = []
        self.batch_info: Optional[SimpleNamespace] = None

    def __del__(self):
        try:
            if getattr(self, 'ctx', None):
                self.lib.ddk_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f'{what} failed ({rc}): {self.lib.ddk_last_error(self.ctx).decode()}')

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------------------------------ batch
    def set_batch(self, data, assume_copies: bool = False) -> SimpleNamespace:
        """Upload the step-invariant part of a PyG-style batch (utils/sampling.py:56-67) and run the per-batch
        setup kernels.  Returns sizes / index info of the batch."""
        lig, rec = data['ligand'], data['receptor']
        B = int(data.num_graphs)
        host, mr_flat, RB = batch_index_arrays(data, self.hyper.no_torsion)
        rec_ptr = host.rec_ptr
        L = self.hyper.latent_dim
        unc = self.hyper.latent_droprate > 0
        nr0 = int(rec_ptr[1])
        share_rec = assume_copies and B > 1 and rec.x.shape[0] == B * nr0
        return self._upload(B, RB, host, mr_flat, lig_x=lig.x, bond_attr=data['ligand', 'ligand'].edge_attr,
                            rec_x=rec.x[:nr0] if share_rec else rec.x, rec_pos=rec.pos[:nr0] if share_rec else rec.pos,
                            rec_repeat=B if share_rec else 1, lig_repeat=1,
                            lig_latent=lig.latent_h if L > 0 else None, rec_latent=rec.latent_h if L > 0 else None,
                            lig_uncond=lig.unconditional if unc and 'unconditional' in lig else None,
                            rec_uncond=rec.unconditional if unc and 'unconditional' in rec else None)

    def set_batch_copies(self, proto, B: int) -> SimpleNamespace:
        """``B`` copies of one complex (what ``sampling()`` batches are, utils/sampling.py:57) without materialising the
        PyG batch on the host: the complex is shipped once and replicated on the device."""
        return self.set_batch_groups([(proto, B)])

    def set_batch_groups(self, groups) -> SimpleNamespace:
        """A batch made of runs of copies: ``groups = [(complex, n_copies), ...]`` in batch order (several complexes of an
        evaluate.py loop sampled together).  Every complex is shipped once -- its 1.5 MB receptor embedding included -- and
        replicated on the device; the host only builds the index arrays."""
        if self.hyper.latent_dim > 0:
            raise RuntimeError('set_batch_groups does not carry per-copy latents; use set_batch')
        dev = self.device
        host, mr_flat, RB = group_index_arrays(groups, self.hyper.no_torsion)
        h2d = 0

        def up(x, dtype, n):
            nonlocal h2d
            if not x.is_cuda:
                h2d += x.numel() * torch.empty(0, dtype=dtype).element_size()
            y = x.to(dev, dtype, non_blocking=True)
            return y.repeat(n, 1) if n > 1 else y

        cat = lambda parts: parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)
        lig_x = cat([up(g['ligand'].x, torch.int32, n) for g, n in groups])
        bond_attr = cat([up(g['ligand', 'ligand'].edge_attr, torch.float32, n) for g, n in groups])
        rec_x = cat([up(g['receptor'].x, torch.float32, n) for g, n in groups])
        rec_pos = cat([up(g['receptor'].pos, torch.float32, n) for g, n in groups])
        B = int(sum(n for _, n in groups))
        return self._upload(B, RB, host, mr_flat, lig_x=lig_x, bond_attr=bond_attr, rec_x=rec_x, rec_pos=rec_pos,
                            rec_repeat=1, lig_repeat=1, lig_latent=None, rec_latent=None, lig_uncond=None, rec_uncond=None,
                            h2d0=h2d)

    def _upload(self, B, RB, host, mr_flat, lig_x, bond_attr, rec_x, rec_pos, rec_repeat, lig_repeat, lig_latent, rec_latent,
                lig_uncond, rec_uncond, h2d0=0) -> SimpleNamespace:
        dev = self.device
        h2d = h2d0 + sum(getattr(host, k).nbytes for k in ('lig_ptr', 'rec_ptr', 'bond_index', 'bond_ptr', 'edge_mask',
                                                           'rec_index', 'rec_eptr', 'mr_off'))

        def up(x, dtype, repeat=1, flat=False):
            nonlocal h2d
            if x is None:
                return None
            if not x.is_cuda:
                h2d += x.numel() * torch.empty(0, dtype=dtype).element_size()
            y = x.to(dev, dtype, non_blocking=True)
            if flat:
                y = y.reshape(-1)
            if repeat > 1:
                y = y.repeat(repeat, *([1] * (y.dim() - 1)))
            return y.contiguous()

        t = SimpleNamespace()
        t.lig_x = up(lig_x, torch.int32, lig_repeat)
        t.bond_attr = up(bond_attr, torch.float32, lig_repeat)
        t.rec_x = up(rec_x, torch.float32, rec_repeat)
        t.rec_pos = up(rec_pos, torch.float32, rec_repeat)
        t.mask_rotate = up(torch.from_numpy(mr_flat), torch.uint8) if mr_flat is not None else None
        t.lig_latent = up(lig_latent, torch.float32)
        t.rec_latent = up(rec_latent, torch.float32)
        t.lig_uncond = up(lig_uncond, torch.float32, flat=True)
        t.rec_uncond = up(rec_uncond, torch.float32, flat=True)
        assert t.rec_x.shape[1] == 1281 and t.lig_x.shape[1] == 16
        b = DdkBatch(B=B, NL=int(host.lig_ptr[-1]), NR=int(host.rec_ptr[-1]), EB=host.bond_index.shape[1],
                     ER=host.rec_index.shape[1], RB=RB,
                     lig_ptr_h=_np_ptr(host.lig_ptr), rec_ptr_h=_np_ptr(host.rec_ptr), bond_index_h=_np_ptr(host.bond_index),
                     bond_ptr_h=_np_ptr(host.bond_ptr), edge_mask_h=_np_ptr(host.edge_mask), rec_index_h=_np_ptr(host.rec_index),
                     rec_edge_ptr_h=_np_ptr(host.rec_eptr), mask_rotate_off_h=_np_ptr(host.mr_off),
                     lig_x=_ptr(t.lig_x), bond_attr=_ptr(t.bond_attr), rec_x=_ptr(t.rec_x), rec_pos=_ptr(t.rec_pos),
                     mask_rotate=_ptr(t.mask_rotate), lig_latent=_ptr(t.lig_latent), rec_latent=_ptr(t.rec_latent),
                     lig_uncond=_ptr(t.lig_uncond), rec_uncond=_ptr(t.rec_uncond))
        assert t.lig_x.shape[0] == b.NL and t.rec_x.shape[0] == b.NR and t.bond_attr.shape[0] == b.EB
        with torch.cuda.device(dev):
            self._check(self.lib.ddk_set_batch(self.ctx, C.byref(b), self.stream()), 'ddk_set_batch')
        self._keep = [t, host]     # the context references rec_pos / bond_attr / masks / latents of the caller
        self.batch_info = SimpleNamespace(B=B, NL=b.NL, NR=b.NR, EB=b.EB, ER=b.ER, RB=RB, lig_ptr=host.lig_ptr,
                                          rec_ptr=host.rec_ptr, h2d_bytes=h2d)
        return self.batch_info

    # ------------------------------------------------------------------------------------------ steps
    def _step_inputs(self, semb, cutoff, tr_sigma, rot_scale, tor_scale):
        keep = [x.to(self.device, torch.float32).contiguous() if x is not None else None
                for x in (semb, cutoff, tr_sigma, rot_scale, tor_scale)]
        si = DdkStepInputs(sigma_emb=_ptr(keep[0]), cross_cutoff=_ptr(keep[1]), tr_sigma=_ptr(keep[2]),
                           rot_scale=_ptr(keep[3]), tor_scale=_ptr(keep[4]))
        return si, keep

    def score(self, pos, semb, cutoff, tr_sigma, rot_scale, tor_scale):
        """ddk_score: pos [NL,3] (device) -> tr [B,3], rot [B,3], tor [RB]."""
        bi = self.batch_info
        pos = pos.to(self.device, torch.float32).contiguous()
        si, keep = self._step_inputs(semb, cutoff, tr_sigma, rot_scale, tor_scale)
        tr = torch.empty(bi.B, 3, device=self.device)
        rot = torch.empty(bi.B, 3, device=self.device)
        use_tor = bi.RB > 0 and not self.hyper.no_torsion
        tor = torch.empty(bi.RB, device=self.device) if use_tor else None
        with torch.cuda.device(self.device):
            self._check(self.lib.ddk_score(self.ctx, _ptr(pos), C.byref(si), _ptr(tr), _ptr(rot), _ptr(tor), self.stream()),
                        'ddk_score')
        self._keep_step = (pos, keep)
        return tr, rot, (tor if use_tor else torch.empty(0, device=self.device))

    def embed(self, pos, semb, cutoff, tr_sigma, rot_scale, tor_scale):
        bi = self.batch_info
        pos = pos.to(self.device, torch.float32).contiguous()
        si, keep = self._step_inputs(semb, cutoff, tr_sigma, rot_scale, tor_scale)
        lig = torch.empty(bi.NL, 84, device=self.device)
        rec = torch.empty(bi.NR, 84, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.ddk_embed(self.ctx, _ptr(pos), C.byref(si), self.stream()), 'ddk_embed')
            self._check(self.lib.ddk_get_node_features(self.ctx, _ptr(lig), _ptr(rec), self.stream()), 'ddk_get_node_features')
        self._keep_step = (pos, keep)
        return lig, rec

    def update(self, pos, tr, rot, tor, z_tr, z_rot, z_tor, coef):
        cf = DdkStepCoef(*[float(x) for x in coef])
        with torch.cuda.device(self.device):
            self._check(self.lib.ddk_update(self.ctx, _ptr(pos), _ptr(tr), _ptr(rot), _ptr(tor if tor is not None and tor.numel() else None),
                                            _ptr(z_tr), _ptr(z_rot), _ptr(z_tor if z_tor is not None and z_tor.numel() else None),
                                            C.byref(cf), self.stream()), 'ddk_update')
        return pos

    def sample(self, pos, steps: 'StepTables', noise: Optional[Dict[str, torch.Tensor]]):
        """ddk_sample on device tensors: pos [NL,3] is updated in place and returned."""
        n = steps.n_steps
        si, keep = self._step_inputs(steps.semb, steps.cutoff, steps.tr_sigma, steps.rot_scale, steps.tor_scale)
        z = {k: (noise[k].to(self.device, torch.float32).contiguous() if noise is not None and noise.get(k) is not None else None)
             for k in ('tr', 'rot', 'tor')}
        coef = (DdkStepCoef * n)(*[DdkStepCoef(*[float(x) for x in row]) for row in steps.coef])
        with torch.cuda.device(self.device):
            self._check(self.lib.ddk_sample(self.ctx, _ptr(pos), n, C.byref(si), _ptr(z['tr']), _ptr(z['rot']),
                                            _ptr(z['tor'] if z['tor'] is not None and z['tor'].numel() else None), coef,
                                            self.stream()), 'ddk_sample')
        self._keep_step = (pos, keep, z)
        return pos

    def sample_host(self, pos_h: torch.Tensor, steps: 'StepTables', noise: Optional[Dict[str, torch.Tensor]]):
        """ddk_sample_host: every buffer lives in (pinned) host memory; returns the final poses in ``pos_h``."""
        n = steps.n_steps
        f32 = lambda x: None if x is None else x.detach().to('cpu', torch.float32).contiguous()
        arrs = [f32(x) for x in (steps.semb, steps.cutoff, steps.tr_sigma, steps.rot_scale, steps.tor_scale)]
        si = DdkStepInputs(*[_ptr(a) for a in arrs])
        z = {k: (f32(noise[k]) if noise is not None and noise.get(k) is not None else None) for k in ('tr', 'rot', 'tor')}
        coef = (DdkStepCoef * n)(*[DdkStepCoef(*[float(x) for x in row]) for row in steps.coef])
        assert pos_h.device.type == 'cpu' and pos_h.dtype == torch.float32 and pos_h.is_contiguous()
        with torch.cuda.device(self.device):
            self._check(self.lib.ddk_sample_host(self.ctx, _ptr(pos_h), n, C.byref(si), _ptr(z['tr']), _ptr(z['rot']),
                                                 _ptr(z['tor'] if z['tor'] is not None and z['tor'].numel() else None), coef),
                        'ddk_sample_host')
        return pos_h

    # ------------------------------------------------------------------------------------------ introspection
    def kernel_launches(self) -> int:
        return int(self.lib.ddk_kernel_launches(self.ctx))

    def edge_total(self) -> int:
        return int(self.lib.ddk_edge_total(self.ctx))

    def segment_total(self) -> int:
        return int(self.lib.ddk_segment_total(self.ctx))

    def group_totals(self):
        """Cumulative (edges[11], segments[11]) per work list: groups 0 lig-lig, 1 lig<-rec, 2 rec-rec, 3 rec<-lig, and 4 + h =
        group 2 restricted to residues within h receptor-contact hops of a residue with a cross edge."""
        e, s = np.zeros(11, np.int64), np.zeros(11, np.int64)
        self._check(self.lib.ddk_group_totals(self.ctx, _np_ptr(e), _np_ptr(s)), 'ddk_group_totals')
        return e, s

    def last_edge_count(self) -> int:
        return int(self.lib.ddk_last_edge_count(self.ctx))

    PROFILE_CLASSES = ['setup', 'graph', 'node_proj', 'conv_accum_lv0', 'conv_accum_lv1', 'conv_accum_lv2', 'conv_accum_lv3',
                       'conv_contract', 'heads', 'update', 'edge_hidden', 'conv_tc_lv0', 'conv_tc_lv1', 'conv_tc_lv2', 'conv_tc_lv3']

    def profile_enable(self, on=True):
        self._check(self.lib.ddk_profile_enable(self.ctx, int(on)), 'ddk_profile_enable')

    def profile_read(self):
        """{class: (milliseconds, launches)} since the last read (synchronises)."""
        ms = np.zeros(len(self.PROFILE_CLASSES), np.float64)
        n = np.zeros(len(self.PROFILE_CLASSES), np.int64)
        self._check(self.lib.ddk_profile_read(self.ctx, _np_ptr(ms), _np_ptr(n)), 'ddk_profile_read')
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.PROFILE_CLASSES)}

    def debug_read(self, name: str, dtype=np.float32) -> np.ndarray:
        n = C.c_size_t()
        self._check(self.lib.ddk_debug_read(self.ctx, name.encode(), None, 0, C.byref(n)), 'ddk_debug_read')
        out = np.empty(n.value // np.dtype(dtype).itemsize, dtype=dtype)
        self._check(self.lib.ddk_debug_read(self.ctx, name.encode(), _np_ptr(out), out.nbytes, C.byref(n)), 'ddk_debug_read')
        return out


# ---------------------------------------------------------------------------------------------- model-level glue
_SO3 = None
_TORUS = None


def score_norm_tables():
    """The reference's precomputed LUTs (utils/so3.py:51-66 ``_exp_score_norms``; utils/torus.py:72-76
    ``score_norm_``), exported once by oracle/make_tables.py from those very files."""
    global _SO3, _TORUS
    if _SO3 is None:
        _SO3 = np.load(os.path.join(_HERE, 'tables', 'so3_exp_score_norms.npy'))
        _TORUS = np.load(os.path.join(_HERE, 'tables', 'torus_score_norm.npy'))
    return _SO3, _TORUS


def so3_score_norm(eps: torch.Tensor) -> torch.Tensor:
    """utils/so3.py:91-95 for a float32 tensor on any device (no host synchronisation)."""
    so3, _ = score_norm_tables()
    table = torch.from_numpy(so3).to(eps.device)
    idx = (torch.log10(eps.float()).double() - np.log10(0.01)) / (np.log10(2) - np.log10(0.01)) * 1000
    idx = torch.clamp(torch.round(idx), 0, 999).long()
    return table[idx].float()


def torus_score_norm(sigma: torch.Tensor) -> torch.Tensor:
    """utils/torus.py:79-83."""
    _, torus = score_norm_tables()
    table = torch.from_numpy(torus).to(sigma.device)
    s = torch.log(sigma.float() / float(np.float32(np.pi))).double()
    s = (s - np.log(3e-3)) / (np.log(2) - np.log(3e-3)) * 5000
    idx = torch.round(torch.clamp(s, 0, 5000)).long()
    return table[idx].float()


def _graph_inputs(model, data):
    """Per-graph step inputs of one forward call, derived exactly like score_model.py:187, 203, 276, 286, 303-306."""
    ct = data.complex_t
    tr_sigma, rot_sigma, tor_sigma = model.t_to_sigma(*[ct[k] for k in ('tr', 'rot', 'tor')])
    tr_sigma, rot_sigma, tor_sigma = [torch.as_tensor(x).float() for x in (tr_sigma, rot_sigma, tor_sigma)]
    semb = model.timestep_emb_func(ct['tr'])
    if model.dynamic_max_cross:
        cutoff = tr_sigma * 3 + 20
    else:
        cutoff = torch.full_like(tr_sigma, model.cross_max_distance)
    if model.scale_by_sigma:
        rot_scale = so3_score_norm(rot_sigma)
        tor_scale = torch.sqrt(torus_score_norm(tor_sigma))
        trs = tr_sigma
    else:
        rot_scale, tor_scale, trs = torch.ones_like(rot_sigma), torch.ones_like(tor_sigma), torch.ones_like(tr_sigma)
    return semb, cutoff, trs, rot_scale, tor_scale, (tr_sigma, rot_sigma, tor_sigma)


def _batch_key(data):
    """Identity of every step-invariant input of a forward call, plus the tensors themselves.

    The key is (data_ptr, shape, _version) of each tensor.  That identifies CONTENT only while the tensor is alive -- the
    caching allocator hands a freed block to the next batch of the same shapes -- so the engine keeps the returned tensors
    referenced for as long as it keeps the key: a live tensor with the same address, shape and version counter is the same
    data (in-place writes bump ``_version``)."""
    lig, rec = data['ligand'], data['receptor']
    ll, rr = data['ligand', 'ligand'], data['receptor', 'receptor']
    parts = [lig.x, rec.x, rec.pos, ll.edge_index, ll.edge_attr, rr.edge_index, lig.edge_mask]
    if 'batch' in lig:
        parts += [lig.batch, rec.batch]
    if 'latent_h' in lig:
        parts += [lig.latent_h, rec.latent_h]
    if 'unconditional' in lig:
        parts += [lig.unconditional, rec.unconditional]
    mr = lig.mask_rotate if 'mask_rotate' in lig else None
    key = tuple((p.data_ptr(), tuple(p.shape), p._version, p.dtype) for p in parts) + (id(mr), int(data.num_graphs))
    return key, (parts, mr)


def _prepare(model, data):
    eng = model.engine(data['ligand'].pos.device if data['ligand'].pos.is_cuda else model.device)
    key, alive = _batch_key(data)
    if getattr(eng, '_batch_key', None) != key:
        eng.set_batch(data)
        eng._batch_key, eng._batch_alive = key, alive
    return eng


def forward_batch(model, data):
    eng = _prepare(model, data)
    semb, cutoff, trs, rot_scale, tor_scale, _ = _graph_inputs(model, data)
    lig = data['ligand']
    # side effects of the reference forward (score_model.py:276, 312, 348)
    data.graph_sigma_emb = semb
    lig.node_sigma_emb = semb[lig.batch.to(semb.device)]
    data['receptor'].node_sigma_emb = semb[data['receptor'].batch.to(semb.device)]
    tr, rot, tor = eng.score(lig.pos, semb, cutoff, trs, rot_scale, tor_scale)
    return tr, rot, tor


def embed_batch(model, data):
    eng = _prepare(model, data)
    semb, cutoff, trs, rot_scale, tor_scale, sig = _graph_inputs(model, data)
    lig_h, rec_h = eng.embed(data['ligand'].pos, semb, cutoff, trs, rot_scale, tor_scale)
    return lig_h, rec_h, sig[0], sig[1], sig[2]


class StepTables:
    """Per-step, per-graph scalars of a whole reverse-diffusion run, computed on the host from the schedule
    (t is uniform over the batch inside sampling(), utils/sampling.py:106-113)."""

    def __init__(self, n_steps, semb, cutoff, tr_sigma, rot_scale, tor_scale, coef):
        self.n_steps, self.semb, self.cutoff, self.tr_sigma = n_steps, semb, cutoff, tr_sigma
        self.rot_scale, self.tor_scale, self.coef = rot_scale, tor_scale, coef
