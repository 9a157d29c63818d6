"""Seeded synthetic protein-ligand complexes with the tensor schema of the reference's pre-processing.

There is no PDBBind, rdkit or ESM on the benchmark box, so workloads of the named shapes
(BASELINE.json ``configs``) are generated here.  The schema follows what
``/root/reference/datasets_utils/process_mols.py`` writes:

  * ``data['ligand'].x`` int64 [N_l,16] categorical atom features (:62-79), ``pos`` f32 [N_l,3];
    ``('ligand','lig_bond','ligand')`` bonds directed both ways, interleaved (u,v),(v,u) (:253-257) with a
    one-hot bond type [E_b,4] (:261); ``edge_mask`` bool [E_b] -- one direction per rotatable bond, pointing
    from the fixed side to the rotating side -- and ``mask_rotate`` bool [R,N_l]
    (``/root/reference/utils/torsion.py:15-45``).
  * ``data['receptor'].x`` f32 [N_r,1+1280] (amino-acid index as float, then the language-model embedding,
    :119-123, :371), ``pos`` = C-alpha minus protein centroid; ``('receptor','rec_contact','receptor')`` edges
    src-major, up to ``c_alpha_max_neighbors`` nearest within ``receptor_radius`` (:341-353).

Geometry: the receptor is a 3.8 A self-avoiding walk confined to a sphere of protein density, the ligand
a random tree with 1.5 A bonds.
"""
from __future__ import annotations

import numpy as np
import torch

from .data import Batch, HeteroData


def _unit(rng, n=1):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def make_receptor(rng, n_res, max_neighbors=24, radius=15.0, esm_dim=1280, esm_scale=0.08):
    R = (n_res * 135.0 * 3 / (4 * np.pi)) ** (1 / 3)
    pos = np.zeros((n_res, 3))
    pos[0] = _unit(rng)[0] * rng.uniform(0, 0.3 * R)
    i = 1
    while i < n_res:
        placed = False
        anchor = pos[i - 1]
        for attempt in range(400):
            if attempt == 200:      # stuck: restart the chain from a random earlier residue (a chain break)
                anchor = pos[rng.integers(0, i)]
            cand = anchor + 3.8 * _unit(rng)[0]
            if np.linalg.norm(cand) > R:
                continue
            if np.min(np.linalg.norm(pos[:i] - cand, axis=1)) < 3.0:
                continue
            pos[i] = cand
            placed = True
            break
        if not placed:              # give up on the sphere constraint for this residue
            pos[i] = anchor + 3.8 * _unit(rng)[0]
        i += 1
    pos -= pos.mean(0, keepdims=True)
    d = np.linalg.norm(pos[:, None] - pos[None], axis=-1)
    src, dst = [], []
    for a in range(n_res):
        nb = np.where(d[a] < radius)[0]
        nb = nb[nb != a]
        if len(nb) > max_neighbors:
            nb = np.argsort(d[a])[1:max_neighbors + 1]
        if len(nb) == 0:
            nb = np.argsort(d[a])[1:2]
        src += [a] * len(nb)
        dst += list(nb)
    aa = rng.integers(0, 20, size=(n_res, 1)).astype(np.float32)
    esm = rng.normal(scale=esm_scale, size=(n_res, esm_dim)).astype(np.float32)
    return (pos.astype(np.float32), np.concatenate([aa, esm], 1),
            np.asarray([src, dst], dtype=np.int64))


def make_ligand(rng, n_atoms, n_rot_target=(6, 10)):
    pos = np.zeros((n_atoms, 3))
    parent = -np.ones(n_atoms, dtype=int)
    deg = np.zeros(n_atoms, dtype=int)
    for i in range(1, n_atoms):
        for attempt in range(2000):
            if rng.random() < 0.8 and deg[i - 1] < 3:
                par = i - 1
            else:
                par = int(rng.integers(0, i))
                if deg[par] >= 4:
                    continue
            cand = pos[par] + 1.5 * _unit(rng)[0]
            dd = np.linalg.norm(pos[:i] - cand, axis=1)
            dd[par] = 10.0
            if dd.min() < 1.9:
                continue
            pos[i], parent[i] = cand, par
            deg[i] += 1
            deg[par] += 1
            break
        else:
            raise RuntimeError('could not grow the ligand tree')
    bonds = [(int(parent[i]), i) for i in range(1, n_atoms)]
    # a few ring closures between spatially close, non-bonded atoms
    bonded = set(bonds) | {(b, a) for a, b in bonds}
    d = np.linalg.norm(pos[:, None] - pos[None], axis=-1)
    cand = [(a, b) for a in range(n_atoms) for b in range(a + 1, n_atoms)
            if (a, b) not in bonded and d[a, b] < 2.6 and deg[a] < 4 and deg[b] < 4]
    rng.shuffle(cand)
    ring_bonds = []
    for a, b in cand[:3]:
        ring_bonds.append((a, b))
        deg[a] += 1
        deg[b] += 1
    all_bonds = bonds + ring_bonds
    btype = [0] * len(bonds) + [3] * len(ring_bonds)
    for k in range(len(bonds)):
        if rng.random() < 0.15:
            btype[k] = 1
    # bridges with both sides >= 2 atoms are rotatable candidates
    adj = [[] for _ in range(n_atoms)]
    for a, b in all_bonds:
        adj[a].append(b)
        adj[b].append(a)

    def component(start, cut):
        seen, stack = {start}, [start]
        while stack:
            u = stack.pop()
            for w in adj[u]:
                if (u, w) == cut or (w, u) == cut or w in seen:
                    continue
                seen.add(w)
                stack.append(w)
        return seen

    rot = {}
    for k, (a, b) in enumerate(all_bonds):
        ca = component(a, (a, b))
        if b in ca:
            continue
        cb = set(range(n_atoms)) - ca
        small = ca if len(ca) <= len(cb) else cb
        if len(small) > 1:
            rot[k] = small
    n_rot = int(rng.integers(n_rot_target[0], n_rot_target[1] + 1))
    keys = sorted(rng.permutation(sorted(rot.keys()))[:n_rot].tolist())
    row, col, et, emask, mrot = [], [], [], [], []
    for k, (a, b) in enumerate(all_bonds):
        row += [a, b]
        col += [b, a]
        et += [btype[k], btype[k]]
        if k in keys:
            small = rot[k]
            # the True direction points from the fixed side to the rotating (smaller) side (torsion.py:24-33)
            first = b in small
            emask += [first, not first]
            m = np.zeros(n_atoms, dtype=bool)
            m[sorted(small)] = True
            mrot.append(m)
        else:
            emask += [False, False]
    x = np.zeros((n_atoms, 16), dtype=np.int64)
    elem = rng.choice([5, 6, 7], size=n_atoms, p=[0.7, 0.15, 0.15])
    arom = np.zeros(n_atoms, dtype=int)
    for a, b in ring_bonds:
        arom[a] = arom[b] = 1
    x[:, 0] = elem
    x[:, 2] = np.clip(deg, 0, 10)
    x[:, 3] = 5
    x[:, 4] = np.clip(4 - deg, 0, 6) * (elem == 5)
    x[:, 5] = np.clip(4 - deg, 0, 8) * (elem == 5)
    x[:, 7] = np.where(arom == 1, 1, 2)
    x[:, 8] = arom
    x[:, 9] = arom
    x[:, 13] = arom
    eattr = np.zeros((len(row), 4), dtype=np.float32)
    eattr[np.arange(len(row)), et] = 1.0
    mask_rotate = np.stack(mrot) if mrot else np.zeros((0, n_atoms), dtype=bool)
    return (pos.astype(np.float32), x, np.asarray([row, col], dtype=np.int64), eattr,
            np.asarray(emask, dtype=bool), mask_rotate)


def make_complex(seed, n_lig=60, n_rec=300, max_neighbors=24, receptor_radius=15.0, esm_scale=0.08,
                 latent_dim=0, name=None) -> HeteroData:
    """One synthetic complex as the per-complex graph the reference's dataset would yield."""
    rng = np.random.default_rng(seed)
    rpos, rx, redges = make_receptor(rng, n_rec, max_neighbors, receptor_radius, esm_scale=esm_scale)
    lpos, lx, bonds, battr, emask, mrot = make_ligand(rng, n_lig)
    Rg = (n_rec * 135.0 * 3 / (4 * np.pi)) ** (1 / 3)
    pocket = _unit(rng)[0] * 0.5 * Rg
    lpos = lpos - lpos.mean(0, keepdims=True) + pocket
    g = HeteroData()
    g['ligand'].x = torch.from_numpy(lx)
    g['ligand'].pos = torch.from_numpy(lpos.astype(np.float32))
    g['ligand'].edge_mask = torch.from_numpy(emask)
    g['ligand'].mask_rotate = mrot
    g['ligand', 'lig_bond', 'ligand'].edge_index = torch.from_numpy(bonds)
    g['ligand', 'lig_bond', 'ligand'].edge_attr = torch.from_numpy(battr)
    g['receptor'].x = torch.from_numpy(rx)
    g['receptor'].pos = torch.from_numpy(rpos)
    g['receptor', 'rec_contact', 'receptor'].edge_index = torch.from_numpy(redges)
    g.name = name or f'synth_{seed}'
    g.original_center = torch.zeros(1, 3)
    return g


def as_loader_item(g: HeteroData) -> Batch:
    """What ``DataLoader(dataset, batch_size=1)`` hands to evaluate.py:221 (a Batch of one graph, so that
    ``mask_rotate`` is a list and ``name`` a list, sampling.py:57, evaluate.py:249)."""
    return Batch.from_data_list([g])


def max_radius_degree(pos: torch.Tensor, r=5.0) -> int:
    d = torch.cdist(pos, pos)
    return int((d < r).sum(1).max()) - 1
