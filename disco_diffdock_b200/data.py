"""Duck-typed stand-ins for the torch_geometric containers the reference's hot path touches.

PyG is not installed in this image, and the sampler only needs a very small part of it:
``HeteroData`` (attribute stores keyed by node type / edge type), ``Batch.from_data_list`` and a
``DataLoader`` that yields batches.  The containers below reproduce exactly the access patterns used
by the reference (``/root/reference/utils/sampling.py:55-67, 200-221``,
``/root/reference/models/score_model.py:310-438``, ``/root/reference/utils/diffusion_utils.py:37-41,
101-117``); real PyG objects expose the same attributes, so the sampler accepts either.

Edge stores can be addressed with 2-tuples (``data['ligand', 'ligand']``) or the 3-tuples the
reference's pre-processing uses (``('ligand', 'lig_bond', 'ligand')``,
``/root/reference/datasets_utils/process_mols.py:265-266, 375``); both resolve to the same store.
"""
from __future__ import annotations

import copy
from typing import Any, Dict, Iterable, List, Sequence

import numpy as np
import torch


def _move(v, device):
    if torch.is_tensor(v):
        return v.to(device)
    if isinstance(v, dict):
        return {k: _move(x, device) for k, x in v.items()}
    return v


class Store:
    """Attribute bag for one node type or one edge type."""

    def __init__(self, **kw):
        object.__setattr__(self, '_d', {})
        for k, v in kw.items():
            self._d[k] = v

    def __getattr__(self, k):
        if k.startswith('__') and k.endswith('__'):
            raise AttributeError(k)
        d = object.__getattribute__(self, '_d')
        if k in d:
            return d[k]
        if k == 'num_nodes':
            for cand in ('x', 'pos', 'batch'):
                if cand in d and d[cand] is not None:
                    return int(d[cand].shape[0])
            raise AttributeError(k)
        if k == 'num_edges':
            if 'edge_index' in d:
                return int(d['edge_index'].shape[1])
            raise AttributeError(k)
        raise AttributeError(k)

    def __setattr__(self, k, v):
        self._d[k] = v

    def __delattr__(self, k):
        del self._d[k]

    def __contains__(self, k):
        return k in self._d

    def __getitem__(self, k):
        return self._d[k]

    def __setitem__(self, k, v):
        self._d[k] = v

    def keys(self):
        return self._d.keys()

    def items(self):
        return self._d.items()

    def to(self, device):
        for k in list(self._d.keys()):
            self._d[k] = _move(self._d[k], device)
        return self

    def __deepcopy__(self, memo):
        s = Store()
        for k, v in self._d.items():
            s._d[k] = copy.deepcopy(v, memo)
        return s

    def __repr__(self):
        def sh(v):
            if torch.is_tensor(v) or isinstance(v, np.ndarray):
                return list(v.shape)
            return type(v).__name__
        return 'Store(' + ', '.join(f'{k}={sh(v)}' for k, v in self._d.items()) + ')'


_EDGE_ALIASES = {
    ('ligand', 'ligand'): ('ligand', 'lig_bond', 'ligand'),
    ('receptor', 'receptor'): ('receptor', 'rec_contact', 'receptor'),
}


class HeteroData:
    """Minimal heterogeneous graph: ``data['ligand']``, ``data['ligand','ligand']``, graph-level attrs."""

    def __init__(self):
        object.__setattr__(self, '_node_stores', {})
        object.__setattr__(self, '_edge_stores', {})
        object.__setattr__(self, '_attrs', {})

    # --- store access -------------------------------------------------------------------------
    def _edge_key(self, key):
        key = tuple(key)
        if len(key) == 3:
            return key
        if len(key) == 2:
            for k in self._edge_stores:
                if k[0] == key[0] and k[2] == key[1]:
                    return k
            return _EDGE_ALIASES.get(key, (key[0], 'to', key[1]))
        raise KeyError(key)

    def __getitem__(self, key):
        if isinstance(key, str):
            if key in self._node_stores:
                return self._node_stores[key]
            if key in self._attrs:
                return self._attrs[key]
            st = Store()
            self._node_stores[key] = st
            return st
        k = self._edge_key(key)
        if k not in self._edge_stores:
            self._edge_stores[k] = Store()
        return self._edge_stores[k]

    def __setitem__(self, key, value):
        if isinstance(key, str):
            if isinstance(value, Store):
                self._node_stores[key] = value
            else:
                self._attrs[key] = value
        else:
            self._edge_stores[self._edge_key(key)] = value

    def __contains__(self, key):
        if isinstance(key, str):
            return key in self._node_stores or key in self._attrs
        return self._edge_key(key) in self._edge_stores

    # --- graph level attributes ---------------------------------------------------------------
    def __getattr__(self, k):
        if k.startswith('__') and k.endswith('__'):
            raise AttributeError(k)
        attrs = object.__getattribute__(self, '_attrs')
        if k in attrs:
            return attrs[k]
        if k == 'num_graphs':
            return 1
        raise AttributeError(k)

    def __setattr__(self, k, v):
        self._attrs[k] = v

    @property
    def node_types(self):
        return list(self._node_stores.keys())

    @property
    def edge_types(self):
        return list(self._edge_stores.keys())

    def to(self, device):
        for s in self._node_stores.values():
            s.to(device)
        for s in self._edge_stores.values():
            s.to(device)
        for k in list(self._attrs.keys()):
            self._attrs[k] = _move(self._attrs[k], device)
        return self

    def shallow_copy(self):
        """New container and new stores that share the underlying tensors (cheap stand-in for the per-sample
        ``copy.deepcopy(orig_complex_graph)`` of evaluate.py:232 when only ``['ligand'].pos`` will be rebound)."""
        out = self.__class__()
        for k, s in self._node_stores.items():
            out._node_stores[k] = Store(**s._d)
        for k, s in self._edge_stores.items():
            out._edge_stores[k] = Store(**s._d)
        out._attrs.update(self._attrs)
        return out

    def __deepcopy__(self, memo):
        out = self.__class__()
        for k, s in self._node_stores.items():
            out._node_stores[k] = copy.deepcopy(s, memo)
        for k, s in self._edge_stores.items():
            out._edge_stores[k] = copy.deepcopy(s, memo)
        for k, v in self._attrs.items():
            out._attrs[k] = copy.deepcopy(v, memo)
        return out

    def __repr__(self):
        return (f'{self.__class__.__name__}(nodes={self._node_stores}, edges={self._edge_stores}, '
                f'attrs={list(self._attrs.keys())})')


def public_view(item):
    """(node stores, edge stores, graph attributes) of a graph as plain dicts, read through the PUBLIC container API only
    (``node_types``, ``edge_types``, ``item[key].items()``, ``to_dict()``), so that real ``torch_geometric`` ``HeteroData`` /
    ``HeteroDataBatch`` objects -- which have none of this module's private fields -- collate exactly like our own."""
    if isinstance(item, HeteroData):
        return ({k: s._d for k, s in item._node_stores.items()}, {k: s._d for k, s in item._edge_stores.items()}, item._attrs)
    nts = list(item.node_types)
    ets = [tuple(et) for et in item.edge_types]
    node = {nt: dict(item[nt].items()) for nt in nts}
    edge = {et: dict(item[et].items()) for et in ets}
    attrs: Dict[str, Any] = {}
    if hasattr(item, 'to_dict'):
        d = item.to_dict()
        if isinstance(d.get('_global_store'), dict):                 # PyG >= 2.0 keeps graph-level attributes there
            attrs.update(d['_global_store'])
        for k, v in d.items():
            if isinstance(k, str) and k != '_global_store' and k not in node:
                attrs[k] = v
    else:                                                            # last resort: the names the hot path reads
        for k in ('name', 'original_center', 'complex_t', 'rmsd_matching', 'success'):
            try:
                attrs[k] = getattr(item, k)
            except AttributeError:
                pass
    attrs.pop('num_graphs', None)
    return node, edge, attrs


def _store_num_nodes(d):
    for cand in ('x', 'pos', 'batch'):
        if cand in d and d[cand] is not None:
            return int(d[cand].shape[0])
    raise AttributeError('num_nodes')


class Batch(HeteroData):
    """Disjoint union of graphs, PyG ``Batch.from_data_list`` semantics for the attributes on the path.

    Tensors on node stores are concatenated along dim 0, ``edge_index`` is offset by the running node
    count of its source / destination node type, other tensors on edge stores are concatenated along
    dim 0, every node store gets ``batch`` (graph id per node) and ``ptr`` (CSR offsets).  Non-tensor
    attributes (``mask_rotate`` numpy arrays, ``name`` strings) are collected into Python lists, which is
    what ``sampling()`` relies on (``data_list[0]['ligand'].mask_rotate[0]``, sampling.py:57, applies to a
    per-complex graph that itself came out of a batch-size-1 loader, evaluate.py:138).
    """

    @classmethod
    def from_data_list(cls, data_list: Sequence[Any]) -> 'Batch':
        out = cls()
        n = len(data_list)
        views = [public_view(d) for d in data_list]
        nodes0, edges0, attrs0 = views[0]
        offsets: Dict[str, List[int]] = {}
        for nt in nodes0:
            stores = [v[0][nt] for v in views]
            st = Store()
            counts = [_store_num_nodes(s) for s in stores]
            off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
            offsets[nt] = off.tolist()
            for k in stores[0].keys():
                if k in ('batch', 'ptr'):                  # items that are batches themselves (loader items): regenerated below
                    continue
                vals = [s[k] for s in stores]
                if torch.is_tensor(vals[0]):
                    if vals[0].dim() == 0:
                        st._d[k] = torch.stack(vals)
                    else:
                        st._d[k] = torch.cat(vals, dim=0)
                elif isinstance(vals[0], dict) and all(torch.is_tensor(x) for x in vals[0].values()):
                    st._d[k] = {kk: torch.cat([v[kk] for v in vals], dim=0) for kk in vals[0]}
                else:
                    st._d[k] = list(vals)
            dev = st._d['pos'].device if 'pos' in st._d else None
            st._d['batch'] = torch.repeat_interleave(torch.arange(n), torch.tensor(counts)).to(dev)
            st._d['ptr'] = torch.from_numpy(off).to(dev)
            out._node_stores[nt] = st
        for et in edges0:
            stores = [v[1][et] for v in views]
            st = Store()
            for k in stores[0].keys():
                vals = [s[k] for s in stores]
                if k == 'edge_index':
                    so, do = offsets[et[0]], offsets[et[2]]
                    shifted = []
                    for i, v in enumerate(vals):
                        add = torch.tensor([[so[i]], [do[i]]], dtype=v.dtype, device=v.device)
                        shifted.append(v + add)
                    st._d[k] = torch.cat(shifted, dim=1)
                elif torch.is_tensor(vals[0]):
                    st._d[k] = torch.cat(vals, dim=0)
                else:
                    st._d[k] = list(vals)
            out._edge_stores[et] = st
        for k in attrs0:
            vals = [v[2][k] for v in views]
            if torch.is_tensor(vals[0]):
                out._attrs[k] = torch.cat([v.reshape(1, *v.shape[1:]) if v.dim() > 0 and v.shape[0] == 1 else v.unsqueeze(0)
                                           for v in vals], dim=0)
            elif isinstance(vals[0], dict) and all(torch.is_tensor(x) for x in vals[0].values()):
                out._attrs[k] = {kk: torch.cat([v[kk] for v in vals], dim=0) for kk in vals[0]}
            else:
                out._attrs[k] = list(vals)
        out._attrs['num_graphs'] = n
        return out


class DataLoader:
    """``torch_geometric.loader.DataLoader(data_list, batch_size)`` without shuffling (sampling.py:56)."""

    def __init__(self, dataset: Sequence[HeteroData], batch_size: int = 1, shuffle: bool = False, **_):
        assert not shuffle, 'the sampler never shuffles'
        self.dataset = dataset
        self.batch_size = int(batch_size)

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for i in range(0, len(self.dataset), self.batch_size):
            yield Batch.from_data_list(self.dataset[i:i + self.batch_size])
