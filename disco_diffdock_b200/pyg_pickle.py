"""Reader for the reference's pre-processed complex cache -- ``heterographs.pkl`` written by
``/root/reference/datasets_utils/pdbbind.py:177-189`` and read back at ``:112-117`` -- on a box WITHOUT torch_geometric.

The file is a plain ``pickle`` of a list of ``torch_geometric.data.HeteroData``.  Unpickling normally imports torch_geometric;
here every ``torch_geometric.*`` class is replaced by a state recorder, and the recorded state -- PyG's ``_global_store`` /
``_node_store_dict`` / ``_edge_store_dict``, each a storage whose attributes live in ``_mapping`` (torch_geometric/data/storage.py,
hetero_data.py; the layout of every 2.x release) -- is rebuilt as this package's ``HeteroData``.  Tensors and numpy arrays
unpickle through their own reducers.  ``rdkit_ligands.pkl`` holds rdkit molecules and needs rdkit; the sampler does not read it.

Only classes of torch_geometric, torch, numpy and the builtins are accepted: a pickle is code, so anything else raises.
"""
from __future__ import annotations

import io
import pickle
from typing import Any, List

from .data import HeteroData

_SAFE_ROOTS = ('torch', 'numpy', 'collections', 'builtins', 'copyreg', '_codecs')


class _Recorded:
    """Stands for any torch_geometric object: keeps whatever state the pickle hands it."""

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):      # (dict state, slots state)
            merged = dict(state[0] or {})
            merged.update(state[1])
            state = merged
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__['_state'] = state


class _Unpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str) -> Any:
        root = module.split('.')[0]
        if root == 'torch_geometric':
            return type(name, (_Recorded,), {'__module__': module})
        if root in _SAFE_ROOTS:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f'refusing to import {module}.{name} from a complex cache')


def _mapping(store) -> dict:
    if store is None:
        return {}
    d = store.__dict__ if isinstance(store, _Recorded) else dict(store)
    m = d.get('_mapping', d)
    return {k: v for k, v in m.items() if not (isinstance(k, str) and k.startswith('_'))}


def to_hetero(obj) -> HeteroData:
    """A recorded torch_geometric HeteroData / HeteroDataBatch -> this package's container."""
    if isinstance(obj, HeteroData):
        return obj
    d = obj.__dict__
    if '_node_store_dict' not in d:
        raise ValueError(f'not a torch_geometric HeteroData pickle (recorded fields: {sorted(d)[:8]})')
    g = HeteroData()
    for nt, store in d['_node_store_dict'].items():
        for k, v in _mapping(store).items():
            setattr(g[nt], k, v)
    for et, store in d['_edge_store_dict'].items():
        for k, v in _mapping(store).items():
            setattr(g[tuple(et)], k, v)
    for k, v in _mapping(d.get('_global_store')).items():
        setattr(g, k, v)
    return g


def load_heterographs(path_or_bytes) -> List[HeteroData]:
    """``pickle.load(open('<cache>/heterographs.pkl', 'rb'))`` of pdbbind.py:113-114 without torch_geometric."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        f = io.BytesIO(path_or_bytes)
        obj = _Unpickler(f).load()
    else:
        with open(path_or_bytes, 'rb') as f:
            obj = _Unpickler(f).load()
    if not isinstance(obj, (list, tuple)):
        obj = [obj]
    return [to_hetero(o) for o in obj]
