"""the reference's ``heterographs.pkl`` cache (datasets_utils/pdbbind.py:112-117, 177-189) read without
torch_geometric.  The pickle is produced here by stand-in classes that live under the real module paths
(``torch_geometric.data.hetero_data.HeteroData``, ``torch_geometric.data.storage.*``) and keep their state exactly where PyG 2.x
keeps it (``_global_store`` / ``_node_store_dict`` / ``_edge_store_dict``, attributes in ``_mapping``, ``_parent`` dropped by
``__getstate__``); the stand-in modules are removed again before the reader runs."""
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from disco_diffdock_b200 import data as ddata
from disco_diffdock_b200 import pyg_pickle, synthetic


def _fake_pyg_pickle(graphs):
    mods = {}
    for name in ('torch_geometric', 'torch_geometric.data', 'torch_geometric.data.storage', 'torch_geometric.data.hetero_data'):
        mods[name] = types.ModuleType(name)

    def storage_cls(name):
        def __init__(self, mapping, key=None):
            self._mapping, self._key, self._parent = dict(mapping), key, object()        # _parent: a weakref in PyG

        def __getstate__(self):
            out = self.__dict__.copy()
            out['_parent'] = None                                                        # storage.py drops the weakref
            return out
        return type(name, (), {'__init__': __init__, '__getstate__': __getstate__, '__module__': 'torch_geometric.data.storage'})
    Base, Node, Edge = storage_cls('BaseStorage'), storage_cls('NodeStorage'), storage_cls('EdgeStorage')
    Hetero = type('HeteroData', (), {'__module__': 'torch_geometric.data.hetero_data'})
    for c in (Base, Node, Edge):
        setattr(mods['torch_geometric.data.storage'], c.__name__, c)
    mods['torch_geometric.data.hetero_data'].HeteroData = Hetero
    sys.modules.update(mods)
    try:
        objs = []
        for g in graphs:
            node, edge, attrs = ddata.public_view(g)
            h = Hetero()
            h.__dict__['_global_store'] = Base(attrs)
            h.__dict__['_node_store_dict'] = {k: Node(v, k) for k, v in node.items()}
            h.__dict__['_edge_store_dict'] = {k: Edge(v, k) for k, v in edge.items()}
            objs.append(h)
        return pickle.dumps(objs)
    finally:
        for name in mods:
            sys.modules.pop(name, None)


def test_heterographs_pickle_reads_without_torch_geometric():
    graphs = [synthetic.make_complex(5, 9, 14), synthetic.make_complex(6, 13, 20)]
    for g in graphs:
        g['ligand'].orig_pos = g['ligand'].pos.numpy().copy()
        g.rmsd_matching = 0.25
    blob = _fake_pyg_pickle(graphs)
    assert 'torch_geometric' not in sys.modules
    out = pyg_pickle.load_heterographs(blob)
    assert len(out) == 2 and 'torch_geometric' not in sys.modules
    for a, b in zip(graphs, out):
        for nt in ('ligand', 'receptor'):
            assert torch.equal(a[nt].x, b[nt].x) and torch.equal(a[nt].pos, b[nt].pos)
        assert torch.equal(a['ligand', 'ligand'].edge_index, b['ligand', 'lig_bond', 'ligand'].edge_index)
        assert torch.equal(a['ligand', 'ligand'].edge_attr, b['ligand', 'ligand'].edge_attr)
        assert torch.equal(a['receptor', 'receptor'].edge_index, b['receptor', 'rec_contact', 'receptor'].edge_index)
        assert torch.equal(a['ligand'].edge_mask, b['ligand'].edge_mask)
        assert np.array_equal(np.asarray(a['ligand'].mask_rotate), np.asarray(b['ligand'].mask_rotate))
        assert np.array_equal(a['ligand'].orig_pos, b['ligand'].orig_pos)
        assert a.name == b.name and b.rmsd_matching == 0.25
    # what the sampler does with them: collate copies of a complex (utils/sampling.py:56) and group them
    from disco_diffdock_b200 import sampling as dsampling
    items = [synthetic.as_loader_item(out[0]) for _ in range(3)]
    batch = ddata.Batch.from_data_list(items)
    assert batch.num_graphs == 3 and batch['ligand'].pos.shape[0] == 27
    assert [n for _, n in dsampling.group_copies(items)] == [3]


def test_pickle_reader_refuses_foreign_classes():
    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ('true',))
    with pytest.raises(pickle.UnpicklingError):
        pyg_pickle.load_heterographs(pickle.dumps([Evil()]))
