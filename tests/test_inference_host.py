"""Host logic of the evaluate.py-style driver: runs of copies are recognised, the index arrays built
from one graph per run equal those of the collated PyG-style batch, the complex pack round-trips, the pose metrics follow
evaluate.py:301-337."""
import copy

import numpy as np
import torch

from disco_diffdock_b200 import data as ddata
from disco_diffdock_b200 import synthetic
from disco_diffdock_b200.engine import batch_index_arrays, group_index_arrays
from disco_diffdock_b200.inference import load_complex_pack, pose_metrics, save_complex_pack
from disco_diffdock_b200.sampling import group_copies


def _mixed_list():
    gs = [synthetic.make_complex(31, 14, 30), synthetic.make_complex(32, 22, 41), synthetic.make_complex(33, 9, 17)]
    gs[2]['ligand'].edge_mask = torch.zeros_like(gs[2]['ligand'].edge_mask)          # a ligand without rotatable bonds
    gs[2]['ligand'].mask_rotate = np.zeros((0, 9), dtype=bool)
    counts = [3, 1, 2]
    items = []
    for g, n in zip(gs, counts):
        for k in range(n):
            c = copy.deepcopy(g) if k % 2 == 0 else g.shallow_copy()                   # both kinds of copies
            c['ligand'].pos = c['ligand'].pos + float(k)                               # poses differ between copies
            items.append(c)
    return gs, counts, items


def test_group_copies_finds_runs():
    gs, counts, items = _mixed_list()
    groups = group_copies(items)
    assert [n for _, n in groups] == counts
    assert all(g is items[sum(counts[:i])] for i, (g, _) in enumerate(groups))
    # two different complexes of identical sizes are not merged
    a, b = synthetic.make_complex(41, 10, 20), synthetic.make_complex(42, 10, 20)
    assert [n for _, n in group_copies([a, copy.deepcopy(a), b])] == [2, 1]


def test_group_index_arrays_match_collated_batch():
    gs, counts, items = _mixed_list()
    want, want_mr, want_rb = batch_index_arrays(ddata.Batch.from_data_list(items), no_torsion=False)
    got, got_mr, got_rb = group_index_arrays(group_copies(items), no_torsion=False)
    assert got_rb == want_rb
    for k in ('lig_ptr', 'rec_ptr', 'bond_ptr', 'rec_eptr', 'bond_index', 'rec_index', 'edge_mask'):
        a, b = getattr(got, k), getattr(want, k)
        assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), k
    # mask_rotate blocks: offsets may differ, the block each graph points at may not
    for g in range(len(items)):
        nl = int(want.lig_ptr[g + 1] - want.lig_ptr[g])
        r = int(want.edge_mask[want.bond_ptr[g]:want.bond_ptr[g + 1]].sum())
        a = got_mr[got.mr_off[g]:got.mr_off[g] + r * nl]
        b = want_mr[want.mr_off[g]:want.mr_off[g] + r * nl]
        assert np.array_equal(a, b)
    # no_torsion: no mask at all
    assert group_index_arrays(group_copies(items), no_torsion=True)[1] is None


def test_complex_pack_round_trip(tmp_path):
    gs = [synthetic.make_complex(51, 12, 25), synthetic.as_loader_item(synthetic.make_complex(52, 8, 16))]
    gs[0]['ligand'].orig_pos = gs[0]['ligand'].pos.numpy() + 1.0
    p = str(tmp_path / 'pack.npz')
    save_complex_pack(p, gs)
    back = load_complex_pack(p)
    assert len(back) == 2 and back[0].name == 'synth_51' and back[1].name == 'synth_52'
    for a, b in zip(gs, back):
        for st in ('ligand', 'receptor'):
            for f in ('x', 'pos'):
                assert torch.equal(a[st][f], b[st][f]) and a[st][f].dtype == b[st][f].dtype
        assert torch.equal(a['ligand', 'ligand'].edge_index, b['ligand', 'ligand'].edge_index)
        assert torch.equal(a['ligand', 'ligand'].edge_attr, b['ligand', 'ligand'].edge_attr)
        assert torch.equal(a['receptor', 'receptor'].edge_index, b['receptor', 'receptor'].edge_index)
        assert torch.equal(a['ligand'].edge_mask, b['ligand'].edge_mask)
    assert np.array_equal(back[0]['ligand'].orig_pos, gs[0]['ligand'].orig_pos)
    # a loaded complex groups with its own copies and yields the same index arrays as the original
    w, _, _ = group_index_arrays([(gs[0], 2)], False)
    g, _, _ = group_index_arrays(group_copies([back[0], copy.deepcopy(back[0])]), False)
    assert np.array_equal(w.bond_index, g.bond_index) and np.array_equal(w.rec_index, g.rec_index)


def test_pose_metrics_follow_evaluate():
    g = synthetic.make_complex(61, 10, 20)
    g['ligand'].x[3, 0] = 0                                                           # one "hydrogen": filtered out
    g['ligand'].orig_pos = g['ligand'].pos.numpy().copy()
    dl = [copy.deepcopy(g) for _ in range(3)]
    dl[1]['ligand'].pos = dl[1]['ligand'].pos + torch.tensor([[3.0, 0.0, 4.0]])
    m = pose_metrics(g, dl)
    assert m['ligand_pos'].shape == (3, 9, 3)
    assert np.allclose(m['rmsd'], [0.0, 5.0, 0.0], atol=1e-5) and np.allclose(m['centroid_distance'], [0.0, 5.0, 0.0], atol=1e-5)
    heavy = g['ligand'].pos.numpy()[np.arange(10) != 3]
    d = np.linalg.norm(heavy[:, None] - heavy[None], axis=-1) + np.eye(9) * 1e9
    assert np.isclose(m['min_self_distance'][0], d.min(), atol=1e-5)
    c = np.linalg.norm(g['receptor'].pos.numpy()[:, None] - heavy[None], axis=-1)
    assert np.isclose(m['min_cross_distance'][0], c.min(), atol=1e-5)
