"""CPU tests of the host-side pieces of the drop-in against the reference's own functions (run behind the leaf-op shims where
/root/reference is mounted; committed fixtures replay the same checks anywhere):
  * randomize_position (utils/sampling.py:12-46) -- the start poses both paths share;
  * GenericEncoder.encode_ar with MULTINOMIAL decoding (models/model_classes.py:9-49, temperature < 100);
  * strict loading of the three shipped checkpoints into the drop-in modules (evaluate.py:160-181);
  * collation of graphs that expose only the PUBLIC torch_geometric API (no private fields of our own containers)."""
import copy
import os

import numpy as np
import pytest
import torch

from disco_diffdock_b200 import data as ddata
from disco_diffdock_b200 import latent as dlatent
from disco_diffdock_b200 import sampling as dsampling
from disco_diffdock_b200 import synthetic
from oracle import ref_loader
from tests import helpers
from tests.test_oracle_golden import GOLD

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason='reference tree not mounted')


# ---------------------------------------------------------------------------------------------- randomize_position
def _rp_inputs(n=3):
    g = synthetic.make_complex(41, 22, 30)
    return [synthetic.as_loader_item(copy.deepcopy(g)) for _ in range(n)]


def _seed(s=17):
    np.random.seed(s)
    torch.manual_seed(s)


@needs_ref
@pytest.mark.parametrize('no_random', [False, True])
def test_randomize_position_matches_reference(no_random):
    """Same RNG call order as the reference: with the same seeds the start poses agree to fp32 rounding."""
    mods = ref_loader.modules()
    a, b = _rp_inputs(), _rp_inputs()
    _seed(); mods.sampling.randomize_position(a, False, no_random, 19.0)
    _seed(); dsampling.randomize_position(b, False, no_random, 19.0)
    for x, y in zip(a, b):
        assert float((x['ligand'].pos - y['ligand'].pos).abs().max()) < 2e-5
    if no_random:
        assert float(b[0]['ligand'].pos.mean(0).abs().max()) < 1e-5      # centred on the protein, sampling.py:36-40


def test_randomize_position_matches_committed_fixture():
    z = np.load(os.path.join(GOLD, 'randomize_position.npz'))
    for key, no_random in (('random', False), ('no_random', True)):
        b = _rp_inputs()
        _seed(); dsampling.randomize_position(b, False, no_random, 19.0)
        got = torch.stack([x['ligand'].pos for x in b]).numpy()
        assert np.abs(got - z[key]).max() < 2e-5


# ---------------------------------------------------------------------------------------------- AR multinomial decoding
class _StubScore:
    """Stands for the pretrained score model inside PretrainedScoreEncoder: deterministic node features that depend on the
    latents decoded so far, so the second decoding step sees the first (same stub under the reference's encoder and ours)."""
    num_conv_layers = 5

    def embed(self, data):
        outs = []
        for nt, seed in (('ligand', 3), ('receptor', 4)):
            n = data[nt].pos.shape[0]
            gen = torch.Generator().manual_seed(seed)
            base = torch.randn(n, 84, generator=gen)
            lat = data[nt].latent_h.float()
            outs.append(base + 0.7 * lat.sum(1, keepdim=True) * torch.roll(base, 1, 0) + data[nt].pos.sum(1, keepdim=True) * 0.01)
        return outs[0], outs[1], None, None, None


def _ar_case():
    g = synthetic.make_complex(43, 12, 20)
    B = 4
    lst = [synthetic.as_loader_item(copy.deepcopy(g)) for _ in range(B)]
    heads = helpers.make_ar_heads(5, ns=16)
    return lst, heads, B


def _our_encode(temp, seed):
    lst, heads, B = _ar_case()
    ar = dlatent.PretrainedScoreEncoder(_StubScore(), 16, 1, 1, input_latent_dim=2)
    ar.load_state_dict(heads, strict=True)
    ar.eval()
    torch.manual_seed(seed)
    return ar.encode_ar(ddata.Batch.from_data_list(lst), temp)


@needs_ref
@pytest.mark.parametrize('temp', [1.0, 3.0, 100.0])
def test_encode_ar_multinomial_matches_reference(temp):
    """torch.multinomial on the same global-RNG state picks the same nodes in the reference's encode_ar and in ours."""
    import contextlib, io
    ref_loader.modules()
    with contextlib.redirect_stdout(io.StringIO()):
        from models.pretrained_score_encoder import PretrainedScoreEncoder as RefEnc
    lst, heads, B = _ar_case()
    ref = RefEnc(pretrained_score_model=_StubScore(), ns=16, latent_dim=1, latent_vocab=1, input_latent_dim=2)
    ref.load_state_dict(heads, strict=True)
    ref.eval()
    torch.manual_seed(123)
    with torch.no_grad():
        want = ref.encode_ar(ddata.Batch.from_data_list(lst), temp)
        got = _our_encode(temp, 123)
    assert torch.equal(want[0], got[0]) and torch.equal(want[1], got[1])
    assert float(got[0].sum() + got[1].sum()) == 2 * B            # one node per latent dimension and graph


def test_encode_ar_multinomial_matches_committed_fixture():
    z = np.load(os.path.join(GOLD, 'encode_ar_multinomial.npz'))
    with torch.no_grad():
        for temp in (1.0, 3.0):
            l, r = _our_encode(temp, 123)
            assert np.array_equal(l.numpy(), z[f'l_{temp}']) and np.array_equal(r.numpy(), z[f'r_{temp}'])
    # different draws at temperature 1 (the distribution is not degenerate)
    with torch.no_grad():
        a, b = _our_encode(1.0, 1), _our_encode(1.0, 2)
    assert not (torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]))


# ---------------------------------------------------------------------------------------------- shipped checkpoints
@pytest.mark.skipif(helpers.checkpoint_path('diffdockS') is None, reason='shipped checkpoints not on this box')
def test_shipped_checkpoints_load_strict():
    """evaluate.py:160-181: score model (DiffDock-S, DisCo) and the AR model (PretrainedScoreEncoder around the DisCo score
    model, ns = 16 from its model_parameters.yml) load with strict=True into the drop-in modules."""
    for name, nkeys in (('diffdockS', 182), ('disco', 187)):
        m, sd, cfg = helpers.make_checkpoint_model(name)
        assert len(sd) == nkeys and len(m.state_dict()) == nkeys
        for k, v in m.state_dict().items():
            assert torch.equal(v.float(), sd[k].float()), k
    sd_ar = torch.load(helpers.checkpoint_path('disco_ar'), map_location='cpu', weights_only=True)
    m, _, cfg = helpers.make_checkpoint_model('disco')
    ar = dlatent.PretrainedScoreEncoder(pretrained_score_model=m, ns=16, latent_dim=1, latent_vocab=1, latent_hidden_dim=128,
                                        input_latent_dim=cfg.latent_dim, apply_gumbel_softmax=True)
    out = ar.load_state_dict(sd_ar, strict=True)
    assert not out.missing_keys and not out.unexpected_keys
    assert tuple(ar.latent_s_predictor[0].weight.shape) == (128, 32)


# ---------------------------------------------------------------------------------------------- public-API collation
class _PygStore:
    """What a torch_geometric NodeStorage / EdgeStorage offers publicly."""

    def __init__(self, d):
        object.__setattr__(self, '_mapping', dict(d))

    def items(self):
        return self._mapping.items()

    def keys(self):
        return self._mapping.keys()

    def __getattr__(self, k):
        try:
            return object.__getattribute__(self, '_mapping')[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self._mapping[k] = v

    def __contains__(self, k):
        return k in self._mapping

    def __getitem__(self, k):
        return self._mapping[k]

    @property
    def num_nodes(self):
        return int(self._mapping['pos'].shape[0])


class _PygLike:
    """A torch_geometric HeteroData look-alike: node_types / edge_types / item access / to_dict() with the graph attributes
    under '_global_store'; any other attribute raises AttributeError (in particular none of our private fields exist)."""

    def __init__(self, g):
        node, edge, attrs = ddata.public_view(g)
        object.__setattr__(self, '_n', {k: _PygStore(v) for k, v in node.items()})
        object.__setattr__(self, '_e', {k: _PygStore(v) for k, v in edge.items()})
        object.__setattr__(self, '_g', dict(attrs))

    node_types = property(lambda self: list(self._n))
    edge_types = property(lambda self: list(self._e))

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._n[key] if key in self._n else self._g[key]
        key = tuple(key)
        if len(key) == 2:
            key = next(k for k in self._e if k[0] == key[0] and k[2] == key[1])
        return self._e[key]

    def __contains__(self, key):
        return key in self._n or key in self._g

    def to_dict(self):
        out = {'_global_store': dict(self._g)}
        out.update({k: dict(v.items()) for k, v in self._n.items()})
        out.update({k: dict(v.items()) for k, v in self._e.items()})
        return out

    def __getattr__(self, k):
        g = object.__getattribute__(self, '_g')
        if k in g:
            return g[k]
        raise AttributeError(k)

    def __setattr__(self, k, v):
        self._g[k] = v


def test_collation_reads_only_the_public_container_api():
    gs = [synthetic.make_complex(51, 9, 14), synthetic.make_complex(52, 13, 20)]
    for g in gs:
        g['ligand'].latent_h = torch.zeros(g['ligand'].pos.shape[0], 2)
    own_items = [synthetic.as_loader_item(g) for g in gs]
    stubs = [_PygLike(x) for x in own_items]
    assert not hasattr(stubs[0], '_node_stores') and not hasattr(stubs[0], '_attrs')
    a, b = ddata.Batch.from_data_list(own_items), ddata.Batch.from_data_list(stubs)
    assert a.num_graphs == b.num_graphs == 2
    for nt in ('ligand', 'receptor'):
        for k in ('x', 'pos', 'batch', 'ptr'):
            assert torch.equal(a[nt][k], b[nt][k]), (nt, k)
    assert torch.equal(a['ligand'].latent_h, b['ligand'].latent_h)
    for et in (('ligand', 'ligand'), ('receptor', 'receptor')):
        assert torch.equal(a[et].edge_index, b[et].edge_index)
    assert torch.equal(a['ligand', 'ligand'].edge_attr, b['ligand', 'ligand'].edge_attr)
    assert a.name == b.name
    # the loader used for confidence_data_list (utils/sampling.py:58-59) and the copy grouping of the fast path
    batches = list(ddata.DataLoader(stubs, batch_size=2))
    assert len(batches) == 1 and torch.equal(batches[0]['ligand'].pos, a['ligand'].pos)
    groups = dsampling.group_copies([_PygLike(own_items[0]), _PygLike(copy.deepcopy(own_items[0])), stubs[1]])
    assert [n for _, n in groups] == [2, 1]
    from disco_diffdock_b200 import engine as dengine
    h1, m1, r1 = dengine.group_index_arrays([(own_items[0], 2), (own_items[1], 1)], False)
    h2, m2, r2 = dengine.group_index_arrays([(stubs[0], 2), (stubs[1], 1)], False)
    assert r1 == r2 and np.array_equal(m1, m2) and np.array_equal(h1.bond_index, h2.bond_index) and np.array_equal(h1.rec_index, h2.rec_index)
