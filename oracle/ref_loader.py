"""TEST INFRASTRUCTURE ONLY -- import the reference's own model / sampler files (build container only).

``/root/reference`` is mounted read-only in the build container and does not exist on the GPU box.  This
loader installs the leaf-op shims (``oracle/ref_shims.py``), puts the reference on ``sys.path`` and exposes
its unmodified ``TensorProductScoreModel`` (models/score_model.py), ``sampling`` (utils/sampling.py) and
``modify_conformer_batch`` (utils/diffusion_utils.py).  ``utils/so3.py`` and ``utils/torus.py`` read/write
multi-hundred-MB caches relative to the CWD and take minutes to build (so3.py:46-66, torus.py:31-40), so
they are replaced by light modules serving the committed score-norm tables
(``disco_diffdock_b200/tables/*.npy``, generated once from those very files by
``oracle/make_tables.py``).
"""
from __future__ import annotations

import os
import sys
import types
from argparse import Namespace
from functools import partial

import numpy as np
import torch

REF = os.environ.get('DDK_REFERENCE', '/root/reference')
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available():
    return os.path.isdir(os.path.join(REF, 'models'))


def load_tables():
    d = os.path.join(_REPO, 'disco_diffdock_b200', 'tables')
    return {'so3_exp_score_norms': np.load(os.path.join(d, 'so3_exp_score_norms.npy')),
            'torus_score_norm': np.load(os.path.join(d, 'torus_score_norm.npy'))}


_mods = None


def modules():
    """Returns a namespace with the reference modules (imported once)."""
    global _mods
    if _mods is not None:
        return _mods
    assert available(), f'reference not found at {REF}'
    from oracle import ref_shims
    ref_shims.install()
    tables = load_tables()

    # utils.so3 / utils.torus: same lookup arithmetic as so3.py:91-95 and torus.py:79-83 on the exported tables
    utils_pkg = types.ModuleType('utils')
    utils_pkg.__path__ = [os.path.join(REF, 'utils')]
    sys.modules['utils'] = utils_pkg
    so3 = types.ModuleType('utils.so3')

    def so3_score_norm(eps):
        eps = eps.numpy()
        idx = (np.log10(eps) - np.log10(0.01)) / (np.log10(2) - np.log10(0.01)) * 1000
        idx = np.clip(np.around(idx).astype(int), a_min=0, a_max=999)
        return torch.from_numpy(tables['so3_exp_score_norms'][idx]).float()
    so3.score_norm = so3_score_norm
    torus = types.ModuleType('utils.torus')

    def torus_score_norm(sigma):
        sigma = np.log(sigma / np.pi)
        sigma = (sigma - np.log(3e-3)) / (np.log(2) - np.log(3e-3)) * 5000
        sigma = np.round(np.clip(sigma, 0, 5000)).astype(int)
        return tables['torus_score_norm'][sigma]
    torus.score_norm = torus_score_norm
    sys.modules['utils.so3'], sys.modules['utils.torus'] = so3, torus
    utils_pkg.so3, utils_pkg.torus = so3, torus

    if REF not in sys.path:
        sys.path.insert(0, REF)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        from models import score_model as sm
        from models import tensor_layers as tl
        from utils import diffusion_utils as du
        from utils import sampling as sp
        from utils import geometry as geo
        from utils import torsion as tor
    _mods = Namespace(score_model=sm, tensor_layers=tl, diffusion_utils=du, sampling=sp, geometry=geo, torsion=tor,
                      tables=tables)
    return _mods


def build_reference_model(cfg, state_dict=None):
    """Instantiate the reference TensorProductScoreModel exactly as utils/model_utils.py:24-68 does."""
    m = modules()
    args = Namespace(**vars(cfg))
    t2s = partial(m.diffusion_utils.t_to_sigma, args=args)
    emb = m.diffusion_utils.get_timestep_embedding('sinusoidal', cfg.sigma_embed_dim, cfg.embedding_scale)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        model = m.score_model.TensorProductScoreModel(
            t_to_sigma=t2s, device=torch.device('cpu'), no_torsion=cfg.no_torsion, timestep_emb_func=emb,
            num_conv_layers=cfg.num_conv_layers, lig_max_radius=cfg.lig_max_radius, scale_by_sigma=cfg.scale_by_sigma,
            sigma_embed_dim=cfg.sigma_embed_dim, ns=cfg.ns, nv=cfg.nv, distance_embed_dim=cfg.distance_embed_dim,
            cross_distance_embed_dim=cfg.cross_distance_embed_dim, batch_norm=True, dropout=0.1, sh_lmax=cfg.sh_lmax,
            use_second_order_repr=False, cross_max_distance=cfg.cross_max_distance,
            dynamic_max_cross=cfg.dynamic_max_cross, lm_embedding_type='esm', use_old_atom_encoder=False,
            latent_dim=cfg.latent_dim, latent_vocab=cfg.latent_vocab, latent_droprate=cfg.latent_droprate)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    model.eval()
    return model, args


class InjectedNormal:
    """Context manager replacing ``torch.normal`` so the reference's ``sampling()`` (sampling.py:146-165)
    consumes pre-drawn noise in its own draw order (tr, rot, tor per step)."""

    def __init__(self, noise, steps):
        self.queue = []
        for s in range(steps):
            self.queue += [noise['tr'][s], noise['rot'][s], noise['tor'][s]]
        self.i = 0

    def __enter__(self):
        self._orig = torch.normal

        def fake(mean=0, std=1, size=None, device=None, **kw):
            z = self.queue[self.i]
            self.i += 1
            assert tuple(z.shape) == tuple(size), (z.shape, size)
            return z.clone()
        torch.normal = fake
        return self

    def __exit__(self, *a):
        torch.normal = self._orig
