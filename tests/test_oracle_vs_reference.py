"""Pins the oracle (oracle/restate.py) to the reference's own code executed behind leaf-op shims.

Runs only where /root/reference is mounted (the build container); the GPU box replays the committed
golden vectors instead (tests/test_oracle_golden.py)."""
import copy
from argparse import Namespace
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from disco_diffdock_b200 import data as ddata
from oracle import ref_loader, restate
from tests import helpers

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason='reference tree not mounted')


@pytest.mark.parametrize('latent', [0, 2])
def test_forward_matches_reference(latent):
    m, sd, cfg = helpers.make_model(0, latent_dim=latent, latent_droprate=0.1 if latent else 0.0)
    ref_model, _ = ref_loader.build_reference_model(cfg, sd)
    tables = ref_loader.load_tables()
    _, lst = helpers.make_pose_batch(3, 20, 50, 3)
    batch = ddata.Batch.from_data_list(lst)
    restate.set_time(batch, 0.6, 0.6, 0.6, 3)
    if latent:
        gen = torch.Generator().manual_seed(5)
        for nt in ('ligand', 'receptor'):
            batch[nt].latent_h = (torch.rand(batch[nt].num_nodes, latent, generator=gen) > 0.9).float()
            batch[nt].unconditional = (torch.rand(batch[nt].num_nodes, 1, generator=gen) > 0.5).float()
    b2 = copy.deepcopy(batch)
    with torch.no_grad():
        ref = ref_model(batch)
        mine = restate.forward(sd, cfg, b2, tables)
    for a, b in zip(ref, mine):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) <= 2e-6 * max(1.0, float(a.abs().max()))


def test_sampling_loop_matches_reference():
    """The reference's unmodified sampling() (utils/sampling.py:49-249) with torch.normal replaced by
    pre-drawn noise == oracle.restate.sample on the same noise, low-temperature sampling on."""
    m, sd, cfg = helpers.make_model(1, gain=5.0)
    ref_model, args = ref_loader.build_reference_model(cfg, sd)
    mods = ref_loader.modules()
    tables = ref_loader.load_tables()
    B, steps = 2, 8
    g, lst = helpers.make_pose_batch(4, 14, 40, B)
    R = g['ligand'].mask_rotate.shape[0]
    noise = helpers.draw_noise(7, steps, B, R)
    sched = mods.diffusion_utils.get_t_schedule(steps)
    from disco_diffdock_b200.synthetic import as_loader_item
    ref_list = [as_loader_item(x) for x in copy.deepcopy(lst)]
    wrapper = SimpleNamespace(score_model=ref_model)
    from functools import partial
    t2s = partial(mods.diffusion_utils.t_to_sigma, args=args)
    with ref_loader.InjectedNormal(noise, steps):
        out_list, _ = mods.sampling.sampling(ref_list, wrapper, steps, sched, sched, sched, torch.device('cpu'), t2s,
                                             args, batch_size=B, no_final_step_noise=False, **helpers.README_TEMPS)
    ref_pos = torch.cat([x['ligand'].pos for x in out_list])
    batch = ddata.Batch.from_data_list(copy.deepcopy(lst))
    with torch.no_grad():
        pos = restate.sample(sd, cfg, batch, tables, sched, noise, inference_steps=steps, **helpers.README_TEMPS)
    rmsd = helpers.rmsd_per_pose(ref_pos, pos, B)
    assert float(rmsd.max()) < 2e-4, rmsd


def test_ar_encoder_has_the_reference_parameter_names():
    """disco_diffdock_b200.latent.PretrainedScoreEncoder must load the reference's AR checkpoints: same head parameter
    names and shapes as models/pretrained_score_encoder.py:23-44."""
    import contextlib, io
    from disco_diffdock_b200 import latent as dlatent
    ref_loader.modules()
    with contextlib.redirect_stdout(io.StringIO()):
        from models.pretrained_score_encoder import PretrainedScoreEncoder as RefEnc
    m, sd, cfg = helpers.make_model(3, latent_dim=2, latent_droprate=0.1)
    ref_model, _ = ref_loader.build_reference_model(cfg, sd)
    ours = dlatent.PretrainedScoreEncoder(m, 24, 1, 1, input_latent_dim=2)
    ref = RefEnc(pretrained_score_model=ref_model, ns=24, latent_dim=1, latent_vocab=1, input_latent_dim=2)
    ko = {k: tuple(v.shape) for k, v in ours.state_dict().items() if k.startswith('latent_')}
    kr = {k: tuple(v.shape) for k, v in ref.state_dict().items() if k.startswith('latent_')}
    assert ko == kr and len(ko) == 2 * (3 * 2 + 2 * 5)
