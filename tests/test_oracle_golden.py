"""Replays the reference-generated golden vectors (oracle/make_golden.py) through the oracle restatement.
Runs anywhere (no /root/reference needed): this is what pins the oracle on the GPU box."""
import copy
import os

import numpy as np
import pytest
import torch

from disco_diffdock_b200 import data as ddata
from oracle import make_golden, restate
from tests import helpers

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def load_tables():
    d = os.path.join(os.path.dirname(GOLD), '..', 'disco_diffdock_b200', 'tables')
    return {'so3_exp_score_norms': np.load(os.path.join(d, 'so3_exp_score_norms.npy')),
            'torus_score_norm': np.load(os.path.join(d, 'torus_score_norm.npy'))}


def check_fp(z, sd, pos, recx):
    assert abs(make_golden.fingerprint(sd) - float(z['weights_fp'])) <= 1e-6 * float(z['weights_fp']), 'weights drifted'
    assert abs(float(pos.double().abs().sum()) - float(z['pos_fp'])) <= 1e-6 * float(z['pos_fp']), 'inputs drifted'
    assert abs(float(recx.double().abs().sum()) - float(z['rec_fp'])) <= 1e-6 * float(z['rec_fp']), 'inputs drifted'


@pytest.mark.parametrize('name', ['forward_small', 'forward_latent', 'forward_cfg1'])
def test_forward_golden(name):
    z = np.load(os.path.join(GOLD, name + '.npz'))
    m, sd, cfg, batch = make_golden.forward_inputs(make_golden.CASES[name])
    check_fp(z, sd, batch['ligand'].pos, batch['receptor'].x)
    tr = {}
    with torch.no_grad():
        out = restate.forward(sd, cfg, batch, load_tables(), tr)
    for a, k in zip(out, ['tr', 'rot', 'tor']):
        ref = torch.from_numpy(z[k])
        assert float((a - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max())), k
    assert float((tr['lig_h'] - torch.from_numpy(z['lig_h'])).abs().max()) < 1e-5
    assert float((tr['rec_h'] - torch.from_numpy(z['rec_h'])).abs().max()) < 1e-5


@pytest.mark.parametrize('name', ['sample_small', 'sample_mid'])
def test_sample_golden(name):
    c = make_golden.CASES[name]
    z = np.load(os.path.join(GOLD, name + '.npz'))
    m, sd, cfg, lst, noise, sched, temps = make_golden.sample_inputs(c)
    check_fp(z, sd, torch.cat([x['ligand'].pos for x in lst]), lst[0]['receptor'].x)
    batch = ddata.Batch.from_data_list(copy.deepcopy(lst))
    with torch.no_grad():
        pos = restate.sample(sd, cfg, batch, load_tables(), sched, noise, inference_steps=c['steps'], **temps)
    rmsd = helpers.rmsd_per_pose(torch.from_numpy(z['pos']), pos, c['B'])
    assert float(rmsd.max()) < 5e-4, rmsd


@pytest.mark.skipif(helpers.checkpoint_path('diffdockS') is None, reason='shipped checkpoints not on this box')
@pytest.mark.parametrize('name', ['pre_forward', 'pre_forward_disco'])
def test_pretrained_forward_golden(name):
    """The oracle with the SHIPPED checkpoints against the reference's own forward (activations up to 7e5): relative to the
    largest entry of each tensor."""
    c = helpers.PRE_CASES[name]
    z = np.load(os.path.join(GOLD, name + '.npz'))
    sd, cfg = helpers.load_checkpoint(c['ckpt'])
    i = len(c['ts']) - 1
    batch = helpers.pre_forward_batch(c, c['ts'][i])
    assert float((batch['ligand'].pos - torch.from_numpy(z[f'pos{i}'])).abs().max()) == 0.0
    tr = {}
    with torch.no_grad():
        out = restate.forward(sd, cfg, batch, load_tables(), tr)
    for a, k in zip(out, ['tr', 'rot', 'tor']):
        ref = torch.from_numpy(z[f'{k}{i}'])
        assert float((a - ref).abs().max()) <= 2e-5 * float(ref.abs().max()), k
    ref = torch.from_numpy(z[f'lig_h{i}'])
    assert float((tr['lig_h'] - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


@pytest.mark.skipif(helpers.checkpoint_path('diffdockS') is None, reason='shipped checkpoints not on this box')
def test_pretrained_trajectory_golden_first_steps():
    """First reverse steps of the pretrained ODE trajectory: oracle scores at the reference's own poses (teacher-forced)."""
    c = helpers.PRE_CASES['pre_traj_ode']
    z = np.load(os.path.join(GOLD, 'pre_traj_ode.npz'))
    sd, cfg = helpers.load_checkpoint(c['ckpt'])
    g, lst, noise, sched, kw = helpers.pre_traj_inputs(c)
    assert float((torch.cat([x['ligand'].pos for x in lst]) - torch.from_numpy(z['start'])).abs().max()) == 0.0
    assert float(z['oracle_vs_reference_rmsd']) < 1e-3 and float(z['oracle_spread_2e-6'].max()) < 1e-3
    for s in (0, 10, 19):
        batch = ddata.Batch.from_data_list(copy.deepcopy(lst))
        batch['ligand'].pos = torch.from_numpy(z['pos_steps'][s]).clone()
        restate.set_time(batch, sched[s], sched[s], sched[s], c['B'])
        with torch.no_grad():
            out = restate.forward(sd, cfg, batch, load_tables())
        for a, k in zip(out, ['tr', 'rot', 'tor']):
            ref = torch.from_numpy(z[k][s])
            assert float((a - ref).abs().max()) <= 2e-5 * float(ref.abs().max()), (s, k)
