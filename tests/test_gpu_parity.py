"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the reference-generated golden vectors.

Run on the B200 box:  python -m pytest tests -m gpu -q
Tolerances (fp32 path, different summation order than the reference): scores 2e-5 relative to the largest score
component, node features 2e-5 absolute, final poses <= 1e-3 A RMSD after a full reverse-diffusion run (the tolerance
BASELINE.json's north_star states).  Edge sets (radius / cross graph construction) must match bit-exactly.
"""
import copy
import json
import os

import numpy as np
import pytest
import torch

from disco_diffdock_b200 import data as ddata
from disco_diffdock_b200 import diffusion_utils as du
from disco_diffdock_b200 import sampling as dsampling
from disco_diffdock_b200 import synthetic
from oracle import make_golden, restate
from tests import helpers
from tests.test_oracle_golden import GOLD, load_tables

pytestmark = pytest.mark.gpu
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')


def dump(name, obj):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, f'diag_{name}.json'), 'w') as f:
        json.dump(obj, f, indent=1, default=float)


@pytest.fixture(scope='module', autouse=True)
def _need_cuda():
    assert torch.cuda.is_available(), 'these tests need the B200'
    from disco_diffdock_b200 import build
    build.build()


def rel_err(a, b):
    """max |a - b| relative to the largest reference entry of the tensor (no absolute floor)."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / max(1e-30, float(b.abs().max())))


def oracle_forward(sd, cfg, batch):
    tr = {}
    with torch.no_grad():
        out = restate.forward(sd, cfg, copy.deepcopy(batch), load_tables(), tr)
    return out, tr


def edge_sets_from_engine(eng, info):
    """(src, dst, slot) triples of the four groups as the kernels listed them."""
    cnt = eng.debug_read('seg_cnt', np.int32)
    base = eng.debug_read('seg_base', np.int32)
    lst = eng.debug_read('seg_list', np.int32).reshape(-1, 2)
    NL, NR = info.NL, info.NR
    groups = {0: [], 1: [], 2: [], 3: []}
    for node in range(NL + NR):
        for which in range(2):
            seg = 2 * node + which
            g = which if node < NL else 2 + which
            for e in range(cnt[seg]):
                slot, dst = lst[base[seg] + e]
                groups[g].append((node, int(dst), int(slot)))
    return groups


@pytest.mark.parametrize('layers', [3, 4, 5])
def test_forward_stages(layers):
    """Every stage of one score evaluation against the oracle, for models with 3..5 conv layers
    (3 layers exercise basis levels 0-2, the 4th adds level 3)."""
    m, sd, cfg = helpers.make_model(10 + layers, num_conv_layers=layers)
    m = m.to('cuda')
    _, lst = helpers.make_pose_batch(3, 20, 50, 3)
    batch = ddata.Batch.from_data_list(lst)
    restate.set_time(batch, 0.6, 0.6, 0.6, 3)
    (tr_o, rot_o, tor_o), trc = oracle_forward(sd, cfg, batch)
    tr, rot, tor = m(batch)
    eng = m.engine()
    info = eng.batch_info
    diag = {}
    # --- graph construction: identical edge sets
    groups = edge_sets_from_engine(eng, info)
    NL = info.NL
    rr = batch['receptor', 'receptor'].edge_index
    want = {0: set(zip(trc['ll_src'].tolist(), trc['ll_dst'].tolist())),
            1: set(zip(trc['lr_src'].tolist(), (trc['lr_dst'] + NL).tolist())),
            2: set(zip((rr[0] + NL).tolist(), (rr[1] + NL).tolist())),
            3: set(zip((trc['lr_dst'] + NL).tolist(), trc['lr_src'].tolist()))}
    # group 0 holds bonded pairs twice (bond edge + radius edge): compare as multisets
    from collections import Counter
    got0 = Counter((s, d) for s, d, _ in groups[0])
    want0 = Counter(zip(trc['ll_src'].tolist(), trc['ll_dst'].tolist()))
    diag['edges'] = {g: len(groups[g]) for g in groups}
    assert got0 == want0, 'ligand-ligand edge multiset differs'
    for g in (1, 2, 3):
        assert set((s, d) for s, d, _ in groups[g]) == want[g], f'group {g} edge set differs'
        assert len(groups[g]) == len(want[g])
    assert eng.last_edge_count() == trc['n_edges']
    # --- edge embeddings / harmonics at the listed slots
    ea = eng.debug_read('ea_pool').reshape(-1, 24)
    sh = eng.debug_read('sh_pool').reshape(-1, 4)
    o_lr = {(s, d): i for i, (s, d) in enumerate(zip(trc['lr_src'].tolist(), (trc['lr_dst'] + NL).tolist()))}
    idx_k = np.array([sl for s, d, sl in groups[1]])
    idx_o = np.array([o_lr[(s, d)] for s, d, _ in groups[1]])
    diag['lr_ea'] = float(np.abs(ea[idx_k] - trc['lr_ea'].numpy()[idx_o]).max())
    diag['lr_sh'] = float(np.abs(sh[idx_k] - trc['lr_sh'].numpy()[idx_o]).max())
    o_rr = {(s, d): i for i, (s, d) in enumerate(zip((rr[0] + NL).tolist(), (rr[1] + NL).tolist()))}
    idx_k = np.array([sl for s, d, sl in groups[2]])
    idx_o = np.array([o_rr[(s, d)] for s, d, _ in groups[2]])
    diag['rr_ea'] = float(np.abs(ea[idx_k] - trc['rr_ea'].numpy()[idx_o]).max())
    diag['rr_sh'] = float(np.abs(sh[idx_k] - trc['rr_sh'].numpy()[idx_o]).max())
    # ligand-ligand: radius edges (zero bond attr) and bond edges; match by (src, dst, is_bond)
    nb = batch['ligand', 'ligand'].edge_index.shape[1]
    o_ll = {}
    for i, (s, d) in enumerate(zip(trc['ll_src'].tolist(), trc['ll_dst'].tolist())):
        o_ll[(s, d, i < nb)] = i
    idx_k = np.array([sl for s, d, sl in groups[0]])
    idx_o = np.array([o_ll[(s, d, sl < nb)] for s, d, sl in groups[0]])
    diag['ll_ea'] = float(np.abs(ea[idx_k] - trc['ll_ea'].numpy()[idx_o]).max())
    diag['ll_sh'] = float(np.abs(sh[idx_k] - trc['ll_sh'].numpy()[idx_o]).max())
    # --- node features after the last layer: ligand rows as the score heads saw them, then every node through embed()
    #     (before the heads the last conv layer skips receptor nodes)
    x_score = eng.debug_read('x_final').reshape(-1, 84)[:info.NL].copy()
    m.embed(batch)
    x = eng.debug_read('x_final').reshape(-1, 84)
    diag['x_lig_score_vs_embed'] = float(np.abs(x_score - x[:info.NL]).max())
    assert diag['x_lig_score_vs_embed'] == 0.0
    ref_x = torch.cat([trc['lig_h'], trc['rec_h']]).numpy()
    w = ref_x.shape[1]
    diag['x_final'] = float(np.abs(x[:, :w] - ref_x).max())
    diag['x_pad'] = float(np.abs(x[:, w:]).max()) if w < 84 else 0.0
    diag['x_per_node_max'] = [float(v) for v in np.abs(x[:, :w] - ref_x).max(1)[:8]]
    diag['tr'], diag['rot'], diag['tor'] = rel_err(tr, tr_o), rel_err(rot, rot_o), rel_err(tor, tor_o)
    diag['tr_vals'] = [tr.cpu().tolist(), tr_o.tolist()]
    diag['tor_vals'] = [tor.cpu().tolist()[:6], tor_o.tolist()[:6]]
    dump(f'stages_L{layers}', diag)
    for k in ('lr_ea', 'lr_sh', 'rr_ea', 'rr_sh', 'll_ea', 'll_sh'):
        assert diag[k] < 5e-6, (k, diag[k])
    assert diag['x_final'] < 2e-5 and diag['x_pad'] == 0.0, diag
    assert diag['tr'] < 2e-5 and diag['rot'] < 2e-5 and diag['tor'] < 2e-5, diag


@pytest.mark.parametrize('name', ['forward_small', 'forward_latent', 'forward_cfg1'])
def test_forward_matches_reference_golden(name):
    """GPU forward vs the vectors generated by the reference's own forward (oracle/make_golden.py)."""
    z = np.load(os.path.join(GOLD, name + '.npz'))
    m, sd, cfg, batch = make_golden.forward_inputs(make_golden.CASES[name])
    m = m.to('cuda')
    tr, rot, tor = m(batch)
    lig_h, rec_h = m.embed(batch)[:2]
    d = {'tr': rel_err(tr, torch.from_numpy(z['tr'])), 'rot': rel_err(rot, torch.from_numpy(z['rot'])),
         'tor': rel_err(tor, torch.from_numpy(z['tor'])),
         'lig_h': float((lig_h.cpu() - torch.from_numpy(z['lig_h'])).abs().max()),
         'rec_h': float((rec_h.cpu() - torch.from_numpy(z['rec_h'])).abs().max())}
    dump(name, d)
    assert max(d['tr'], d['rot'], d['tor']) < 2e-5 and max(d['lig_h'], d['rec_h']) < 3e-5, d


def test_forward_mixed_batch_and_no_rotatable_bonds():
    """forward() on a batch of *different* complexes (the reference's forward supports it, only sampling() assumes
    copies) including a ligand without rotatable bonds."""
    m, sd, cfg = helpers.make_model(5)
    m = m.to('cuda')
    gs = [synthetic.make_complex(21, 12, 30), synthetic.make_complex(22, 25, 64), synthetic.make_complex(23, 9, 17)]
    gs[2]['ligand'].edge_mask = torch.zeros_like(gs[2]['ligand'].edge_mask)
    gs[2]['ligand'].mask_rotate = np.zeros((0, 9), dtype=bool)
    batch = ddata.Batch.from_data_list(gs)
    restate.set_time(batch, 0.3, 0.3, 0.3, 3)
    (tr_o, rot_o, tor_o), _ = oracle_forward(sd, cfg, batch)
    tr, rot, tor = m(batch)
    d = {'tr': rel_err(tr, tr_o), 'rot': rel_err(rot, rot_o), 'tor': rel_err(tor, tor_o), 'n_tor': int(tor.numel())}
    dump('mixed', d)
    assert tor.shape == tor_o.shape
    assert max(d['tr'], d['rot'], d['tor']) < 2e-5, d


def test_update_matches_oracle():
    """ddk_update (perturbation + rigid move + sequential torsions + Kabsch) vs modify_conformer_batch."""
    m, sd, cfg = helpers.make_model(6)
    m = m.to('cuda')
    B = 4
    g, lst = helpers.make_pose_batch(8, 24, 40, B)
    batch = ddata.Batch.from_data_list(lst)
    eng = m.engine()
    info = eng.set_batch(batch)
    R = g['ligand'].mask_rotate.shape[0]
    gen = torch.Generator().manual_seed(3)
    tr, rot, tor = torch.randn(B, 3, generator=gen), torch.randn(B, 3, generator=gen) * 0.5, torch.randn(B * R, generator=gen)
    z = helpers.draw_noise(4, 1, B, R, no_final_step_noise=False)
    coef = [0.7, 0.3, 0.4, 0.2, 0.9, 0.5]
    pos = batch['ligand'].pos.clone()
    M = batch['ligand', 'ligand'].edge_index.shape[1] // B
    rot_bonds = batch['ligand', 'ligand'].edge_index[:, :M].T[batch['ligand'].edge_mask[:M]]
    mask_rotate = torch.from_numpy(g['ligand'].mask_rotate)
    want = restate.modify_conformer_batch(pos, B, rot_bonds, mask_rotate,
                                          coef[0] * tr + coef[1] * z['tr'][0], coef[2] * rot + coef[3] * z['rot'][0],
                                          coef[4] * tor + coef[5] * z['tor'][0])
    dev = eng.device
    got = eng.update(pos.to(dev).contiguous(), tr.to(dev), rot.to(dev), tor.to(dev), z['tr'][0].to(dev).contiguous(),
                     z['rot'][0].to(dev).contiguous(), z['tor'][0].to(dev).contiguous(), coef)
    rmsd = helpers.rmsd_per_pose(want, got.cpu(), B)
    dump('update', {'rmsd': rmsd.tolist()})
    assert float(rmsd.max()) < 2e-5, rmsd


@pytest.mark.parametrize('mode', [0, 1, 2])
def test_conv_kernel_paths_match_reference_golden(mode):
    """The three conv-kernel choices (DDK_TC = 0 FFMA2 only, 1 FFMA2 + k_acc_tc, 2 k_conv_tcr) against the same reference
    vectors: config-1 forward (node features, scores) and an 8-step sampling run."""
    from disco_diffdock_b200 import engine as dengine
    dengine.set_tensor_core_path(mode)
    try:
        z = np.load(os.path.join(GOLD, 'forward_cfg1.npz'))
        m, sd, cfg, batch = make_golden.forward_inputs(make_golden.CASES['forward_cfg1'])
        m = m.to('cuda')
        tr, rot, tor = m(batch)
        lig_h, rec_h = m.embed(batch)[:2]
        d = {'tr': rel_err(tr, torch.from_numpy(z['tr'])), 'rot': rel_err(rot, torch.from_numpy(z['rot'])),
             'tor': rel_err(tor, torch.from_numpy(z['tor'])),
             'lig_h': float((lig_h.cpu() - torch.from_numpy(z['lig_h'])).abs().max()),
             'rec_h': float((rec_h.cpu() - torch.from_numpy(z['rec_h'])).abs().max())}
        c = make_golden.CASES['sample_small']
        zs = np.load(os.path.join(GOLD, 'sample_small.npz'))
        m2, sd2, cfg2, lst, noise, sched, temps = make_golden.sample_inputs(c)
        pos = run_gpu_sampling(m2.to('cuda'), cfg2, lst, sched, noise, c['steps'], temps, c['B'])
        d['sample_small_rmsd'] = helpers.rmsd_per_pose(torch.from_numpy(zs['pos']), pos, c['B']).tolist()
        dump(f'conv_path_{mode}', d)
        assert max(d['tr'], d['rot'], d['tor']) < 2e-5 and max(d['lig_h'], d['rec_h']) < 3e-5, d
        assert max(d['sample_small_rmsd']) < 1e-3, d
    finally:
        dengine.set_tensor_core_path(None)


def run_gpu_sampling(m, cfg, lst, sched, noise, steps, temps, B, host_buffers=False):
    from functools import partial
    data_list = [synthetic.as_loader_item(x) for x in copy.deepcopy(lst)]
    t2s = partial(du.t_to_sigma, args=cfg)
    out, _ = dsampling.sampling(data_list, m, steps, sched, sched, sched, torch.device('cuda'), t2s, cfg, batch_size=B,
                                no_final_step_noise=False, noise=noise, host_buffers=host_buffers, **temps)
    return torch.cat([x['ligand'].pos.cpu() for x in out])


@pytest.mark.parametrize('name', ['sample_small', 'sample_mid', 'sample_cfg1'])
def test_sampling_matches_reference_golden(name):
    """Full reverse-diffusion runs through the drop-in sampling() vs poses produced by the reference's sampling()."""
    c = make_golden.CASES[name]
    z = np.load(os.path.join(GOLD, name + '.npz'))
    m, sd, cfg, lst, noise, sched, temps = make_golden.sample_inputs(c)
    m = m.to('cuda')
    pos = run_gpu_sampling(m, cfg, lst, sched, noise, c['steps'], temps, c['B'])
    rmsd = helpers.rmsd_per_pose(torch.from_numpy(z['pos']), pos, c['B'])
    pos_h = run_gpu_sampling(m, cfg, lst, sched, noise, c['steps'], temps, c['B'], host_buffers=True)
    dump(name, {'rmsd_vs_reference': rmsd.tolist(), 'host_vs_device': float((pos_h - pos).abs().max())})
    assert float(rmsd.max()) < 1e-3, rmsd
    assert float((pos_h - pos).abs().max()) == 0.0


def test_sampling_matches_oracle_multi_batch():
    """10 poses in batches of 4 (ragged last batch), 20 steps, README temperatures: GPU vs oracle trajectory."""
    m, sd, cfg = helpers.make_model(1, gain=5.0)
    m = m.to('cuda')
    N, bs, steps = 10, 4, 20
    g, lst = helpers.make_pose_batch(9, 22, 60, N)
    R = g['ligand'].mask_rotate.shape[0]
    noise = helpers.draw_noise(11, steps, N, R)
    sched = du.get_t_schedule(steps)
    from functools import partial
    data_list = [synthetic.as_loader_item(x) for x in copy.deepcopy(lst)]
    out, _ = dsampling.sampling(data_list, m, steps, sched, sched, sched, torch.device('cuda'), partial(du.t_to_sigma, args=cfg),
                                cfg, batch_size=bs, noise=noise, **helpers.README_TEMPS)
    got = torch.cat([x['ligand'].pos.cpu() for x in out])
    want = []
    for b0 in range(0, N, bs):
        b1 = min(N, b0 + bs)
        batch = ddata.Batch.from_data_list(copy.deepcopy(lst[b0:b1]))
        nz = {'tr': noise['tr'][:, b0:b1], 'rot': noise['rot'][:, b0:b1], 'tor': noise['tor'][:, b0 * R:b1 * R]}
        with torch.no_grad():
            want.append(restate.sample(sd, cfg, batch, load_tables(), sched, nz, inference_steps=steps, **helpers.README_TEMPS).clone())
    rmsd = helpers.rmsd_per_pose(torch.cat(want), got, N)
    dump('multi_batch', {'rmsd': rmsd.tolist()})
    assert float(rmsd.max()) < 1e-3, rmsd


def test_disco_ar_latents_and_guidance_match_reference_golden():
    """BASELINE config 4 in miniature: latent-conditioned score model + autoregressive latent sampler (arg-max decoding) +
    classifier-free guidance, against poses / latents produced by the reference's own PretrainedScoreEncoder.encode_ar
    and sampling() (oracle/make_golden.py::write_disco)."""
    from functools import partial
    from disco_diffdock_b200 import latent as dlatent
    c = helpers.DISCO_CASE
    z = np.load(os.path.join(GOLD, 'sample_disco.npz'))
    m, sd, cfg, lst, noise, sched, heads = helpers.disco_inputs(c)
    m = m.to('cuda')
    ar = dlatent.PretrainedScoreEncoder(pretrained_score_model=m, ns=cfg.ns, latent_dim=1, latent_vocab=1, latent_hidden_dim=128,
                                        input_latent_dim=cfg.latent_dim, apply_gumbel_softmax=True)
    missing = ar.load_state_dict(heads, strict=False)
    assert all(k.startswith('pretrained_score_model.') for k in missing.missing_keys) and not missing.unexpected_keys
    ar = ar.to('cuda').eval()
    # latents alone
    probe = ddata.Batch.from_data_list([synthetic.as_loader_item(x) for x in copy.deepcopy(lst)])
    probe['ligand'].pos = probe['ligand'].ar_pos
    lat_l, lat_r = ar.encode_ar(probe.to('cuda'), c['softmax_latent_temperature'])
    assert torch.equal(lat_l.cpu(), torch.from_numpy(z['lat_l'])) and torch.equal(lat_r.cpu(), torch.from_numpy(z['lat_r']))
    # full run: AR latents -> guided reverse diffusion
    data_list = [synthetic.as_loader_item(x) for x in copy.deepcopy(lst)]
    for a, b in zip(data_list, lst):
        a['ligand'].ar_pos = b['ligand'].ar_pos.clone()
    out, _ = dsampling.sampling(data_list, m, c['steps'], sched, sched, sched, torch.device('cuda'), partial(du.t_to_sigma, args=cfg), cfg,
                                batch_size=c['B'], no_final_step_noise=False, noise=noise, ar_model=ar,
                                classifier_free_guidance_weight=c['cfg_weight'], cfg_start=c['cfg_start'], cfg_end=c['cfg_end'],
                                softmax_latent_temperature=c['softmax_latent_temperature'], **helpers.README_TEMPS)
    pos = torch.cat([x['ligand'].pos.cpu() for x in out])
    rmsd = helpers.rmsd_per_pose(torch.from_numpy(z['pos']), pos, c['B'])
    dump('sample_disco', {'rmsd_vs_reference': rmsd.tolist(), 'latent_str': [x.latent_str for x in out]})
    assert [x.latent_str for x in out] == [str(s) for s in z['latent_str']]
    assert float(rmsd.max()) < 1e-3, rmsd


def test_forward_large_receptor_stress():
    """BASELINE config 5 shape: 2000 C-alpha residues, 120-atom ligand, dynamic cross cut-off at t = 0.9 (all 240 000 cross
    pairs listed per direction) -- two poses against the oracle."""
    m, sd, cfg = helpers.make_model(4)
    m = m.to('cuda')
    _, lst = helpers.make_pose_batch(31, 120, 2000, 2)
    batch = ddata.Batch.from_data_list(lst)
    restate.set_time(batch, 0.9, 0.9, 0.9, 2)
    with torch.no_grad():
        tr_o, rot_o, tor_o = restate.forward(sd, cfg, copy.deepcopy(batch), load_tables())[:3]
    tr, rot, tor = m(batch)
    eng = m.engine()
    d = {'tr': rel_err(tr, tr_o), 'rot': rel_err(rot, rot_o), 'tor': rel_err(tor, tor_o), 'edges': eng.last_edge_count()}
    dump('large_receptor', d)
    assert d['edges'] > 2 * 2 * 120 * 2000 * 0.95
    assert max(d['tr'], d['rot'], d['tor']) < 5e-5, d


def test_batched_inference_matches_per_complex_calls():
    """the evaluate.py-style driver sampling several complexes in one call (runs of copies, one ddk batch)
    gives bit-identical poses to one sampling() call per complex, and those match the oracle."""
    from functools import partial
    from disco_diffdock_b200 import inference
    m, sd, cfg = helpers.make_model(2, gain=5.0)
    m = m.to('cuda')
    N, steps = 4, 12
    gs = [synthetic.make_complex(71, 18, 40), synthetic.make_complex(72, 26, 64), synthetic.make_complex(73, 9, 17)]
    gs[2]['ligand'].edge_mask = torch.zeros_like(gs[2]['ligand'].edge_mask)
    gs[2]['ligand'].mask_rotate = np.zeros((0, 9), dtype=bool)
    for g in gs:
        g['ligand'].orig_pos = g['ligand'].pos.numpy().copy()
    complexes = [synthetic.as_loader_item(g) for g in gs]
    Rs = [int(g['ligand'].edge_mask.sum()) for g in gs]
    zs = [helpers.draw_noise(20 + i, steps, N, Rs[i]) for i in range(3)]

    def noise_fn(first, flat):
        sel = zs[first:first + len(flat) // N]
        return {k: torch.cat([z[k] for z in sel], dim=1) for k in ('tr', 'rot', 'tor')}

    t2s = partial(du.t_to_sigma, args=cfg)
    res = {}
    for per_call in (1, 3, 2):
        np.random.seed(5); torch.manual_seed(5)
        res[per_call] = inference.run_inference(complexes, m, cfg, torch.device('cuda'), t2s, samples_per_complex=N,
                                                inference_steps=steps, complexes_per_call=per_call, noise_fn=noise_fn,
                                                no_final_step_noise=True, **helpers.README_TEMPS)
    diffs = {}
    for per_call in (3, 2):
        for i in range(3):
            a = np.stack([g['ligand'].pos.cpu().numpy() for g in res[1]['data_lists'][i]])
            b = np.stack([g['ligand'].pos.cpu().numpy() for g in res[per_call]['data_lists'][i]])
            diffs[f'{per_call}_{i}'] = float(np.abs(a - b).max())
        assert res[per_call]['names'] == ['synth_71', 'synth_72', 'synth_73']
        assert np.array_equal(np.array(res[per_call]['rmsds']), np.array(res[1]['rmsds']))
    # oracle on complex 1 from the same start poses
    np.random.seed(5); torch.manual_seed(5)
    start = []
    for i in range(3):
        dl = [copy.deepcopy(complexes[i]) for _ in range(N)]
        dsampling.randomize_position(dl, False, False, cfg.tr_sigma_max)
        start.append(dl)
    lst = []
    for k in range(N):
        g = copy.deepcopy(gs[1]); g['ligand'].pos = start[1][k]['ligand'].pos.clone(); lst.append(g)
    batch = ddata.Batch.from_data_list(lst)
    sched = du.get_t_schedule(steps)
    with torch.no_grad():
        want = restate.sample(sd, cfg, batch, load_tables(), sched, zs[1], inference_steps=steps, **helpers.README_TEMPS)
    got = torch.cat([g['ligand'].pos.cpu() for g in res[3]['data_lists'][1]])
    rmsd = helpers.rmsd_per_pose(want, got, N)
    dump('batched_inference', {'max_abs_diff_vs_per_complex': diffs, 'rmsd_vs_oracle': rmsd.tolist()})
    assert max(diffs.values()) == 0.0, diffs
    assert float(rmsd.max()) < 1e-3, rmsd


def test_radius_truncation_at_32_neighbours():
    """torch_cluster's max_num_neighbors rule: a compact ligand in which atoms have more than 32 neighbours
    within 5 A (radius_graph asks for 33 incl. the self loop, keeps the first by ascending index) and rotatable-bond
    midpoints with more than 32 atoms in range (score_model.py:425-438).  Edge multiset bit-exact, scores vs oracle."""
    from collections import Counter
    m, sd, cfg = helpers.make_model(31)
    m = m.to('cuda')
    g = synthetic.make_complex(81, 48, 40)
    c = g['ligand'].pos.mean(0, keepdim=True)
    g['ligand'].pos = (g['ligand'].pos - c) * 0.3 + c                                  # 48 atoms inside a ~4 A ball
    lst = []
    gen = torch.Generator().manual_seed(5)
    for k in range(3):
        x = copy.deepcopy(g)
        x['ligand'].pos = x['ligand'].pos + torch.randn(48, 3, generator=gen) * 0.3 + torch.randn(1, 3, generator=gen)
        lst.append(x)
    deg = max(int((torch.cdist(x['ligand'].pos, x['ligand'].pos) < 5.0).sum(1).max()) - 1 for x in lst)
    assert deg > 32, deg                                                             # the rule is exercised
    batch = ddata.Batch.from_data_list(lst)
    restate.set_time(batch, 0.4, 0.4, 0.4, 3)
    (tr_o, rot_o, tor_o), trc = oracle_forward(sd, cfg, batch)
    tr, rot, tor = m(batch)
    eng = m.engine()
    groups = edge_sets_from_engine(eng, eng.batch_info)
    got0 = Counter((s, d) for s, d, _ in groups[0])
    want0 = Counter(zip(trc['ll_src'].tolist(), trc['ll_dst'].tolist()))
    n_radius = len(trc['ll_src']) - batch['ligand', 'ligand'].edge_index.shape[1]
    d = {'max_degree': deg, 'radius_edges': n_radius, 'untruncated': int(sum(int((torch.cdist(x['ligand'].pos, x['ligand'].pos) < 5.0).sum()) - 48 for x in lst)),
         'tr': rel_err(tr, tr_o), 'rot': rel_err(rot, rot_o), 'tor': rel_err(tor, tor_o)}
    dump('truncation', d)
    assert d['radius_edges'] < d['untruncated']                                      # something was cut
    assert got0 == want0, 'ligand-ligand edge multiset differs under truncation'
    assert eng.last_edge_count() == trc['n_edges']
    assert max(d['tr'], d['rot'], d['tor']) < 2e-5, d


def test_sharded_inference_is_bit_identical_to_unsharded():
    """SURVEY 8e / north_star: sharding samples_per_complex x complexes over ranks.  The shards of world = 2 and 3, run one after
    the other on this GPU, reproduce the world = 1 poses BIT FOR BIT (start poses and noise are keyed by (seed, complex, sample);
    a pose does not depend on its batch), and those match the oracle."""
    from functools import partial
    from disco_diffdock_b200 import inference
    m, sd, cfg = helpers.make_model(2, gain=1.0)
    m = m.to('cuda')
    N, steps = 5, 8
    gs = [synthetic.make_complex(91, 14, 40), synthetic.make_complex(92, 22, 64), synthetic.make_complex(93, 10, 24)]
    complexes = [synthetic.as_loader_item(g) for g in gs]
    t2s = partial(du.t_to_sigma, args=cfg)
    kw = dict(samples_per_complex=N, inference_steps=steps, seed=11, no_final_step_noise=True, broadcast_weights=False, **helpers.README_TEMPS)
    full = inference.run_inference_sharded(complexes, m, cfg, torch.device('cuda'), t2s, rank=0, world=1, poses_per_call=400, **kw)
    assert not any(torch.isnan(p).any() for p in full['ligand_pos'])
    diffs = {}
    for world in (2, 3):
        merged = [torch.full_like(p, float('nan')) for p in full['ligand_pos']]
        for r in range(world):
            part = inference.run_inference_sharded(complexes, m, cfg, torch.device('cuda'), t2s, rank=r, world=world, poses_per_call=6,
                                                   gather=False, **kw)
            for ci in range(3):
                mask = ~torch.isnan(part['ligand_pos'][ci][:, 0, 0])
                merged[ci][mask] = part['ligand_pos'][ci][mask]
        diffs[world] = max(float((a - b).abs().max()) for a, b in zip(merged, full['ligand_pos']))
    # oracle on complex 1 from the same start poses and noise
    dl = inference.seeded_start_poses(complexes[1], 1, N, 11, False, False, cfg.tr_sigma_max)
    lst = []
    for k in range(N):
        g = copy.deepcopy(gs[1]); g['ligand'].pos = dl[k]['ligand'].pos.clone(); lst.append(g)
    R = int(gs[1]['ligand'].edge_mask.sum())
    zs = [inference.pose_noise(11, 1, k, steps, R) for k in range(N)]
    noise = {'tr': torch.stack([z['tr'] for z in zs], 1), 'rot': torch.stack([z['rot'] for z in zs], 1), 'tor': torch.cat([z['tor'] for z in zs], 1)}
    for k in noise:
        noise[k][-1] = 0
    with torch.no_grad():
        want = restate.sample(sd, cfg, ddata.Batch.from_data_list(lst), load_tables(), du.get_t_schedule(steps), noise,
                              inference_steps=steps, **helpers.README_TEMPS)
    rmsd = helpers.rmsd_per_pose(want, full['ligand_pos'][1].reshape(-1, 3), N)
    dump('sharded_inference', {'max_abs_diff_vs_unsharded': diffs, 'rmsd_vs_oracle': rmsd.tolist()})
    assert max(diffs.values()) == 0.0, diffs
    assert float(rmsd.max()) < 1e-3, rmsd


def test_sampling_large_receptor_through_late_sparse_steps():
    """BASELINE config 5 shape (2000 C-alpha residues, 120-atom ligand, dynamic cross cut-off) through the reverse-diffusion loop at
    LATE times: t = 0.3, 0.2, 0.1 give a cross cut-off of 21.4 .. 20.6 A, so only a fraction of the receptor is listed (sparse cross
    graph, empty cross segments, receptor-contact hops of the needed-residue sets) -- against the oracle on the same noise."""
    from functools import partial
    m, sd, cfg = helpers.make_model(4, gain=2.0)
    m = m.to('cuda')
    B, steps = 2, 3
    g, lst = helpers.make_pose_batch(33, 120, 2000, B)
    R = g['ligand'].mask_rotate.shape[0]
    noise = helpers.draw_noise(13, steps, B, R)
    sched = np.array([0.3, 0.2, 0.1])
    data_list = [synthetic.as_loader_item(x) for x in copy.deepcopy(lst)]
    out, _ = dsampling.sampling(data_list, m, steps, sched, sched, sched, torch.device('cuda'), partial(du.t_to_sigma, args=cfg), cfg,
                                batch_size=B, noise=noise, **helpers.README_TEMPS)
    got = torch.cat([x['ligand'].pos.cpu() for x in out])
    eng = m.engine()
    edges_last = eng.last_edge_count()
    batch = ddata.Batch.from_data_list(copy.deepcopy(lst))
    with torch.no_grad():
        want = restate.sample(sd, cfg, batch, load_tables(), sched, noise, inference_steps=steps, **helpers.README_TEMPS)
    rmsd = helpers.rmsd_per_pose(want, got, B)
    dump('large_receptor_sparse_steps', {'rmsd': rmsd.tolist(), 'edges_last_step': edges_last, 'all_pairs_would_be': 2 * B * 120 * 2000})
    assert edges_last < 0.5 * 2 * B * 120 * 2000                     # the cross graph is sparse at these times
    assert float(rmsd.max()) < 1e-3, rmsd
