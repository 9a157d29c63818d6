"""TEST INFRASTRUCTURE: how close is a sampling golden to an edge flip?

Runs the CPU oracle (oracle/restate.py) on a sampling case of oracle/make_golden.py from the nominal start pose and
from start poses perturbed by `eps` Angstrom, and prints the final-pose RMSD of every perturbed run against the
nominal one.  A well-conditioned case keeps all of them within ~2e-4 A; a case next to a radius-graph edge flip shows a
second mode ~3e-3 A away.  Usage: python tools/golden_conditioning.py sample_cfg1 [cseed ...]
"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from disco_diffdock_b200 import data as ddata  # noqa: E402
from oracle import make_golden, restate  # noqa: E402
from tests import helpers  # noqa: E402
from tests.test_oracle_golden import load_tables  # noqa: E402


def run(c, eps, seed, tables):
    m, sd, cfg, lst, noise, sched, temps = make_golden.sample_inputs(c)
    l2 = copy.deepcopy(lst)
    if eps:
        g = torch.Generator().manual_seed(seed)
        for x in l2:
            x['ligand'].pos = x['ligand'].pos + eps * torch.randn(x['ligand'].pos.shape, generator=g)
    batch = ddata.Batch.from_data_list(l2)
    with torch.no_grad():
        return restate.sample(sd, cfg, batch, tables, sched, noise, inference_steps=c['steps'], **temps).clone()


if __name__ == '__main__':
    name = sys.argv[1] if len(sys.argv) > 1 else 'sample_cfg1'
    seeds = [int(v) for v in sys.argv[2:]] or [make_golden.CASES[name]['cseed']]
    tables = load_tables()
    for cseed in seeds:
        c = dict(make_golden.CASES[name]); c['cseed'] = cseed
        base = run(c, 0, 0, tables)
        sp = [float(helpers.rmsd_per_pose(base, run(c, 2e-6, s, tables), c['B']).max()) for s in range(4)]
        print(name, 'cseed', cseed, 'rmsd of 4 runs perturbed by 2e-6 A:', ['%.2e' % v for v in sp], flush=True)
