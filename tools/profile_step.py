"""Small driver for ncu: a few reverse steps over a resident batch (2 complexes x 40 samples by default)."""
import argparse
import copy
import os
import sys
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from disco_diffdock_b200 import data as ddata, diffusion_utils as du, sampling as dsampling, synthetic  # noqa: E402
from tests import helpers  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--complexes', type=int, default=2)
ap.add_argument('--samples', type=int, default=40)
ap.add_argument('--rev-steps', type=int, default=2)
ap.add_argument('--start-step', type=int, default=0)
ap.add_argument('--pocket', action='store_true', help='keep ligands at the protein centre (dense cross graph)')
args = ap.parse_args()
dev = torch.device('cuda')
m, sd, cfg = helpers.make_model(0, gain=5.0)
m = m.to(dev)
eng = m.engine(dev)
np.random.seed(0); torch.manual_seed(0)
flat = []
for c in range(args.complexes):
    item = synthetic.as_loader_item(synthetic.make_complex(1000 + c, 60, 300))
    dl = [copy.deepcopy(item) for _ in range(args.samples)]
    dsampling.randomize_position(dl, False, args.pocket, 19.0)
    flat += dl
big = ddata.Batch.from_data_list(flat)
info = eng.set_batch(big)
sched = du.get_t_schedule(20)[args.start_step:]
tab = dsampling.build_step_tables(m, cfg, partial(du.t_to_sigma, args=cfg), sched, sched, sched, args.rev_steps, info.B,
                                  **helpers.README_TEMPS)
pos = big['ligand'].pos.to(dev).contiguous()
eng.sample(pos, tab, None)
torch.cuda.synchronize()
print('edges last step', eng.last_edge_count(), 'per pose', eng.last_edge_count() / info.B)
