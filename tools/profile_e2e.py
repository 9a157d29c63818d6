"""cProfile of the end-to-end path of bench.py (sampling() with host buffers, one complex at a time)."""
import copy, cProfile, os, pstats, sys, time
from functools import partial
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from disco_diffdock_b200 import diffusion_utils as du, sampling as dsampling
from tests import helpers
dev = torch.device('cuda')
m, sd, cfg = helpers.make_model(0, gain=5.0)
m = m.to(dev); m.engine(dev)
t2s = partial(du.t_to_sigma, args=cfg)
sched = du.get_t_schedule(bench.REV_STEPS)
complexes = bench.build_workload(0, 4)
def step(seed):
    g = torch.Generator().manual_seed(seed)
    for data_list in complexes:
        dl = [x.shallow_copy() for x in data_list]
        dsampling.sampling(dl, m, bench.REV_STEPS, sched, sched, sched, dev, t2s, cfg, batch_size=bench.N_SAMPLES,
                           no_final_step_noise=True, generator=g, host_buffers=True, **helpers.README_TEMPS)
step(0); step(1)
torch.cuda.synchronize(); t0 = time.perf_counter(); step(2); torch.cuda.synchronize(); print('4 complexes e2e s:', time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); step(3); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
