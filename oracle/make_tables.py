# Generates the SO(3) / torus score-norm LUTs by importing the reference's own utils/so3.py and utils/torus.py
# (so3.py:51-66, torus.py:31-76). numpy is seeded before the torus import because torus.py:72-76 is Monte-Carlo.
import sys, time, numpy as np
sys.path.insert(0, '/root/reference')
t0 = time.time()
np.random.seed(0)
from utils import torus
np.save('torus_score_norm.npy', torus.score_norm_.astype(np.float64))
print('torus done', time.time() - t0, flush=True)
from utils import so3
np.save('so3_exp_score_norms.npy', so3._exp_score_norms.astype(np.float64))
print('so3 done', time.time() - t0, flush=True)
