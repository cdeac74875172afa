"""Full-size golden for BASELINE.json configs[1] (N=20 000, T=15, FHT, sigma=0.1, lambda=1/N^2), made by
EXECUTING THE UNMODIFIED REFERENCE (fit ~3 min on 8 cores, ~20 GB RAM).  Stores the fitted value functions
subsampled (every 10th sample) plus the safety probabilities of 2 000 of the 250 000 test points.
Run in the authoring container only:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_c2.py"""
import os
import sys
import time
import warnings
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "gym_shim"))
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import gym  # noqa: E402
from gym_socks.kernel.metrics import rbf_kernel  # noqa: E402
from gym_socks.algorithms.reach.kernel_sr import KernelSR  # noqa: E402

from socks_b200 import synthetic as syn  # noqa: E402

n, T, m_sub = 20000, 15, 2000
X, U, Y = syn.integrator_sample(n, dim=2, seed=0)
pts = syn.uniform_points(250000, 2, seed=1)[:m_sub]
box = lambda b: gym.spaces.Box(low=-b, high=b, shape=(2,), dtype=np.float32)
alg = KernelSR(time_horizon=T, constraint_tube=[box(1.0)] * T, target_tube=[box(0.5)] * T, problem="FHT",
               kernel_fn=partial(rbf_kernel, sigma=0.1), regularization_param=1 / n ** 2)
t0 = time.time()
alg.fit(list(zip(X, U, Y)))
t1 = time.time()
out = alg.predict(pts)
t2 = time.time()
print(f"reference fit {t1 - t0:.1f} s, predict({m_sub}) {t2 - t1:.1f} s")
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sr_c2_n20000_fht.npz"),
                    n=n, T=T, sigma=0.1, lam=1 / n ** 2, m_sub=m_sub, vf_stride=10,
                    w_diag_sub=np.diag(alg.W)[::10].copy(), value_functions_sub=alg.value_functions[:, ::10].copy(),
                    out=out, ref_fit_s=t1 - t0, ref_predict_s=t2 - t1, cond_est=0.0)
print("wrote sr_c2_n20000_fht.npz")
