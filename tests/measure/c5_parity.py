"""Full-size parity check of BASELINE.json configs[4] (N = 40 000, d = 4, T = 15): GPU KernelSR vs the CPU oracle
(numpy LU inverse of the 40 000 x 40 000 regularised Gram, ~5 min on the box's cores) on a sample of the test points.
Usage (on a B200, >= 64 GB host RAM): python tests/measure/c5_parity.py [N] [out.json]"""
import json
import os
import sys
import time
from functools import partial

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

from oracle import socks_oracle as orc
from socks_b200 import synthetic as syn
from socks_b200.algorithms.reach.kernel_sr import KernelSR
from socks_b200.kernel.metrics import rbf_kernel

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 40000
outp = [a for a in sys.argv[1:] if a.endswith(".json")]
d, T, m = 4, 15, 1000
X, U, Y = syn.integrator_sample(n, dim=d, seed=0)
pts = syn.uniform_points(4000000, d, seed=1)[:m]
ct, tt = syn.integrator_tubes(T, d)
t0 = time.perf_counter()
a = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", kernel_fn=partial(rbf_kernel, sigma=0.1),
             regularization_param=1 / n ** 2)
a.fit_arrays(X, Y)
got = a.predict(pts)
t_gpu = time.perf_counter() - t0
t0 = time.perf_counter()
o = orc.KernelSROracle(T, ct, tt, "FHT", 0.1, 1 / n ** 2).fit(X, U, Y)
want = o.predict(pts)
t_cpu = time.perf_counter() - t0
res = {"N": n, "d": d, "T": T, "points_compared": m, "gpu_fit_predict_s": t_gpu, "cpu_oracle_fit_predict_s": t_cpu,
       "max_abs_err_probabilities": float(np.nanmax(np.abs(got - want))),
       "max_abs_err_value_functions": float(np.nanmax(np.abs(a.value_functions - o.value_functions))),
       "max_rel_err_diag_W": float(np.max(np.abs(a._w.cpu().numpy() - o.w_diag) / o.w_diag)),
       "nan_positions_equal": bool(np.array_equal(np.isnan(got), np.isnan(want))), "cores": os.cpu_count()}
print(json.dumps(res))
if outp:
    json.dump(res, open(outp[0], "w"), indent=1)
