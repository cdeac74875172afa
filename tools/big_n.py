"""Robustness at the top of one B200's memory: KernelSR fit + predict at large N (default 60 000).
Checks: positive definite factorisation, probabilities in [0, 1], FMA-vs-DMMA agreement on a slice."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from socks_b200 import _device as dv, synthetic as syn
from socks_b200.algorithms.reach.kernel_sr import KernelSR

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
m, T = 250000, 15
X, U, Y = syn.integrator_sample(n, seed=0)
ct, tt = syn.integrator_tubes(T)
alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", regularization_param=1 / n ** 2)
alg._keep_factors = False
torch.cuda.synchronize()
t0 = time.perf_counter()
alg.fit_arrays(X, Y)
torch.cuda.synchronize()
t_fit = time.perf_counter() - t0
pts = dv.to_dev(syn.uniform_points(m, seed=1))
alg.predict_dev(pts[:4096])
torch.cuda.synchronize()
t0 = time.perf_counter()
out = alg.predict_dev(pts)
torch.cuda.synchronize()
t_pred = time.perf_counter() - t0
alg.predict_variant = 1
ref = alg.predict_dev(pts[:20000])
npad = dv.padded(n)
print(json.dumps({"N": n, "fit_s": t_fit, "fit_tflops": 2 * npad ** 3 / 3 / t_fit / 1e12, "predict_s": t_pred,
                  "pts_per_s": m / t_pred, "out_min": float(out.min()), "out_max": float(out.max()),
                  "nan": int(torch.isnan(out).sum()), "dmma_vs_fma_maxdiff": float((out[:, :20000] - ref).abs().max()),
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))
