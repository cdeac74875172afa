"""KernelMaximalSR.predict: integer tensor-core contraction (Ozaki splitting, csrc/ozaki.cu) against the FP64 DMMA path.
Usage: python tools/srmax_i8_check.py [small|c2] [out.json]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from socks_b200 import _device as dv, synthetic as syn
from socks_b200.algorithms.reach.kernel_sr_max import KernelMaximalSR

mode = sys.argv[1] if len(sys.argv) > 1 else "small"
n, m, T, P = (1200, 4000, 6, 100) if mode == "small" else (20000, 250000, 15, 100)
X, U, Y = syn.integrator_sample(n, seed=0)
pts = dv.to_dev(syn.uniform_points(m, seed=1))
A = np.linspace(-1, 1, P).reshape(-1, 1)
ct, tt = syn.integrator_tubes(T)
alg = KernelMaximalSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", regularization_param=1.0)
alg.fit_arrays(X, U, Y, A)


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r


res = []
alg.predict_engine = "f64"
t64, ref = timed(lambda: alg.predict_dev(pts))
res.append({"engine": "f64 DMMA", "N": n, "M": m, "T": T, "P": P, "s": t64, "pts_per_s": m / t64})
print(json.dumps(res[-1]), flush=True)
for S in (8, 7, 6):
    alg.predict_engine, alg.predict_slices = "i8", S
    t, out = timed(lambda: alg.predict_dev(pts))
    err = float((out - ref).abs().max())
    res.append({"engine": f"i8 x {S} slices", "N": n, "M": m, "s": t, "pts_per_s": m / t, "speedup": t64 / t,
                "max_abs_diff_vs_f64": err, "flagged": alg.last_flagged})
    print(json.dumps(res[-1]), flush=True)
if len(sys.argv) > 2:
    json.dump(res, open(sys.argv[2], "w"), indent=1)
