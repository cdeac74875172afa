"""Host-to-host predict() latency for the per-rank shard sizes of the strong-scaling run (one GPU), for the column-block
counts and the direct-to-host store option of KernelSR.predict.  Usage: python tools/e2e_small.py [out.json]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from socks_b200 import synthetic as syn
from socks_b200.algorithms.reach.kernel_sr import KernelSR

n, T, d = 20000, 15, 2
X, U, Y = syn.integrator_sample(n, d, seed=0)
ct, tt = syn.integrator_tubes(T, d)
alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", regularization_param=1 / n ** 2)
alg._keep_factors = False
alg.fit_arrays(X, Y)
res = []
for m in (31250, 62500, 125000, 250000):
    pts = torch.from_numpy(syn.uniform_points(m, d, seed=1)).pin_memory().numpy()
    ref = None
    for label, chunks, direct in (("1 block", 1, False), ("2 blocks", 2, False), ("3 blocks", 3, False),
                                  ("4 blocks", 4, False), ("auto", None, False), ("direct host stores", None, True)):
        alg.predict_chunks, alg.predict_direct_host = chunks, direct
        for _ in range(3):
            out = alg.predict(pts)
        if ref is None:
            ref = out.copy()
        torch.cuda.synchronize()
        ts = []
        for _ in range(15):
            t0 = time.perf_counter()
            out = alg.predict(pts)
            ts.append(time.perf_counter() - t0)
        rec = {"M": m, "mode": label, "ms_median": 1e3 * float(np.median(ts)), "ms_min": 1e3 * float(np.min(ts)),
               "Mpts_per_s": m / float(np.median(ts)) / 1e6, "identical": bool(np.array_equal(out, ref, equal_nan=True))}
        print(json.dumps(rec), flush=True)
        res.append(rec)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
