"""Small driver for ncu captures. Usage: python tools/profile_target.py predict|potrf|fit [N] [M]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from socks_b200 import _device as dv, synthetic as syn
from socks_b200.algorithms.reach.kernel_sr import KernelSR
from socks_b200.kernel import metrics as mt

what = sys.argv[1] if len(sys.argv) > 1 else "predict"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
m = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
T = 15
X, U, Y = syn.integrator_sample(n, seed=0)
if what == "potrf":
    fac = mt.Factorization(dv.to_dev(X), (-50.0, 0), 1.0 / n)
    fac.diag()
    torch.cuda.synchronize()
    fac.check()
else:
    ct, tt = syn.integrator_tubes(T)
    alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", regularization_param=1 / n ** 2)
    alg.fit_arrays(X, Y)
    if what == "predict":
        pts = dv.to_dev(syn.uniform_points(m, seed=1))
        for _ in range(3):
            out = alg.predict_dev(pts)
    torch.cuda.synchronize()
print("ok", what, n, m)
