"""ncu target: one fit at configs[1] size, then a few fused predicts.  Usage: python tools/predict_profile_target.py [variant] [M]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from socks_b200 import _device as dv, synthetic as syn
from socks_b200.algorithms.reach.kernel_sr import KernelSR

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
m = int(sys.argv[2]) if len(sys.argv) > 2 else 250000
n, d, T = 20000, 2, 15
X, U, Y = syn.integrator_sample(n, d, seed=0)
ct, tt = syn.integrator_tubes(T, d)
alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", regularization_param=1 / n ** 2)
alg._keep_factors = False
alg.fit_arrays(X, Y)
alg.predict_variant = variant
pts = dv.to_dev(syn.uniform_points(m, d, seed=1))
for _ in range(3):
    out = alg.predict_dev(pts)
torch.cuda.synchronize()
print("ok")
