import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from socks_b200 import _device as dv, synthetic as syn
from socks_b200.algorithms.reach.kernel_sr_max import KernelMaximalSR
n, m, T, P = 20000, 16384, 15, 100
X, U, Y = syn.integrator_sample(n, seed=0)
pts = dv.to_dev(syn.uniform_points(m, seed=1))
A = np.linspace(-1, 1, P).reshape(-1, 1)
ct, tt = syn.integrator_tubes(T)
alg = KernelMaximalSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", regularization_param=1.0)
alg.fit_arrays(X, U, Y, A)
alg.predict_slices = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for _ in range(2):
    out = alg.predict_dev(pts)
torch.cuda.synchronize()
print("ok")
