"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck).
Usage: compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys
from functools import partial

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from socks_b200 import synthetic as syn
from socks_b200.algorithms.control.kernel_control_bwd import KernelControlBwd
from socks_b200.algorithms.control.kernel_control_fwd import KernelControlFwd
from socks_b200.algorithms.reach.kernel_sr import KernelSR
from socks_b200.algorithms.reach.kernel_sr_max import KernelMaximalSR
from socks_b200.algorithms.reach.separating_kernel import SeparatingKernelClassifier
from socks_b200.kernel.metrics import abel_kernel, rbf_kernel, regularized_inverse

n, T = 300, 5
X, U, Y = syn.integrator_sample(n, seed=0)
pts = syn.uniform_points(333, seed=1)
ct, tt = syn.integrator_tubes(T)
print("rbf", rbf_kernel(X[:37], Y[:23], sigma=0.3).sum(), rbf_kernel(X[:50]).sum())
print("W", np.abs(regularized_inverse(X, U=U, regularization_param=1e-2, kernel_fn=partial(rbf_kernel, sigma=0.5))).sum())
a = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", regularization_param=1e-2)
a.fit_arrays(X, Y)
print("sr", a.predict(pts).sum(), a.W.shape)
a.predict_variant = 1
print("sr fma", a.predict(pts).sum())
b = KernelMaximalSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="THT", regularization_param=1e-1)
b.fit_arrays(X, U, Y, np.linspace(-1, 1, 7).reshape(-1, 1))
print("srmax", b.predict(pts).sum())
Xu, Uu, Yu = syn.unicycle_sample(260, seed=3)
A = syn.unicycle_actions(5, 6)
cons = lambda time=0, state=None: np.abs(state[:, 1]) - 0.5
f = KernelControlFwd(cost_fn=syn.tracking_cost_fn(4), constraint_fn=cons, heuristic=True, regularization_param=1e-3,
                     kernel_fn=partial(rbf_kernel, sigma=1.0), verbose=False)
f.train_arrays(Xu, Uu, Yu, A)
print("fwd", f(time=1, state=[[0.1, 0.2, 0.3]]))
g = KernelControlBwd(time_horizon=4, cost_fn=syn.tracking_cost_fn(4), constraint_fn=cons, heuristic=True,
                     regularization_param=1e-3, kernel_fn=partial(rbf_kernel, sigma=1.0), verbose=False)
g.train_arrays(Xu, Uu, Yu, A)
print("bwd", g(time=1, state=[[0.1, 0.2, 0.3]]))
c = SeparatingKernelClassifier(kernel_fn=partial(abel_kernel, sigma=0.3), regularization_param=1e-2).fit(X)
print("sep", c.predict(pts).sum())
