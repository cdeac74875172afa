"""Achieved GPU-vs-reference errors on every golden fixture (the tests only assert the tolerances).
Usage (on a B200): python tests/measure/parity_report.py [out.json]"""
import json
import os
import sys
from functools import partial

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from conftest import load_golden, tubes_from
from socks_b200 import synthetic as syn
from socks_b200.algorithms.control.kernel_control_bwd import KernelControlBwd
from socks_b200.algorithms.control.kernel_control_fwd import KernelControlFwd
from socks_b200.algorithms.reach.kernel_sr import KernelSR
from socks_b200.algorithms.reach.kernel_sr_max import KernelMaximalSR
from socks_b200.algorithms.reach.separating_kernel import SeparatingKernelClassifier
from socks_b200.kernel.metrics import abel_kernel, rbf_kernel, regularized_inverse

res = {}


def cons(time=0, state=None):
    return np.abs(state[:, 1]) - (0.45 + 0.01 * time)


g = load_golden("kernel_metrics.npz")
res["rbf_kernel max rel err"] = float(np.max(np.abs(rbf_kernel(g["X"], g["Y"], sigma=0.7) - g["K_xy_s07"]) / g["K_xy_s07"]))
W = regularized_inverse(g["X"], U=g["U"], kernel_fn=partial(rbf_kernel, sigma=0.7), regularization_param=1e-3)
res["regularized_inverse rel-Frobenius err (cond %.1e)" % np.linalg.cond(g["W_xu_s07_lam"])] = float(
    np.linalg.norm(W - g["W_xu_s07_lam"]) / np.linalg.norm(g["W_xu_s07_lam"]))
for name in ("sr_c1_fht.npz", "sr_c1_tht.npz", "sr_d3_timevarying_fht.npz", "sr_d4_tht.npz"):
    g = load_golden(name)
    a = KernelSR(time_horizon=int(g["T"]), constraint_tube=tubes_from(g["ctube"]), target_tube=tubes_from(g["ttube"]),
                 problem=str(g["problem"]), kernel_fn=partial(rbf_kernel, sigma=float(g["sigma"])),
                 regularization_param=float(g["lam"]))
    a.fit_arrays(g["X"], g["Y"])
    res[f"KernelSR {name} max abs err (value functions, probabilities)"] = [
        float(np.max(np.abs(a.value_functions - g["value_functions"]))), float(np.max(np.abs(a.predict(g["pts"]) - g["out"])))]
g = load_golden("sr_c2_n20000_fht.npz")
n, T = int(g["n"]), int(g["T"])
X, U, Y = syn.integrator_sample(n, dim=2, seed=0)
ct, tt = syn.integrator_tubes(T)
a = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", kernel_fn=partial(rbf_kernel, sigma=0.1),
             regularization_param=float(g["lam"]))
a.fit_arrays(X, Y)
pts = syn.uniform_points(250000, 2, seed=1)[: int(g["m_sub"])]
res["KernelSR N=20000 (configs[1]) max abs err (value functions, probabilities, diag(W) rel)"] = [
    float(np.max(np.abs(a.value_functions[:, ::10] - g["value_functions_sub"]))),
    float(np.max(np.abs(a.predict(pts) - g["out"]))),
    float(np.max(np.abs(a._w.cpu().numpy()[::10] - g["w_diag_sub"]) / g["w_diag_sub"]))]
for name in ("srmax_fht.npz", "srmax_tht_batch.npz", "srmax_n1200_p100_fht.npz"):
    g = load_golden(name)
    a = KernelMaximalSR(time_horizon=int(g["T"]), constraint_tube=tubes_from(g["ctube"]), target_tube=tubes_from(g["ttube"]),
                        problem=str(g["problem"]), regularization_param=float(g["lam"]),
                        kernel_fn=partial(rbf_kernel, sigma=float(g["sigma"])))
    a.fit_arrays(g["X"], g["U"], g["Y"], g["A"])
    res[f"KernelMaximalSR {name} max abs err (value functions, probabilities)"] = [
        float(np.max(np.abs(a.value_functions - g["value_functions"]))), float(np.max(np.abs(a.predict(g["pts"]) - g["out"])))]
for name in ("fwd_unconstrained.npz", "fwd_constrained.npz", "fwd_tracking_regime.npz"):
    g = load_golden(name)
    T = int(g["T"])
    p = KernelControlFwd(cost_fn=syn.tracking_cost_fn(T), constraint_fn=cons if bool(g["constrained"]) else None,
                         heuristic=True, regularization_param=float(g["lam"]),
                         kernel_fn=partial(rbf_kernel, sigma=float(g["sigma"])), verbose=False)
    p.train_arrays(g["X"], g["U"], g["Y"], g["A"])
    err, same = 0.0, True
    for t in range(T):
        y = p.scores(time=t, state=[g["states"][t]]).cpu().numpy()
        err = max(err, float(np.max(np.abs(y[0] - g["C"][t])) / np.max(np.abs(g["C"][t]))))
        same = same and bool(np.array_equal(p(time=t, state=[g["states"][t]]), g["actions"][t]))
    res[f"KernelControlFwd {name} (cond {float(g['cond']):.1e}) max rel err of scores, actions identical"] = [err, same]
for name in ("bwd_unconstrained.npz", "bwd_constrained.npz"):
    g = load_golden(name)
    T = int(g["T"])
    p = KernelControlBwd(time_horizon=T, cost_fn=syn.tracking_cost_fn(T), constraint_fn=cons if bool(g["constrained"]) else None,
                         heuristic=True, regularization_param=float(g["lam"]),
                         kernel_fn=partial(rbf_kernel, sigma=float(g["sigma"])), verbose=False)
    p.train_arrays(g["X"], g["U"], g["Y"], g["A"])
    same = all(np.array_equal(p(time=t, state=[g["states"][t]]), g["actions"][t]) for t in range(T))
    res[f"KernelControlBwd {name} (cond {float(g['cond']):.1e}) max rel err of value functions, actions identical"] = [
        float(np.max(np.abs(p.value_functions - g["value_functions"])) / np.max(np.abs(g["value_functions"]))), bool(same)]
g = load_golden("separating_ring.npz")
c = SeparatingKernelClassifier(kernel_fn=partial(abel_kernel, sigma=float(g["sigma"])), regularization_param=float(g["lam"])).fit(g["X"])
res[f"SeparatingKernelClassifier (cond {float(g['cond']):.1e}) max abs err of scores, labels identical"] = [
    float(np.max(np.abs(c.scores(g["T"]) - g["C"]))), bool(np.array_equal(c.predict(g["T"]), g["labels"]))]
print(json.dumps(res, indent=1))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
