"""CPU baselines of the other BASELINE.json configs: the UNMODIFIED reference (gym_socks via oracle/ref_loader) on this
box's host cores, float64, all BLAS threads (SURVEY 8(d) "CPU baseline beside it").  Reported beside the GPU numbers of
tests/measure/config_times.py; never a target.  Usage: python tests/measure/cpu_baselines.py [c3 c4 c5] [out.json]"""
import json
import os
import sys
import time
from functools import partial

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

from oracle import ref_loader
from socks_b200 import synthetic as syn


def wall(fn):
    t0 = time.perf_counter()
    r = fn()
    return time.perf_counter() - t0, r


def c3(n=5000, P_grid=(25, 40), T=20):
    """configs[2]: KernelControlFwd train + policy calls (examples/control/tracking.py regime), full size."""
    ref_loader.load()
    from gym_socks.algorithms.control.kernel_control_fwd import KernelControlFwd
    from gym_socks.kernel.metrics import rbf_kernel

    X, U, Y = syn.unicycle_sample(n, seed=12345)
    A = syn.unicycle_actions(*P_grid)
    pol = KernelControlFwd(cost_fn=syn.tracking_cost_fn(T), heuristic=True, regularization_param=1e-7,
                           kernel_fn=partial(rbf_kernel, sigma=3.0), verbose=False)
    t_train, _ = wall(lambda: pol.train(list(zip(X, U, Y)), A))
    state = np.array([[-0.8, 0.0, 0.0]])
    ts = [wall(lambda: pol(time=t, state=state))[0] for t in range(3)]
    return {"N": n, "P": len(A), "reference_train_s": t_train, "reference_call_s_median": float(np.median(ts))}


def c4(n=20000):
    """configs[3]: regularized_inverse; measured at N = 20 000, N = 40 000 extrapolated x8 (2 N^3 flops)."""
    ref = ref_loader.load()
    from sklearn.metrics.pairwise import rbf_kernel as sk_rbf

    X, _, _ = syn.integrator_sample(n, seed=0)
    t, W = wall(lambda: ref["regularized_inverse"](X, regularization_param=1 / n ** 2,
                                                   kernel_fn=partial(sk_rbf, gamma=50.0)))
    return {"N_measured": n, "reference_regularized_inverse_s": t, "extrapolated_N40000_s": 8 * t,
            "kernel_fn": "sklearn rbf_kernel(gamma=50)", "W_shape": list(W.shape)}


def c5(n=40000, d=4, T=15, sample=2000):
    """configs[4]: KernelSR.predict at N = 40 000, d = 4 on a subset of the 4 M test points, fitted state taken from
    the GPU fit (W only enters through its diagonal, kernel_sr.py:45)."""
    import torch

    from socks_b200.algorithms.reach.kernel_sr import KernelSR as GpuSR

    ref = ref_loader.load()
    gym = ref["gym"]
    from sklearn.metrics.pairwise import rbf_kernel as sk_rbf

    X, U, Y = syn.integrator_sample(n, d, seed=0)
    ct, tt = syn.integrator_tubes(T, d)
    g = GpuSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", regularization_param=1 / n ** 2)
    g._keep_factors = False
    g.fit_arrays(X, Y)
    w, vf = g._w.cpu().numpy()[:n], g._vf.cpu().numpy()[:, :n]
    box = lambda b: gym.spaces.Box(low=-b, high=b, shape=(d,), dtype=np.float32)
    alg = ref["KernelSR"](time_horizon=T, constraint_tube=[box(1.0)] * T, target_tube=[box(0.5)] * T, problem="FHT",
                          kernel_fn=partial(sk_rbf, gamma=50.0), regularization_param=1 / n ** 2)
    alg._validate_params(None)
    W = np.zeros((n, n))
    W[np.diag_indices(n)] = w
    alg.X, alg.W, alg.value_functions = X, W, vf
    pts = syn.uniform_points(sample, d, seed=1)
    t, out = wall(lambda: alg.predict(pts))
    got = g.predict(pts)
    return {"N": n, "d": d, "sample_points": sample, "reference_predict_s": t, "reference_points_per_s": sample / t,
            "max_abs_diff_gpu_vs_reference": float(np.nanmax(np.abs(got - out)))}


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.endswith(".json")]
    out = [a for a in sys.argv[1:] if a.endswith(".json")]
    res = {"cores": os.cpu_count()}
    for name in args or ["c3", "c4", "c5"]:
        res[name] = {"c3": c3, "c4": c4, "c5": c5}[name]()
        print(name, json.dumps(res[name]), flush=True)
    if out:
        json.dump(res, open(out[0], "w"), indent=1)
