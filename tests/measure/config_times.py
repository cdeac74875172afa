"""Timings of the other BASELINE.json configs through the public API (CUDA events / wall clock), with the
algorithmic work of SURVEY §8(d').  Usage (on a B200): python tests/measure/config_times.py [c1 c2max c3 c4 c5] [out.json]"""
import json
import os
import sys
import time
from functools import partial

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from socks_b200 import _device as dv, synthetic as syn
from socks_b200.algorithms.control.kernel_control_fwd import KernelControlFwd
from socks_b200.algorithms.reach.kernel_sr import KernelSR
from socks_b200.algorithms.reach.kernel_sr_max import KernelMaximalSR
from socks_b200.kernel.metrics import RBFSpec, factorize, rbf_kernel


def wall(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, r


def c1():
    """configs[0]: N=1000, T=15, 50x50 grid, sigma=0.1, lambda=1/N^2 -- whole kernel_sr call, host to host."""
    from oracle import socks_oracle as orc

    n, T = 1000, 15
    X, U, Y = syn.integrator_sample(n, seed=0)
    pts = syn.grid_points(50, 2)
    ct, tt = syn.integrator_tubes(T)
    res = {}
    for problem in ("FHT", "THT"):
        def run():
            a = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem=problem,
                         kernel_fn=partial(rbf_kernel, sigma=0.1), regularization_param=1 / n ** 2)
            a.fit_arrays(X, Y)
            return a.predict(pts)
        run()
        ts = [wall(run)[0] for _ in range(5)]
        t0 = time.perf_counter()
        o = orc.KernelSROracle(T, ct, tt, problem, 0.1, 1 / n ** 2).fit(X, U, Y).predict(pts)
        t_cpu = time.perf_counter() - t0
        res[problem] = {"gpu_s_median": float(np.median(ts)), "cpu_oracle_s": t_cpu,
                        "max_abs_diff": float(np.max(np.abs(run() - o)))}
    return res


def c2max(n=20000, m=250000, T=15, P=100):
    X, U, Y = syn.integrator_sample(n, seed=0)
    pts = syn.uniform_points(m, seed=1)
    A = np.linspace(-1, 1, P).reshape(-1, 1)
    ct, tt = syn.integrator_tubes(T)
    alg = KernelMaximalSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT",
                          kernel_fn=partial(rbf_kernel, sigma=0.1), regularization_param=1.0)
    t_fit, _ = wall(lambda: alg.fit_arrays(X, U, Y, A))
    t_fit2, _ = wall(lambda: alg.fit_arrays(X, U, Y, A))
    pd = dv.to_dev(pts)
    alg.predict_dev(pd[:16384])
    t_pred, out = wall(lambda: alg.predict_dev(pd))
    npad = dv.padded(n)
    return {"N": n, "M": m, "T": T, "P": P, "fit_s_first": t_fit, "fit_s": t_fit2, "predict_s": t_pred,
            "predict_pts_per_s": m / t_pred,
            "predict_gemm_tflops": 2.0 * npad * m * (-(-T * P // 128) * 128) / t_pred / 1e12,
            "fit_step_gemm_flops": 2.0 * npad * npad * 256 * (T - 1),
            "vf_range": [float(alg._vf.min()), float(alg._vf.max())],
            "out_range": [float(out.min()), float(out.max())]}


def c3(n=5000, P_grid=(25, 40), T=20):
    X, U, Y = syn.unicycle_sample(n, seed=12345)
    A = syn.unicycle_actions(*P_grid)
    pol = KernelControlFwd(cost_fn=syn.tracking_cost_fn(T), heuristic=True, regularization_param=1e-7,
                           kernel_fn=partial(rbf_kernel, sigma=3.0), verbose=False)
    t_train, _ = wall(lambda: pol.train_arrays(X, U, Y, A))
    state = np.array([[-0.8, 0.0, 0.0]])
    pol(time=0, state=state)
    ts = []
    for t in range(T):
        dt, act = wall(lambda: pol(time=t, state=state))
        ts.append(dt)
        state = state + 0.05  # any state: the cost of a call does not depend on it
    # where a call's time goes: the user's cost_fn on the host (part of the reference API, kernel_control_fwd.py:225) and
    # the device work of socks_fwd_policy (CUDA events around the enqueue, inputs already staged)
    tc = []
    for t in range(T):
        t0 = time.perf_counter()
        np.array(pol.cost_fn(time=t), dtype=np.float32)
        tc.append(time.perf_counter() - t0)
    td = []
    for t in range(T):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        pol._enqueue(t, state, policy=True)
        e1.record()
        torch.cuda.synchronize()
        td.append(e0.elapsed_time(e1) * 1e-3)
    return {"N": n, "P": len(A), "T": T, "train_s": t_train, "call_s_median": float(np.median(ts)),
            "call_s_max": float(np.max(ts)), "host_cost_fn_s_median": float(np.median(tc)),
            "enqueue_to_done_s_median": float(np.median(td)),
            "call_minus_cost_fn_s": float(np.median(ts) - np.median(tc)),
            "hbm_floor_s": 8.0 * n * n / 6546.2e9, "reference_call_flops": 2.0 * n * n * len(A)}


def c4(n=40000):
    X, U, Y = syn.integrator_sample(n, seed=0)
    spec = RBFSpec(sigma=0.1)
    def run():
        fac = factorize(X, None, spec, 1 / n ** 2)
        W = fac.full()
        fac.check()
        return fac
    t, fac = wall(run)
    npad = fac.np
    Wd = fac.full()
    sym = float((Wd[:2048, :2048] - Wd[:2048, :2048].T).abs().max())
    # residual check on a block of columns: (G W)[:, :64] = I[:, :64], G rebuilt from the kernel
    Xd = dv.to_dev(X)
    from socks_b200.kernel.metrics import gram_dev
    G = gram_dev(Xd, Xd[:64], spec.params_for(Xd, Xd))[:, :64]  # G[:, :64] without the ridge
    G[torch.arange(64), torch.arange(64)] += 1.0 / n
    R = Wd[:n, :n] @ G  # torch matmul only as a checker
    R[torch.arange(64), torch.arange(64)] -= 1.0
    return {"N": n, "gram_potrf_trtri_lauum_s": t, "tflops": float(npad) ** 3 / t / 1e12,
            "sym_err": sym, "residual_max": float(R.abs().max()), "bytes_W": 8.0 * npad * npad}


def c5(n=40000, m=4000000, T=15, d=4):
    X, U, Y = syn.integrator_sample(n, dim=d, seed=0)
    ct, tt = syn.integrator_tubes(T, d)
    alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT",
                   kernel_fn=partial(rbf_kernel, sigma=0.1), regularization_param=1 / n ** 2)
    t_fit, _ = wall(lambda: alg.fit_arrays(X, Y))
    alg._fac.release_factors()
    torch.cuda.empty_cache()
    pts = dv.to_dev(syn.uniform_points(m, d, seed=1))
    alg.predict_dev(pts[:65536])
    t_pred, out = wall(lambda: alg.predict_dev(pts))
    nan = int(torch.isnan(out).sum())
    return {"N": n, "M": m, "T": T, "d": d, "fit_s": t_fit, "predict_s": t_pred, "pts_per_s": m / t_pred,
            "fp64_tflops": float(n) * m * (3 * d + 2 + 18 + 2 * T) / t_pred / 1e12, "nan_outputs": nan,
            "out_range": [float(torch.nan_to_num(out).min()), float(torch.nan_to_num(out).max())]}


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.endswith(".json")] or ["c1", "c2max", "c3", "c4", "c5"]
    outp = [a for a in sys.argv[1:] if a.endswith(".json")]
    res = {}
    for nm in names:
        try:
            res[nm] = globals()[nm]()
        except Exception as e:  # keep going: each config is independent
            res[nm] = {"error": f"{type(e).__name__}: {e}"}
        print(nm, json.dumps(res[nm]), flush=True)
        torch.cuda.empty_cache()
    if outp:
        json.dump(res, open(outp[0], "w"), indent=1)
