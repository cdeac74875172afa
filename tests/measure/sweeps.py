"""Sacred-free re-run of the reference's published sweeps: kernel_sr and kernel_sr_max
wall time vs sample size, test-set size, time horizon, dimensionality and number of actions, with the reference's
defaults (d=2, N=1000, M=1000, T=10, sigma=0.1, lambda=1, THT, sklearn rbf_kernel(gamma=1/(2 sigma^2));
`benchmarks/kernel_sr/baseline.py:26-34,127-146`, `benchmarks/kernel_sr_max/baseline.py:33`, sweeps
`benchmarks/*/plot_*_vs_time.py:19`).  Timed region as in the reference: tube construction + fit + predict, one
warm-up discarded, median of 5 repeats; host arrays in, host array out.  The CPU column is the oracle port.

Usage (on a B200): python tests/measure/sweeps.py [out.json] [--no-cpu]"""
import json
import os
import sys
import time
from functools import partial

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from sklearn.metrics.pairwise import rbf_kernel as sk_rbf

from oracle import socks_oracle as orc
from socks_b200 import synthetic as syn
from socks_b200.algorithms.reach.kernel_sr import kernel_sr
from socks_b200.algorithms.reach.kernel_sr_max import kernel_sr_max

SIGMA, LAM = 0.1, 1.0
DEF = dict(d=2, n=1000, m=1000, T=10, P=100)


def timed(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def run_sr(n, m, T, d, cpu):
    X, U, Y = syn.integrator_sample(n, dim=d, seed=0)
    pts = syn.uniform_points(m, d, seed=1)
    S = syn.as_sample(X, U, Y)

    def gpu():
        ct, tt = syn.integrator_tubes(T, d)
        return kernel_sr(S, pts, time_horizon=T, constraint_tube=ct, target_tube=tt, problem="THT",
                         regularization_param=LAM, kernel_fn=partial(sk_rbf, gamma=1 / (2 * SIGMA ** 2)))

    out = {"gpu_s": timed(gpu)}
    if cpu:
        def cpu_fn():
            ct, tt = syn.integrator_tubes(T, d)
            return orc.KernelSROracle(T, ct, tt, "THT", SIGMA, LAM).fit(X, U, Y).predict(pts)
        t0 = time.perf_counter()
        ref = cpu_fn()
        out["cpu_s"] = time.perf_counter() - t0
        out["max_abs_diff"] = float(np.nanmax(np.abs(gpu() - ref)))
    return out


def run_srmax(n, m, T, P, cpu):
    d = 2
    X, U, Y = syn.integrator_sample(n, dim=d, seed=0)
    pts = syn.uniform_points(m, d, seed=1)
    A = np.linspace(-1, 1, P).reshape(-1, 1)
    S = syn.as_sample(X, U, Y)

    def gpu():
        ct, tt = syn.integrator_tubes(T, d)
        return kernel_sr_max(S, A, pts, time_horizon=T, constraint_tube=ct, target_tube=tt, problem="THT",
                             regularization_param=LAM, kernel_fn=partial(sk_rbf, gamma=1 / (2 * SIGMA ** 2)))

    out = {"gpu_s": timed(gpu)}
    if cpu:
        def cpu_fn():
            ct, tt = syn.integrator_tubes(T, d)
            return orc.KernelMaximalSROracle(T, ct, tt, "THT", SIGMA, LAM).fit(X, U, Y, A).predict(pts)
        t0 = time.perf_counter()
        ref = cpu_fn()
        out["cpu_s"] = time.perf_counter() - t0
        out["max_abs_diff"] = float(np.nanmax(np.abs(gpu() - ref)))
    return out


def main():
    cpu = "--no-cpu" not in sys.argv
    outp = [a for a in sys.argv[1:] if a.endswith(".json")]
    res = {"defaults": DEF, "sigma": SIGMA, "lambda": LAM, "problem": "THT", "kernel_sr": {}, "kernel_sr_max": {}}
    D = DEF
    sr = res["kernel_sr"]
    sr["sample_size"] = {n: run_sr(n, D["m"], D["T"], D["d"], cpu) for n in (100, 500, 1000, 1500)}
    sr["test_sample_size"] = {m: run_sr(D["n"], m, D["T"], D["d"], cpu) for m in (100, 500, 1000, 1500)}
    sr["time_horizon"] = {T: run_sr(D["n"], D["m"], T, D["d"], cpu) for T in (1, 5, 10, 20)}
    sr["dimensionality"] = {d: run_sr(D["n"], D["m"], D["T"], d, cpu) for d in (2, 4, 6, 8)}
    mx = res["kernel_sr_max"]
    mx["sample_size"] = {n: run_srmax(n, D["m"], D["T"], D["P"], cpu) for n in (100, 500, 1000, 1500)}
    mx["action_sample_size"] = {P: run_srmax(D["n"], D["m"], D["T"], P, cpu) for P in (25, 100, 250)}
    mx["test_sample_size"] = {m: run_srmax(D["n"], m, D["T"], D["P"], cpu) for m in (100, 1000, 1500)}
    mx["time_horizon"] = {T: run_srmax(D["n"], D["m"], T, D["P"], cpu) for T in (1, 10, 20)}
    res["published_by_reference_plots_s"] = {
        "kernel_sr": {"sample_size": {"100": 0.01, "1000": 0.10, "1500": 0.23}, "test_sample_size": {"100": 0.081, "1500": 0.113},
                      "time_horizon": {"1": 0.090, "20": 0.114}, "dimensionality": {"2": 0.10, "8": 0.114}},
        "kernel_sr_max": {"sample_size": {"100": 0.1, "1000": 2.0, "1500": 3.8}, "action_sample_size": {"25": 0.63, "250": 5.05},
                          "test_sample_size": {"100": 1.15, "1500": 2.5}, "time_horizon": {"1": 1.32, "20": 2.85}},
        "source": "BASELINE.md section 1 (read off docs/benchmarks/*.png, hardware unstated)"}
    print(json.dumps(res))
    if outp:
        json.dump(res, open(outp[0], "w"), indent=1)


if __name__ == "__main__":
    main()
