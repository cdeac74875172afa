"""Sacred-free re-run of the reference's published sweeps: kernel_sr and kernel_sr_max
wall time vs sample size, test-set size, time horizon, dimensionality and number of actions, with the reference's
defaults (d=2, N=1000, M=1000, T=10, sigma=0.1, lambda=1, THT, sklearn rbf_kernel(gamma=1/(2 sigma^2));
`benchmarks/kernel_sr/baseline.py:26-34,127-146`, `benchmarks/kernel_sr_max/baseline.py:33`, sweeps
`benchmarks/*/plot_*_vs_time.py:19`).  Timed region as in the reference: tube construction + fit + predict; host arrays
in, host array out.  Repeats follow the reference's own rule (`benchmarks/experiment_matrix.py:44-62,224-255`): one
discarded warm-up, then at least 3 and at most 32 runs, stopping early as soon as `sample_mean` reports the last run
OUTSIDE the 95 % normal interval of the runs so far (that is the rule as the reference codes it: `if callback_result is
False or len(results) >= max_iter: return`); the mean, the median and the number of runs are reported.  The CPU column is
the oracle port (run once).

Usage (on a B200): python tests/measure/sweeps.py [out.json] [--no-cpu]
M-sweep on 1..8 GPUs (test points sharded, fit on rank 0, NCCL broadcast + gather):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/measure/sweeps.py --msweep [out.json]"""
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


def sample_mean(results, alpha=0.95):
    """The reference's callback (benchmarks/experiment_matrix.py:224-255): True if the last result lies inside the
    `alpha` normal interval around the mean of all results so far (sample standard deviation)."""
    from scipy import stats

    results = np.asarray(results)
    last, mu, sigma = results[-1], results.mean(), results.std(ddof=1)
    lo, hi = stats.norm.interval(alpha, loc=mu, scale=sigma)
    return bool(last >= lo and last < hi)


def reference_repeats(run_once, min_iter=3, max_iter=32, warmup=True, callback=sample_mean):
    """`_result_accumulator` of benchmarks/experiment_matrix.py:17-62, with run_once() -> seconds."""
    if warmup:
        run_once()
    results = []
    while True:
        results.append(run_once())
        if len(results) >= min_iter:
            cb = callback(results) if callback is not None else False
            if cb is False or len(results) >= max_iter:
                return results


def timed(fn):
    def once():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    ts = reference_repeats(once)
    return {"mean": float(np.mean(ts)), "median": float(np.median(ts)), "runs": len(ts)}


def run_sr(n, m, T, d, cpu):
    X, U, Y = syn.integrator_sample(n, dim=d, seed=0)
    pts = syn.uniform_points(m, d, seed=1)
    S = syn.as_sample(X, U, Y)

    def gpu():
        ct, tt = syn.integrator_tubes(T, d)
        return kernel_sr(S, pts, time_horizon=T, constraint_tube=ct, target_tube=tt, problem="THT",
                         regularization_param=LAM, kernel_fn=partial(sk_rbf, gamma=1 / (2 * SIGMA ** 2)))

    t = timed(gpu)
    out = {"gpu_s": t["median"], "gpu_s_mean": t["mean"], "runs": t["runs"]}
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

    t = timed(gpu)
    out = {"gpu_s": t["median"], "gpu_s_mean": t["mean"], "runs": t["runs"]}
    if cpu:
        def cpu_fn():
            ct, tt = syn.integrator_tubes(T, d)
            return orc.KernelMaximalSROracle(T, ct, tt, "THT", SIGMA, LAM).fit(X, U, Y, A).predict(pts)
        t0 = time.perf_counter()
        ref = cpu_fn()
        out["cpu_s"] = time.perf_counter() - t0
        out["max_abs_diff"] = float(np.nanmax(np.abs(gpu() - ref)))
    return out


def msweep():
    """Wall time of tubes + fit + predict vs test-set size on WORLD_SIZE GPUs (the reference's M-sweep,
    benchmarks/kernel_sr/plot_test_sample_size_vs_time.py:19, extended past its 1 500 points): rank 0 fits, the fitted
    state is broadcast, every rank predicts its contiguous block, the blocks are gathered on the device."""
    import torch.distributed as dist

    from socks_b200 import _device as dv
    from socks_b200.algorithms.reach.kernel_sr import KernelSR
    from socks_b200.distributed import broadcast_state, gather_columns, init_from_env, shard_bounds

    rank, world, local = init_from_env()
    outp = [a for a in sys.argv[1:] if a.endswith(".json")]
    res = {"gpus": world, "repeat_rule": "experiment_matrix.py:44-62 (warm-up + 3..32 runs, sample_mean callback)", "rows": []}
    for n in (1000, 20000):
        X, U, Y = syn.integrator_sample(n, dim=2, seed=0)
        for m in (1000, 10000, 100000, 1000000):
            pts_h = syn.uniform_points(m, 2, seed=1)
            lo, hi = shard_bounds(m, world, rank)
            T = 10

            def run():
                ct, tt = syn.integrator_tubes(T, 2)
                alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="THT", regularization_param=LAM,
                               kernel_fn=partial(sk_rbf, gamma=1 / (2 * SIGMA ** 2)))
                alg._validate_params(None)
                alg._keep_factors = False
                if rank == 0:
                    alg.fit_arrays(X, Y)
                else:
                    alg.allocate_state(n, 2)
                broadcast_state(alg.state_tensors(), src=0)
                local_out = alg.predict_dev(dv.to_dev(pts_h[lo:hi]))
                return gather_columns(local_out, m) if world > 1 else local_out

            def once():
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                run()
                torch.cuda.synchronize()
                dt = torch.tensor([time.perf_counter() - t0], device="cuda")
                if world > 1:
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX)  # every rank sees the same time => the same stopping point
                return float(dt)

            ts = reference_repeats(once)
            row = {"N": n, "M": m, "T": T, "s_mean": float(np.mean(ts)), "s_median": float(np.median(ts)), "runs": len(ts)}
            if rank == 0:
                print(json.dumps(row), flush=True)
            res["rows"].append(row)
    if rank == 0 and outp:
        json.dump(res, open(outp[0], "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


def main():
    if "--msweep" in sys.argv:
        return msweep()
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
