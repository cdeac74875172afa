"""bench.py — KernelSR test points/sec (N=20k samples, T=15, fp64) on 1..8 B200 + Gram/solve seconds.

Workload (BASELINE.json configs[1], SURVEY.md §8(d) "C2"): 2-D stochastic integrator, N = 20 000 samples,
T = 15, FHT, sigma = 0.1, lambda = 1/N^2, M = 250 000 test points per GPU.  A "step" is one pass of the
fused predict path (socks_sr_predict: pack + the DMMA kernel) over the rank's M resident test points; the
fit (Gram + Cholesky + triangular inverse + diag + T-1 GEMV passes) runs once on rank 0, is reported as
`gram_solve_s` / `fit_s`, and its state (X, w, value functions: ~2.9 MB) is broadcast over NCCL.
Test points are independent, so ranks need no data-path collective: weak scaling, M points per rank.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

`--impl reference` times the CPU oracle port of the reference's predict (numpy, all host threads) on a
bounded sample of the same workload and never touches the CUDA library.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SIGMA, PROBLEM = 0.1, "FHT"
DIM = 2  # state dimension of the headline workload; --d overrides it (BASELINE.json configs[4] uses d = 4)
# algorithmic FP64 work per (sample, test point) pair of the fused predict:
# distance 3d+2, shipped exp (exp_f64.cuh exp_scaled_x4: 5 DFMA + DADD + DMUL = 12 flop), contraction 2T
EXP_FLOPS = 12


def pair_flops(d, T):
    return 3 * d + 2 + EXP_FLOPS + 2 * T


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons during the timed region (nvidia-smi based)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            while not self._stop_evt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as e:  # clocks are evidence, not a dependency of the measurement
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_predict_sample(n, T, m_sample, w=None, vf=None, X=None, seed=1):
    """Oracle (numpy port of kernel_sr.py:329-342) predict on m_sample test points at N = n. Returns (pts/s, desc)."""
    from oracle import socks_oracle as orc
    from socks_b200 import synthetic as syn

    ct, tt = syn.integrator_tubes(T, DIM)
    alg = orc.KernelSROracle(T, ct, tt, PROBLEM, SIGMA, 1 / n ** 2, chunk=1000)
    rng = np.random.default_rng(123)
    if X is None:
        X = syn.integrator_sample(n, DIM, seed=0)[0]
    alg.X = X
    alg.w_diag = w if w is not None else 1.0 + rng.random(n)
    alg.value_functions = vf if vf is not None else rng.random((T, n))
    pts = syn.uniform_points(m_sample, DIM, seed=seed)
    alg.predict(pts[:200])  # warm-up
    t0 = time.perf_counter()
    alg.predict(pts)
    dt = time.perf_counter() - t0
    return m_sample / dt, dt


def run_reference(args):
    """CPU arm: the oracle port of the reference's predict path on the host cores (no CUDA)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, T, m_s = args.n, args.T, args.cpu_sample
    cores = os.cpu_count()
    try:  # torchrun exports OMP_NUM_THREADS=1; the reference arm may use every host thread
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=cores)
    except Exception:
        pass
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_predict_sample(n, T, m_s, seed=10 + i)
        if i >= args.warmup:
            vals.append(v)
            secs.append(dt)
    value = m_s * len(secs) / sum(secs)
    sample = (f"{m_s} of the {args.m} test points per step at N={n}, T={T} (numpy port of kernel_sr.py:329-342, "
              f"BLAS threads unrestricted); stand-in fitted state (w, value_functions): per-point cost does not depend "
              f"on their values, and a CPU fit at N={n} alone takes ~130 s (BASELINE.md)")
    line = {
        "impl": "reference", "metric": "KernelSR test points/sec (N=20k, T=15, fp64)", "value": value,
        "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"KernelSR predict, {DIM}-D integrator, N={n}, T={T}, M={args.m}/GPU, sigma=0.1, FHT",
                   "sample_points_per_step": m_s},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from functools import partial

    from socks_b200 import _device as dv, _lib, synthetic as syn
    from socks_b200.algorithms.reach.kernel_sr import KernelSR
    from socks_b200.kernel.metrics import rbf_kernel

    n, m, T, d = args.n, args.m, args.T, DIM
    ct, tt = syn.integrator_tubes(T, d)
    alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem=PROBLEM,
                   kernel_fn=partial(rbf_kernel, sigma=SIGMA), regularization_param=1 / n ** 2)
    alg._validate_params(None)
    X, U, Y = syn.integrator_sample(n, d, seed=0)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    fit_info = {}
    if rank == 0:
        small = syn.integrator_sample(512, d, seed=9)
        KernelSR(time_horizon=3, constraint_tube=ct, target_tube=tt, problem=PROBLEM,
                 regularization_param=1e-3).fit_arrays(small[0], small[2])  # module load / attribute warm-up
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        alg.fit_arrays(X, Y)
        e1.record()
        torch.cuda.synchronize()
        fit_info["fit_s"] = e0.elapsed_time(e1) * 1e-3
        # Gram + solve alone (gram_sym + potrf + trtri + diag), re-timed on the resident inputs
        from socks_b200.kernel.metrics import factorize

        e0, e1 = ev(), ev()
        e0.record()
        fac = factorize(X, None, alg._spec, alg.regularization_param)
        fac.diag()
        e1.record()
        torch.cuda.synchronize()
        fit_info["gram_solve_s"] = e0.elapsed_time(e1) * 1e-3
        npad = fac.np
        fit_info["gram_solve_tflops"] = (2 * npad ** 3 / 3) / fit_info["gram_solve_s"] / 1e12
        del fac
        alg._fac.release_factors()
        torch.cuda.empty_cache()
    else:
        alg.allocate_state(n, d)
    if world > 1:  # fit state broadcast over NCCL/NVLink: the only collective on the path
        from socks_b200.distributed import broadcast_state

        fit_info["broadcast_bytes"] = broadcast_state(alg.state_tensors(), src=0)

    if args.scaling == "strong":  # one global set of M points, contiguous block per rank
        from socks_b200.distributed import shard_bounds

        lo, hi = shard_bounds(m, world, rank)
        pts_host = np.ascontiguousarray(syn.uniform_points(m, d, seed=1)[lo:hi])
        m_total, m = m, hi - lo
    else:
        pts_host = syn.uniform_points(m, d, seed=1 + rank)
        m_total = world * m
    pts = dv.to_dev(pts_host)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step():
        return alg.predict_dev(pts)

    for _ in range(max(args.warmup, 3)):
        out = step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    starts, ends = [ev() for _ in range(args.steps)], [ev() for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()  # evict L2 between timed steps (outside the per-step event pair)
        starts[i].record()
        out = step()
        ends[i].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t_wall0
    dev_s = sum(s.elapsed_time(e) for s, e in zip(starts, ends)) * 1e-3
    clocks = sampler.stop()

    # end to end through the public API: host ndarray in (H2D inside), host ndarray out (D2H inside)
    pinned = torch.from_numpy(pts_host).pin_memory()
    pts_np = pinned.numpy()
    for _ in range(2):
        res = alg.predict(pts_np)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = alg.predict(pts_np)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    tmax = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_s, e2e_s = float(tmax[0]), float(tmax[1])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from probe_fp64 import probe

        fp64_peak = probe(0, iters=8000)["fp64_fma_tflops"]
        dmma_peak = probe(1, iters=8000)["fp64_dmma_tflops"]
        ms_step = 1e3 * dev_s / args.steps
        flops_launch = float(n) * m * pair_flops(d, T)  # rank 0's launch
        achieved = flops_launch / (ms_step * 1e-3) / 1e12
        line = {
            "metric": "KernelSR test points/sec (N=20k, T=15, fp64)",
            "value": m_total * args.steps / dev_s, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"KernelSR predict, {d}-D integrator, N={n}, T={T}, M={args.m}"
                                   f"{'/GPU' if args.scaling == 'weak' else ' in total'}, sigma=0.1, FHT, "
                                   f"lambda=1/N^2 (BASELINE.json configs[1])",
                       "l2": "flushed between timed steps (256 MiB memset outside the per-step event pair)",
                       "parallelism": f"test points sharded, {world} rank(s), fit on rank 0 + NCCL broadcast"},
            "e2e": {"value": m_total * args.steps / e2e_s, "unit": "points/s", "h2d_bytes_per_step": int(pts_host.nbytes),
                    "d2h_bytes_per_step": int(res.nbytes)},
            "gpu_launches": 4 * args.steps,  # per step: factored kernel, direct kernel (exits at once unless the factored form is inadmissible), finish, flagged-column kernel
            "roofline": {"bound": "tensor", "peak_kind": "fp64 DMMA (mma.sync.m8n8k4.f64), measured live",
                         "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s",
                         "frac": achieved / dmma_peak,
                         # dram__bytes_read + dram__bytes_write of one launch at this exact shape, from the ncu
                         # --set full capture summarised in profiles/r1_ncu_sr_predict_dmma_n20000_m250000.txt
                         "traffic": 279.1e6 if (n, m, T, d) == (20000, 250000, 15, 2) else None,
                         "kernel": "sr_predict_dmma_kernel<2>",
                         "note": "algorithmic flops = N*M*(3d+2+E+2T), E=12 (exp_f64.cuh); FMA and DMMA share the "
                                 "FP64 pipe on B200 (tools/probe_fp64.py: 36.5 / 37.1 TFLOP/s alone, 35.7 combined), so the fp64 "
                                 "tensor peak bounds the whole kernel; measured live by socks_probe_fp64 "
                                 "(MEASURED_PEAKS.json has HBM and bf16 only)",
                         "fp64_fma_peak": fp64_peak, "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "clocks": clocks,
            "wall_s_timed_region": wall,
        }
        line.update(fit_info)
        if world == 1 and not args.no_extras and d == 2:
            line["kernel_maximal_sr"] = maximal_sr_extra(n, m, T, d, ct, tt, X, U, Y, pts)
        if world == 1:
            wh, vfh = alg._w.cpu().numpy(), alg._vf.cpu().numpy()
            v, dt = cpu_predict_sample(n, T, args.cpu_sample, w=wh, vf=vfh, X=X)
            line["cpu_baseline"] = {
                "value": v, "unit": "points/s", "cores": os.cpu_count(), "kind": "port",
                "sample": f"oracle predict of {args.cpu_sample} of the {m} test points at N={n}, T={T} on the host "
                          f"cores ({dt:.1f} s), fitted state taken from the GPU fit"}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def maximal_sr_extra(n, m, T, d, ct, tt, X, U, Y, pts, P=100):
    """BASELINE.json configs[1] names KernelMaximalSR at the same sizes: report it beside the headline (one fit and one
    predict over the same M resident test points, P = 100 admissible actions as benchmarks/kernel_sr_max/baseline.py:33)."""
    import torch
    from functools import partial

    from socks_b200 import _device as dv
    from socks_b200.algorithms.reach.kernel_sr_max import KernelMaximalSR
    from socks_b200.kernel.metrics import rbf_kernel

    A = np.linspace(-1, 1, P).reshape(-1, 1)
    alg = KernelMaximalSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem=PROBLEM,
                          kernel_fn=partial(rbf_kernel, sigma=SIGMA), regularization_param=1.0)

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3, r

    t_fit, _ = timed(lambda: alg.fit_arrays(X, U, Y, A))
    alg.predict_dev(pts[:16384])
    t_pred, out = timed(lambda: alg.predict_dev(pts))
    npad = dv.padded(n)
    cols = -(-T * P // 128) * 128
    res = {"P": P, "fit_s": t_fit, "predict_s": t_pred, "points_per_s": m / t_pred,
           "predict_gemm_tflops": 2.0 * npad * m * cols / t_pred / 1e12,
           "note": "predict = Gram chunks + one DMMA GEMM K(X,pts)^T [diag(coef_r) CUA]_r per chunk + fused max/step"}
    del alg, out
    torch.cuda.empty_cache()
    return res


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything except the final JSON line goes to stderr (NCCL and torchrun print banners on fd 1)."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--samples", dest="n", type=int, default=20000)
    ap.add_argument("--m", "--points", dest="m", type=int, default=250000)
    ap.add_argument("--T", "--horizon", dest="T", type=int, default=15)
    ap.add_argument("--cpu-sample", type=int, default=4000)
    ap.add_argument("--no-extras", action="store_true", help="skip the KernelMaximalSR side measurement")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --m test points per GPU (default, the driver's contract); strong: --m in total")
    ap.add_argument("--d", "--dim", dest="d", type=int, default=DIM, help="state dimension (chain of integrators)")
    args = ap.parse_args()
    globals()["DIM"] = args.d
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
