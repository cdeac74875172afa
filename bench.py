"""bench.py — KernelSR test points/sec (N=20k samples, T=15, fp64) on 1..8 B200 + Gram/solve seconds.

Workload (BASELINE.json configs[1], SURVEY.md §8(d) "C2"): 2-D stochastic integrator, N = 20 000 samples, T = 15, FHT,
sigma = 0.1, lambda = 1/N^2, M = 250 000 test points.  A "step" is one pass of the fused predict path
(socks_sr_predict_packed: the tensor-core kernel + finish) over the rank's resident test points.  The fit (Gram +
Cholesky/triangular inverse + diag + T-1 GEMV passes + pack) runs once on rank 0, is reported stage by stage
(`gram_solve_s`, `potrf_inv_tflops`, `fit_recursion_gbs`, ...), and its state (X, w, value functions: ~2.9 MB) is
broadcast over NCCL.  Test points are independent columns, so ranks need no data-path collective.

Scaling: with --gpus N > 1 the DEFAULT is strong scaling -- the 250 000 points of configs[1] are split over the ranks
(north_star: ">= 7x scaling of test-point throughput from 1 to 8 B200") -- and the weak-scaling number (250 000 points
per rank) is measured in the same run and reported under "weak".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling strong|weak]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

`--impl reference` times the UNMODIFIED reference (`gym_socks.algorithms.reach.kernel_sr.KernelSR`, imported from
/root/reference or from the verbatim archive oracle/vendor_ref.py leaves in oracle/_ref) on the host cores: one reference
fit at full N (untimed, reported), then `predict` on a bounded sample of the same test points per step.  It never
touches the CUDA library.  If the reference cannot be imported, the numpy oracle port is timed instead
(`cpu_baseline.kind: "port"`).
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
METRIC = "KernelSR test points/sec (N=20k, T=15, fp64)"
# algorithmic FP64 work per (sample, test point) pair of the fused predict: distance 3d+2, exp E = 12
# (the figure of the round-1 table exp, kept so that fractions stay comparable across rounds; the shipped factored
# exponent issues fewer instructions than that), contraction 2T
EXP_FLOPS = 12


def pair_flops(d, T):
    return 3 * d + 2 + EXP_FLOPS + 2 * T


def workload_name(n, m, T, d):
    return (f"KernelSR predict, {d}-D integrator, N={n}, T={T}, M={m}, sigma={SIGMA}, {PROBLEM}, lambda=1/N^2 "
            f"(BASELINE.json configs[1])")


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons during the timed region (NVML)."""

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
                time.sleep(0.01)
        except Exception as e:  # clocks are evidence, not a dependency of the measurement
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------- CPU side (oracle/)
def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm may use every host thread."""
    cores = os.cpu_count()
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=cores)
    except Exception:
        pass
    return cores


def _blas_name():
    try:
        from threadpoolctl import threadpool_info

        return ",".join(sorted({f"{i.get('internal_api')}:{i.get('num_threads')}" for i in threadpool_info()}))
    except Exception:
        return "unknown"


def reference_predictor(n, T, d, X=None, w=None, vf=None):
    """-> (predict(pts) callable, info dict).  The unmodified reference's KernelSR with the kernel_fn its own benchmark
    uses (sklearn rbf_kernel, gamma = 1/(2 sigma^2): benchmarks/kernel_sr/baseline.py:20,134-144).  With (X, w, vf) given
    the fitted state is injected (W only enters through its diagonal, kernel_sr.py:45, so a zero matrix carrying diag(W)
    reproduces predict exactly); otherwise the reference fits itself on the synthetic sample (minutes at N = 20 000)."""
    from functools import partial

    from oracle import ref_loader
    from sklearn.metrics.pairwise import rbf_kernel as sk_rbf
    from socks_b200 import synthetic as syn

    ref = ref_loader.load()
    gym = ref["gym"]
    box = lambda b: gym.spaces.Box(low=-b, high=b, shape=(d,), dtype=np.float32)
    alg = ref["KernelSR"](time_horizon=T, constraint_tube=[box(1.0)] * T, target_tube=[box(0.5)] * T, problem=PROBLEM,
                          kernel_fn=partial(sk_rbf, gamma=1 / (2 * SIGMA ** 2)), regularization_param=1 / n ** 2)
    info = {"root": ref["root"], "kernel_fn": "sklearn.metrics.pairwise.rbf_kernel(gamma=1/(2 sigma^2))"}
    if X is None:
        Xs, Us, Ys = syn.integrator_sample(n, d, seed=0)
        t0 = time.perf_counter()
        alg.fit(list(zip(Xs, Us, Ys)))
        info["fit_s"] = time.perf_counter() - t0
        info["state"] = "fitted by the reference itself"
    else:
        alg._validate_params(None)
        W = np.zeros((n, n))  # calloc: only the diagonal's pages are ever touched
        W[np.diag_indices(n)] = w
        alg.X, alg.W, alg.value_functions = np.asarray(X), W, np.asarray(vf)
        info["state"] = "X, diag(W), value functions taken from the GPU fit"
    return alg.predict, info


def port_predictor(n, T, d, X=None, w=None, vf=None):
    """Oracle (numpy port of kernel_sr.py:329-342) predict; stand-in state if none is given."""
    from oracle import socks_oracle as orc
    from socks_b200 import synthetic as syn

    ct, tt = syn.integrator_tubes(T, d)
    alg = orc.KernelSROracle(T, ct, tt, PROBLEM, SIGMA, 1 / n ** 2, chunk=1000)
    rng = np.random.default_rng(123)
    alg.X = X if X is not None else syn.integrator_sample(n, d, seed=0)[0]
    alg.w_diag = w if w is not None else 1.0 + rng.random(n)
    alg.value_functions = vf if vf is not None else rng.random((T, n))
    return alg.predict, {"state": "from the GPU fit" if w is not None else "stand-in (w, value_functions)"}


def time_cpu(predict, pts_list):
    """Seconds per call of predict over the given point batches (first batch is an untimed warm-up)."""
    predict(pts_list[0][:200])
    secs = []
    for p in pts_list:
        t0 = time.perf_counter()
        predict(p)
        secs.append(time.perf_counter() - t0)
    return secs


def run_reference(args):
    """CPU arm: the unmodified reference's predict path on the host cores (no CUDA)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from socks_b200 import synthetic as syn

    n, T, m_s, d = args.n, args.T, args.cpu_sample, DIM
    cores = _use_all_host_threads()
    kind, info = "reference", {}
    try:
        if args.ref_state == "port":
            raise ImportError("--ref-state port")
        predict, info = reference_predictor(n, T, d)
    except Exception as e:  # reference not importable on this box: the numpy port is the CPU arm
        kind = "port"
        predict, info = port_predictor(n, T, d)
        info["why_port"] = f"{type(e).__name__}: {e}"[:200]
    allp = syn.uniform_points(args.m, d, seed=1)
    k = args.warmup + args.steps
    batches = [allp[(i * m_s) % max(args.m - m_s, 1):][:m_s] for i in range(k)]
    secs = time_cpu(predict, batches)[args.warmup:]
    value = m_s * len(secs) / sum(secs)
    sample = (f"{m_s} of the {args.m} test points per step at N={n}, T={T}: unmodified gym_socks KernelSR.predict "
              f"(kernel_sr.py:311-344), {info.get('kernel_fn', 'numpy port')}, state {info.get('state')}; BLAS {_blas_name()}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, args.m, T, d)},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "reference_fit_s": info.get("fit_s"), "reference_root": info.get("root"),
    }
    if "why_port" in info:
        line["reference_unavailable"] = info["why_port"]
    _emit(line)


# ------------------------------------------------------------------------------------------------- GPU side
def _ncu_traffic(kernel_prefix, shape):
    """dram bytes per launch from the committed ncu capture of this exact kernel and shape (tools/ncu_summary.py
    writes profiles/*.json next to the text summary); None when there is no capture for the shape."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for name in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if not (name.startswith("r2_ncu_") and name.endswith(".json")):
            continue
        try:
            rec = json.load(open(os.path.join(pdir, name)))
        except Exception:
            continue
        if rec.get("kernel", "").startswith(kernel_prefix) and rec.get("shape") == shape:
            best = {"bytes": rec.get("dram_bytes_read", 0) + rec.get("dram_bytes_write", 0), "source": "profiles/" + name}
    return best


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

    from socks_b200 import _device as dv, synthetic as syn
    from socks_b200.algorithms.reach.kernel_sr import KernelSR
    from socks_b200.distributed import broadcast_state, shard_bounds
    from socks_b200.kernel.metrics import rbf_kernel

    n, m, T, d = args.n, args.m, args.T, DIM
    ct, tt = syn.integrator_tubes(T, d)
    alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem=PROBLEM,
                   kernel_fn=partial(rbf_kernel, sigma=SIGMA), regularization_param=1 / n ** 2)
    alg._validate_params(None)
    alg._keep_factors = False
    X, U, Y = syn.integrator_sample(n, d, seed=0)

    fit_info = {}
    if rank == 0:
        small = syn.integrator_sample(512, d, seed=9)
        KernelSR(time_horizon=3, constraint_tube=ct, target_tube=tt, problem=PROBLEM,
                 regularization_param=1e-3).fit_arrays(small[0], small[2])  # module load / attribute warm-up
        alg.fit_arrays(X, Y)  # one discarded warm-up at full size, as the reference's harness does
        torch.cuda.synchronize()  # (benchmarks/experiment_matrix.py:44-46): first-touch cudaMalloc of the 3.2 GB buffers
        dv.stage_marks = []
        alg.fit_arrays(X, Y)
        torch.cuda.synchronize()
        st = dv.stage_times()
        dv.stage_marks = None
        npad = dv.padded(n)
        fit_info = {
            "fit_s": sum(st.values()),
            "gram_solve_s": st["gram_sym"] + st["potrf_inv"] + st["diag_inv"],
            "gram_solve_tflops": (2 * npad ** 3 / 3) / (st["gram_sym"] + st["potrf_inv"] + st["diag_inv"]) / 1e12,
            "potrf_inv_s": st["potrf_inv"], "potrf_inv_tflops": (2 * npad ** 3 / 3) / st["potrf_inv"] / 1e12,
            "potrf_inv_engine": "look-ahead Cholesky + inverse; large GEMMs as fp64-accurate digit GEMMs on tcgen05.mma "
                                "kind::i8 (7 digits in the trailing updates, 8 in the merges), so *_tflops are "
                                "fp64-EQUIVALENT rates (fp64 DMMA peak: roofline.peak)",
            "gram_sym_s": st["gram_sym"], "gram_cross_s": st["gram_cross"],
            "fit_recursion_s": st["fit_recursion"],
            # HBM-bound stage: (T-1) GEMV passes over the resident N x N beta, bytes = (T-1) * 8 * N^2 (the
            # normaliser pass and the step kernels are inside the time but not in the bytes)
            "fit_recursion_gbs": (T - 1) * 8.0 * n * n / st["fit_recursion"] / 1e9,
            "stages_s": st,
        }
    else:
        alg.allocate_state(n, d)
    if world > 1:  # fit state broadcast over NCCL/NVLink: the only collective on the path
        try:  # the library's own communicator (socks_comm_init / socks_bcast); torch.distributed is the rendezvous
            from socks_b200.distributed import SocksComm

            comm = SocksComm()
            fit_info["broadcast_bytes"] = comm.broadcast_state(alg.state_tensors(), src=0)
            fit_info["broadcast_via"] = "socks_bcast (libsocksb200 -> NCCL)"
            comm.close()
        except Exception as e:  # plumbing only: same bytes through torch.distributed
            fit_info["broadcast_bytes"] = broadcast_state(alg.state_tensors(), src=0)
            fit_info["broadcast_via"] = f"torch.distributed ({type(e).__name__}: {e})"[:160]

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if "fit_recursion_gbs" in fit_info and peaks.get("hbm_gbs"):
        fit_info["fit_recursion_frac_hbm"] = fit_info["fit_recursion_gbs"] / peaks["hbm_gbs"]

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def measure(mode):
        """Device-timed and end-to-end throughput of one sharding mode; returns a dict (identical on all ranks)."""
        if mode == "strong":  # one global set of M points, contiguous block per rank
            lo, hi = shard_bounds(m, world, rank)
            pts_host = np.ascontiguousarray(syn.uniform_points(m, d, seed=1)[lo:hi])
            m_total = m
        else:
            pts_host = syn.uniform_points(m, d, seed=1 + rank)
            m_total = world * m
        pts = dv.to_dev(pts_host)
        for _ in range(max(args.warmup, 3)):
            out = alg.predict_dev(pts)
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
            out = alg.predict_dev(pts)
            ends[i].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - t_wall0
        dev_s = sum(s.elapsed_time(e) for s, e in zip(starts, ends)) * 1e-3
        clocks = sampler.stop()

        # end to end through the public API: host ndarray in (H2D inside), host ndarray out (D2H inside)
        def e2e(arr):
            for _ in range(2):
                res = alg.predict(arr)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                res = alg.predict(arr)
            torch.cuda.synchronize()
            return time.perf_counter() - t0, res

        pinned = torch.from_numpy(pts_host).pin_memory()
        e2e_s, res = e2e(pinned.numpy())
        e2e_pageable_s, _ = e2e(pts_host)  # what a drop-in caller passes: a plain (pageable) ndarray
        tmax = torch.tensor([dev_s, e2e_s, e2e_pageable_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dev_s, e2e_s, e2e_pageable_s = (float(x) for x in tmax)
        del out
        return {"m_rank": int(pts_host.shape[0]), "m_total": m_total, "dev_s": dev_s, "e2e_s": e2e_s,
                "e2e_pageable_s": e2e_pageable_s, "wall": wall, "clocks": clocks, "h2d": int(pts_host.nbytes),
                "d2h": int(res.nbytes), "pts": pts}

    main_mode = args.scaling
    r = measure(main_mode)
    other = None
    if world > 1 and not args.no_extras:
        other = measure("weak" if main_mode == "strong" else "strong")

    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from probe_fp64 import probe

        fp64_peak = probe(0, iters=8000)["fp64_fma_tflops"]
        dmma_peak = probe(1, iters=8000)["fp64_dmma_tflops"]
        ms_step = 1e3 * r["dev_s"] / args.steps
        flops_launch = float(n) * r["m_rank"] * pair_flops(d, T)  # rank 0's launch
        achieved = flops_launch / (ms_step * 1e-3) / 1e12
        traffic = _ncu_traffic("sr_predict_fast_kernel", [n, r["m_rank"], T, d])
        line = {
            "metric": METRIC, "value": r["m_total"] * args.steps / r["dev_s"], "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": main_mode, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, m, T, d),
                       "points": f"{r['m_total']} in total, {r['m_rank']} on rank 0",
                       "l2": "flushed between timed steps (256 MiB memset outside the per-step event pair)",
                       "parallelism": f"test points sharded over {world} rank(s), fit on rank 0 + NCCL broadcast"},
            "e2e": {"value": r["m_total"] * args.steps / r["e2e_s"], "unit": "points/s",
                    "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                    "input": "pinned host ndarray (contract); pageable_input_value = plain ndarray, as a drop-in caller passes",
                    "pageable_input_value": r["m_total"] * args.steps / r["e2e_pageable_s"]},
            "gpu_launches": 4 * args.steps,  # per step: factored kernel, direct-form kernel (exits at once unless
            # the factored form is inadmissible), finish, flagged-column kernel
            "roofline": {"bound": "tensor", "peak_kind": "fp64 DMMA (mma.sync.m8n8k4.f64), measured live",
                         "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s", "frac": achieved / dmma_peak,
                         "traffic": traffic["bytes"] if traffic else None,
                         "traffic_source": traffic["source"] if traffic else None,
                         "kernel": f"sr_predict_fast_kernel<{d}, ...> (csrc/sr_predict.cu)",
                         "note": "algorithmic flops = N*M*(3d+2+E+2T), E=12; FMA and DMMA share the FP64 pipe on B200 "
                                 "(tools/probe_fp64.py: 36.5 / 37.1 TFLOP/s alone, 35.7 combined), so the fp64 tensor "
                                 "peak bounds the whole kernel; measured live by socks_probe_fp64 (MEASURED_PEAKS.json "
                                 "has HBM and bf16 only)",
                         "fp64_fma_peak": fp64_peak, "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "clocks": r["clocks"],
            "wall_s_timed_region": r["wall"],
        }
        if other is not None:
            line["weak" if main_mode == "strong" else "strong"] = {
                "value": other["m_total"] * args.steps / other["dev_s"], "unit": "points/s",
                "ms_per_step": 1e3 * other["dev_s"] / args.steps, "points_total": other["m_total"],
                "e2e_value": other["m_total"] * args.steps / other["e2e_s"], "clocks": other["clocks"]}
        line.update(fit_info)
        if world == 1 and not args.no_extras and d == 2:
            line["kernel_maximal_sr"] = maximal_sr_extra(n, m, T, d, ct, tt, X, U, Y, r["pts"])
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(alg, n, T, d, m, args.cpu_sample, X)
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(alg, n, T, d, m, m_s, X):
    """The reference's predict on the host cores, on a bounded sample of the same workload, with the fitted state of the
    GPU fit injected; the numpy port beside it."""
    from socks_b200 import synthetic as syn

    cores = _use_all_host_threads()
    wh, vfh = alg._w.cpu().numpy(), alg._vf.cpu().numpy()
    pts = syn.uniform_points(m, d, seed=1)[:m_s]
    out = {"unit": "points/s", "cores": cores}
    secs_p = time_cpu(port_predictor(n, T, d, X, wh, vfh)[0], [pts])
    out["port_value"] = m_s / secs_p[0]
    try:
        predict, info = reference_predictor(n, T, d, X, wh, vfh)
        secs = time_cpu(predict, [pts])
        out.update(value=m_s / secs[0], kind="reference",
                   sample=f"unmodified gym_socks KernelSR.predict (kernel_sr.py:311-344, {info['kernel_fn']}) on {m_s} "
                          f"of the {m} test points at N={n}, T={T} ({secs[0]:.1f} s on the host cores; numpy port: "
                          f"{secs_p[0]:.1f} s); {info['state']}; BLAS {_blas_name()}")
    except Exception as e:
        out.update(value=out["port_value"], kind="port",
                   sample=f"oracle port predict of {m_s} of the {m} test points at N={n}, T={T} ({secs_p[0]:.1f} s); "
                          f"reference not importable: {type(e).__name__}: {e}"[:300])
    return out


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
           "engine": f"{alg.predict_engine} ({alg.predict_slices} slices)" if alg.predict_engine == "i8" else "f64",
           "flagged_points": getattr(alg, "last_flagged", None),
           "fit_flagged_samples": getattr(alg, "last_fit_flagged", None),  # 0: the backward recursion ran on the i8 engine too
           "note": "predict = K(X,pts)^T [diag(coef_r) CUA]_r per chunk of test points + fused max/step; engine i8: Ozaki "
                   "digit splitting, exact int8 products on tcgen05.mma kind::i8 (predict_gemm_tflops is then the "
                   "fp64-EQUIVALENT rate of the contraction), FP64 DMMA for flagged points"}
    alg.predict_engine = "f64"
    alg.predict_dev(pts[:16384])
    t64, _ = timed(lambda: alg.predict_dev(pts))
    res["predict_s_f64_engine"] = t64
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
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the KernelMaximalSR side measurement and the second scaling mode")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): --m test points in total, split over the ranks; weak: --m per GPU")
    ap.add_argument("--ref-state", default="reference", choices=["reference", "port"],
                    help="--impl reference: 'reference' = the unmodified gym_socks code (fit + predict), 'port' = numpy port")
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
