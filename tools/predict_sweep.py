"""Time the predict forms / CTA shapes of socks_sr_predict_packed (csrc/sr_predict.cu) with CUDA events.
Usage: python tools/predict_sweep.py [out.json]   (B200; the fit runs once per (N, d))"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from socks_b200 import _device as dv, synthetic as syn
from socks_b200.algorithms.reach.kernel_sr import KernelSR

NAMES = {0: "auto (factored fp64 form on DMMA, shape by M)", 5: "same as 0", 6: "integer tensor-core engine (tcgen05 kind::i8)", 1: "FMA pipe + CUDA exp", 2: "direct form on DMMA (r1 kernel)",
         3: "factored, 8-warp CTAs x3, unroll 2", 4: "factored, 4-warp CTAs x5, unroll 2",
         10: "factored, 8 warps x3, unroll 4", 11: "factored, 8 warps x2, unroll 2", 12: "factored, 8 warps x2, unroll 4",
         13: "attribution (vs 11): +4 integer instr per value", 14: "attribution (vs 11): +1 FP64 add per value",
         15: "attribution (vs 11): no exponent insertion, -2 int (wrong results)",
         16: "attribution (vs 11): no table gather (wrong results)", 17: "attribution (vs 11): +4 int +1 FP64"}
VARIANTS = [int(v) for v in os.environ.get("SWEEP_VARIANTS", "0,2,3,4,1").split(",")]
SIZES = os.environ.get("SWEEP_SIZES", "full")


def timed(fn, reps):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    res = []
    cases = ((20000, 2, 15, (250000, 125000, 62500, 31250)), (40000, 4, 15, (500000, 62500)))
    if SIZES == "c2":
        cases = ((20000, 2, 15, (250000, 31250)),)
    for (n, d, T, ms) in cases:
        X, U, Y = syn.integrator_sample(n, d, seed=0)
        ct, tt = syn.integrator_tubes(T, d)
        alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", regularization_param=1 / n ** 2)
        alg._keep_factors = False
        alg.fit_arrays(X, Y)
        ref = None
        for m in ms:
            ref = None
            pts = dv.to_dev(syn.uniform_points(m, d, seed=1))
            for variant in VARIANTS:
                if variant == 1 and m > 125000:
                    continue
                alg.predict_variant = variant
                med, best = timed(lambda: alg.predict_dev(pts), 7 if variant != 1 else 3)
                out = alg.predict_dev(pts)
                if ref is None:
                    ref = out
                flops = float(n) * m * (3 * d + 2 + 12 + 2 * T)
                rec = {"N": n, "d": d, "T": T, "M": m, "variant": variant, "name": NAMES[variant], "ms_median": med,
                       "ms_best": best, "Mpts_per_s": m / med / 1e3, "tflops_algorithmic": flops / med / 1e9,
                       "max_abs_diff_vs_first": float(torch.nan_to_num(out - ref).abs().max()),
                       "bit_identical_to_first": bool(torch.equal(torch.nan_to_num(out), torch.nan_to_num(ref)))}
                print(json.dumps(rec), flush=True)
                res.append(rec)
        del alg
        torch.cuda.empty_cache()
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
