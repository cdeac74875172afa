"""Per-stage device times of the KernelSR path (CUDA events) with algorithmic work and roofline fractions.
Usage (on a B200): python tools/stage_times.py [N] [M] [T] [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from socks_b200 import _device as dv, _lib, synthetic as syn
from socks_b200.algorithms.reach.common import PROBLEM_CODE, pack_tubes
from socks_b200.kernel import metrics as mt

EXP_FLOPS = 31  # CUDA exp(double) on sm_100a: 14 DFMA (28) + 3 other FP64-pipe ops


def timed(fn, reps=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps, r


def main(n=20000, m=250000, T=15, out_path=None, hbm_gbs=6546.2):
    lib = _lib.load()
    d, sigma = 2, 0.1
    gamma = 1 / (2 * sigma ** 2)
    X, U, Y = syn.integrator_sample(n, seed=0)
    pts = syn.uniform_points(m, seed=1)
    ct, tt = syn.integrator_tubes(T)
    Xd, Yd, Pd = dv.to_dev(X), dv.to_dev(Y), dv.to_dev(pts)
    npad = dv.padded(n)
    st = dv.stream()
    res = {"N": n, "M": m, "T": T, "d": d}
    # warm-up of every kernel on a small problem (module load, smem attribute calls)
    fac0 = mt.Factorization(dv.to_dev(X[:512]), (-gamma, 0), 1e-3)
    fac0.full(); fac0.diag()
    L = dv.empty(npad, npad)
    t, _ = timed(lambda: _lib.call("socks_gram_sym", dv.ptr(Xd), n, d, -gamma, 0, 1.0 / n, dv.ptr(L), npad, st))
    res["gram_sym_s"] = t
    dinv = dv.work(lib.socks_potrf_work_bytes(n))
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    t, _ = timed(lambda: _lib.call("socks_potrf", dv.ptr(L), n, npad, dv.ptr(dinv), dv.ptr(info), st))
    res["potrf_s"] = t
    res["potrf_tflops"] = npad ** 3 / 3 / t / 1e12
    res["potrf_info"] = int(info.item())
    M = dv.empty(npad, npad)
    wk = dv.work(lib.socks_trtri_work_bytes(n))
    t, _ = timed(lambda: _lib.call("socks_trtri", dv.ptr(L), n, npad, dv.ptr(dinv), dv.ptr(M), npad, dv.ptr(wk), st))
    res["trtri_s"] = t
    res["trtri_tflops"] = npad ** 3 / 3 / t / 1e12
    # fused factor + invert (what the algorithms use): G regenerated, then one recursion for L and inv(L)
    L2, M2 = dv.empty(npad, npad), dv.empty(npad, npad)
    info2 = torch.zeros(1, dtype=torch.int32, device="cuda")
    # one untimed call first: internal streams, kernel attributes and the first touch of the workspace are one-off costs
    _lib.call("socks_gram_sym", dv.ptr(Xd), n, d, -gamma, 0, 1.0 / n, dv.ptr(L2), npad, st)
    _lib.call("socks_potrf_inv", dv.ptr(L2), n, npad, dv.ptr(M2), npad, dv.ptr(wk), dv.ptr(info2), 0, st)
    _lib.call("socks_gram_sym", dv.ptr(Xd), n, d, -gamma, 0, 1.0 / n, dv.ptr(L2), npad, st)
    t, _ = timed(lambda: _lib.call("socks_potrf_inv", dv.ptr(L2), n, npad, dv.ptr(M2), npad, dv.ptr(wk), dv.ptr(info2), 0, st))
    res["potrf_inv_s"] = t
    res["potrf_inv_tflops"] = 2 * npad ** 3 / 3 / t / 1e12
    blk = slice(n - 2048, n)  # last rows of inv(L): compare the two routes (rounding differs, ~cond * eps)
    res["potrf_inv_vs_two_pass_reldiff"] = float((torch.tril(M2[blk, blk]) - torch.tril(M[blk, blk])).abs().max()
                                                 / torch.tril(M[blk, blk]).abs().max())
    del L2, M2
    w = dv.empty(n)
    rwk = dv.work(lib.socks_reduce_work_bytes(npad) + 8 * (npad + 2))
    t, _ = timed(lambda: _lib.call("socks_diag_inv", dv.ptr(M), n, npad, dv.ptr(w), dv.ptr(rwk), st))
    res["diag_inv_s"] = t
    res["diag_inv_gbs"] = 8 * npad * npad / 2 / t / 1e9
    W = L  # reuse the 3.2 GB buffer for inv(G)
    t, _ = timed(lambda: _lib.call("socks_lauum", dv.ptr(M), n, npad, dv.ptr(W), npad, st))
    res["lauum_s"] = t
    res["lauum_tflops"] = npad ** 3 / 3 / t / 1e12
    ldb = dv.even(n)
    B = dv.empty(n, ldb)
    t, _ = timed(lambda: mt.gram_dev(Xd, Yd, (-gamma, 0), rowscale=w, out=B, ld=ldb))
    res["gram_cross_s"] = t
    res["gram_cross_gbs_written"] = 8 * n * n / t / 1e9
    res["gram_cross_fp64_tflops"] = n * n * (3 * d + 2 + EXP_FLOPS) / t / 1e12
    tubes = pack_tubes(ct, tt, T, d)
    vf = dv.empty(T, n)
    rec = lambda: _lib.call("socks_sr_recursion", dv.ptr(B), n, n, ldb, dv.ptr(Yd), d, dv.ptr(tubes),
                            PROBLEM_CODE["FHT"], T, dv.ptr(vf), n, dv.ptr(vf), n, dv.ptr(rwk), st)
    rec()  # first launch of each kernel pays CUDA's lazy module loading (~1.3 ms here): keep it out of the timing
    t, _ = timed(lambda: _lib.call("socks_sr_recursion", dv.ptr(B), n, n, ldb, dv.ptr(Yd), d, dv.ptr(tubes),
                                   PROBLEM_CODE["FHT"], T, dv.ptr(vf), n, dv.ptr(vf), n, dv.ptr(rwk), st))
    res["fit_recursion_s"] = t
    res["fit_recursion_gbs"] = T * 8 * n * n / t / 1e9  # T passes: 1 normaliser + T-1 steps
    res["fit_recursion_frac_hbm"] = res["fit_recursion_gbs"] / hbm_gbs
    # steady-state cost of one GEMV pass alone (15 back-to-back passes, no step kernels in between)
    yv = dv.empty(1, n)
    f = lambda: _lib.call("socks_gemv_t", dv.ptr(B), n, n, ldb, dv.ptr(vf), n, 1, dv.ptr(yv), n, dv.ptr(rwk), st)
    f()
    t, _ = timed(f, reps=15)
    res["gemv_pass_s"] = t
    res["gemv_pass_gbs"] = 8 * n * n / t / 1e9
    out = dv.empty(T, m)
    coef = dv.work(lib.socks_sr_predict_work_bytes(n, m))
    for variant in (0, 1):
        f = lambda: _lib.call("socks_sr_predict", dv.ptr(Xd), dv.ptr(w), dv.ptr(vf), n, n, d, -gamma, 0, dv.ptr(Pd), m,
                              dv.ptr(tubes), PROBLEM_CODE["FHT"], T, dv.ptr(out), m, dv.ptr(coef), variant, st)
        f()
        t, _ = timed(f, reps=3)
        key = "predict_dmma" if variant == 0 else "predict_fma"
        res[key + "_s"] = t
        res[key + "_pts_per_s"] = m / t
        # algorithmic flops as in bench.py: distance 3d+2, exp (table variant 12 / CUDA exp 31), contraction 2T
        res[key + "_tflops"] = n * m * (3 * d + 2 + (12 if variant == 0 else EXP_FLOPS) + 2 * T) / t / 1e12
    res["gram_solve_two_pass_s"] = res["gram_sym_s"] + res["potrf_s"] + res["trtri_s"] + res["diag_inv_s"]
    res["gram_solve_s"] = res["gram_sym_s"] + res["potrf_inv_s"] + res["diag_inv_s"]
    res["vf_range"] = [float(vf.min()), float(vf.max())]
    print(json.dumps(res))
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)
    return res


if __name__ == "__main__":
    a = sys.argv[1:]
    out = a.pop() if a and a[-1].endswith(".json") else None  # python tools/stage_times.py [N [M [T]]] [out.json]
    main(int(a[0]) if len(a) > 0 else 20000, int(a[1]) if len(a) > 1 else 250000, int(a[2]) if len(a) > 2 else 15, out)
