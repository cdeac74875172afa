"""Time the fp64 DMMA GEMM (socks_gemm_f64) on the shapes the factorisation uses. Usage: gemm_bench.py [quick]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from socks_b200 import _device as dv, _lib

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"


def run(ta, tb, m, n, K, lower, cfg, flags=0, reps=3):
    A = torch.randn((K, m) if ta else (m, K), dtype=torch.float64, device="cuda")
    B = torch.randn((n, K) if tb else (K, n), dtype=torch.float64, device="cuda")
    C = torch.zeros(m, n, dtype=torch.float64, device="cuda")
    st = dv.stream()
    f = lambda: _lib.call("socks_gemm_f64", ta, tb, dv.ptr(A), A.shape[1], dv.ptr(B), B.shape[1], dv.ptr(C), n, m, n, K,
                          -1.0, 1.0, flags, lower, cfg, st)
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / reps
    tm, tn = m // 128, n // 128
    tiles = tm * (tm + 1) // 2 if lower else tm * tn
    fl = 2.0 * tiles * 128 * 128 * K
    return {"ta": ta, "tb": tb, "m": m, "n": n, "K": K, "lower": lower, "cfg": cfg, "ms": t * 1e3, "tflops": fl / t / 1e12}


cfgs = (0, 1)
shapes = [(0, 1, 10112, 10112, 9984, 1), (0, 0, 10112, 9984, 9984, 0), (0, 1, 2560, 2560, 2560, 1), (0, 1, 640, 640, 640, 0),
          (0, 1, 128, 128, 128, 0)] if quick else [
    (0, 1, 10112, 10112, 9984, 1), (0, 0, 10112, 9984, 9984, 0), (1, 0, 8192, 8192, 8192, 0),
    (0, 1, 5120, 5120, 4992, 1), (0, 1, 10112, 1280, 1280, 0), (0, 1, 2560, 2560, 2560, 1),
    (0, 1, 1280, 1280, 1280, 1), (0, 1, 640, 640, 640, 0), (0, 1, 384, 256, 256, 0), (0, 1, 128, 128, 128, 0),
    (0, 0, 256, 128, 256, 0),
]
for ta, tb, m, n, K, lower in shapes:
    for cfg in cfgs:
        print(json.dumps(run(ta, tb, m, n, K, lower, cfg)), flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "lookahead":
    # shapes of socks_potrf_inv's look-ahead factorisation at N = 20 000 (block columns of 1024)
    print("# rank-1024 trailing updates (lower-only SYRK), column updates, panels (B upper-clipped), merges")
    for m in (19072, 15104, 10112, 5120, 2048):
        r = run(0, 1, m, m, 1024, 1, 0)
        r["what"] = "trailing update A22 -= L21 L21^T"
        print(json.dumps(r), flush=True)
    for m in (19072, 10112, 2048):
        r = run(0, 1, m, 1024, 1024, 0, 0)
        r["what"] = "next-block-column update / panel-shaped GEMM"
        print(json.dumps(r), flush=True)
    for n1 in (1024, 2048, 4096, 10112):
        for flags, name in ((1, "T = M22 L21 (A lower: k <= m)"), (4, "M21 = -T M11 (B: k >= n)")):
            r = run(0, 0, n1, n1, n1, 0, 0, flags=flags)
            r["tflops_triangular"] = r["tflops"] / 2  # about half of the k-range is clipped
            r["what"] = name
            print(json.dumps(r), flush=True)
