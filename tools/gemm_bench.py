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
