"""FP64-equivalent GEMM on the integer tensor cores (socks_dev_ozaki_gemm_nt) against the DMMA GEMM and a float128-free
exact check: accuracy on operands with a wide dynamic range, and speed on the factorisation's update shapes.
Usage: python tools/ozaki_gemm_check.py [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from socks_b200 import _device as dv, _lib

lib = _lib.load()
res = []


def run(m, n, K, syrk, lower, reps=3, wide=False):
    g = torch.Generator(device="cuda").manual_seed(m + n + K)
    A = torch.randn(m, K, dtype=torch.float64, device="cuda", generator=g)
    if wide:  # rows and columns spanning 12 decades
        A = A * torch.pow(10.0, torch.empty(m, 1, dtype=torch.float64, device="cuda").uniform_(-6, 6, generator=g))
        A = A * torch.pow(10.0, torch.empty(1, K, dtype=torch.float64, device="cuda").uniform_(-3, 3, generator=g))
    B = A if syrk else torch.randn(n, K, dtype=torch.float64, device="cuda", generator=g)
    C0 = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
    wk = dv.work(lib.socks_dev_ozaki_gemm_work_bytes(m, n, K) + 2048)
    base = (wk.data_ptr() + 1023) // 1024 * 1024
    st = dv.stream()

    def oz(C):
        _lib.call("socks_dev_ozaki_gemm_nt", A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), n, m, n, K, 1 if syrk else 0,
                  1 if lower else 0, base, st)

    def dm(C):
        _lib.call("socks_gemm_f64", 0, 1, A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), n, m, n, K, -1.0, 1.0, 0,
                  1 if lower else 0, 0, st)

    C1, C2 = C0.clone(), C0.clone()
    oz(C1)
    dm(C2)
    torch.cuda.synchronize()
    # reference on a sample of entries in double-double via torch (split A, B into hi/lo halves for an exact-ish product)
    rows = torch.arange(0, m, max(1, m // 64), device="cuda")[:64]
    Ah, Bh = A[rows], B
    prod = (Ah.unsqueeze(1) * Bh.unsqueeze(0))  # (64, n, K) fp64 products
    # sum in extended precision: sort by magnitude + fp64 fsum-like pairwise (torch.sum is pairwise): use math.fsum on CPU for a few entries
    import math
    cols = list(range(0, n, max(1, n // 16)))[:16]
    errs_oz, errs_dm = [], []
    pc = prod[:, cols, :].cpu()
    for a in range(len(rows)):
        for b, cj in enumerate(cols):
            if lower and cj > int(rows[a]) // 128 * 128 + 127:
                continue
            exact = float(C0[rows[a], cj]) - math.fsum(pc[a, b].tolist())
            scale = float((Ah[a].abs() * Bh[cj].abs()).sum()) + abs(float(C0[rows[a], cj]))
            errs_oz.append(abs(float(C1[rows[a], cj]) - exact) / scale)
            errs_dm.append(abs(float(C2[rows[a], cj]) - exact) / scale)

    def timed(f):
        C = C0.clone()
        f(C)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f(C)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    t_oz, t_dm = timed(oz), timed(dm)
    tiles = (m // 128) * (m // 128 + 1) // 2 if lower else (m // 128) * (n // 128)
    fl = 2.0 * tiles * 128 * 128 * K
    rec = {"m": m, "n": n, "K": K, "syrk": syrk, "lower": lower, "wide_range": wide, "ozaki_ms": t_oz * 1e3,
           "dmma_ms": t_dm * 1e3, "ozaki_tflops_equiv": fl / t_oz / 1e12, "dmma_tflops": fl / t_dm / 1e12,
           "ozaki_max_err_rel_to_sum_abs": max(errs_oz), "dmma_max_err_rel_to_sum_abs": max(errs_dm)}
    print(json.dumps(rec), flush=True)
    res.append(rec)


run(256, 256, 64, False, False)
run(512, 384, 1024, False, False, wide=True)
run(2048, 2048, 1024, True, True, wide=True)
run(10112, 10112, 1024, True, True)
run(19072, 19072, 1024, True, True)
run(19072, 1024, 1024, False, False)
run(10112, 10112, 2048, True, True)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
