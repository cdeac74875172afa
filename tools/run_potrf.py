"""Factorise one synthetic regularised Gram matrix (for ncu launch lists). Usage: python tools/run_potrf.py [N] [what]
what: potrf | inv (potrf + trtri + diag) | full (+ lauum) | sweep (time socks_potrf_inv for several look-ahead block widths)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from socks_b200 import _device as dv, synthetic as syn
from socks_b200.kernel import metrics as mt

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
what = sys.argv[2] if len(sys.argv) > 2 else "potrf"
X, U, Y = syn.integrator_sample(n, seed=0)
if what == "sweep":
    import json

    from socks_b200 import _lib

    lib = _lib.load()
    Xd = dv.to_dev(X)
    npad = dv.padded(n)
    mt.Factorization(dv.to_dev(X[:3000]), (-50.0, 0), 1.0 / 3000)  # warm-up: module load, smem attributes, streams
    out = []
    for nb in [int(a) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else (0, 512, 1024, 1536, 2048):
        tail, ozaki, ozmerge = 1, 1, 1
        if nb >= 400000:  # e.g. 401024: block 1024, integer trailing updates but FP64 DMMA merges
            ozmerge, nb = 0, nb - 400000
        if nb >= 200000:  # e.g. 201024: block 1024 with FP64 DMMA trailing updates
            ozaki, nb = 0, nb - 200000
        if nb >= 100000:  # e.g. 101024: block 1024 without the half-width tail
            tail, nb = 0, nb - 100000
        lib.socks_potrf_inv_set_block(-4 if ozaki else -3)
        lib.socks_potrf_inv_set_block(-6 if ozmerge else -5)
        lib.socks_potrf_inv_set_block(-2 if tail else -1)
        old = lib.socks_potrf_inv_set_block(nb)
        ts = []
        for rep in range(3):
            dv.stage_marks = []
            fac = mt.Factorization(Xd, (-50.0, 0), 1.0 / n)
            torch.cuda.synchronize()
            ts.append(dv.stage_times()["potrf_inv"])
            dv.stage_marks = None
            fac.check()
            del fac
        lib.socks_potrf_inv_set_block(old)
        t = min(ts)
        lib.socks_potrf_inv_set_block(-2)
        lib.socks_potrf_inv_set_block(-4)
        lib.socks_potrf_inv_set_block(-6)
        rec = {"N": n, "block": nb, "half_width_tail": bool(tail),
               "trailing_updates": "int8 tensor cores (Ozaki)" if ozaki else "FP64 DMMA",
               "large_merges": "int8 tensor cores (Ozaki)" if (ozaki and ozmerge) else "FP64 DMMA", "potrf_inv_s": t, "tflops": 2 * npad ** 3 / 3 / t / 1e12, "all_s": ts}
        print(json.dumps(rec), flush=True)
        out.append(rec)
    if len(sys.argv) > 4:
        json.dump(out, open(sys.argv[4], "w"), indent=1)
    sys.exit(0)
fac = mt.Factorization(dv.to_dev(X), (-50.0, 0), 1.0 / n)
if what in ("inv", "full"):
    fac.diag()
if what == "full":
    fac.full()
torch.cuda.synchronize()
fac.check()
print("ok", n, what)
