"""Accuracy of socks_potrf_inv's GEMM engines: residual of W = inv(G + I/N) on 64 columns and diag(W) against the FP64
DMMA-only variant.  Usage: python tools/inverse_residual.py [N] [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from socks_b200 import _device as dv, _lib, synthetic as syn
from socks_b200.kernel import metrics as mt

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
lib = _lib.load()
X, U, Y = syn.integrator_sample(n, seed=0)
Xd = dv.to_dev(X)
res, ref_w = [], None
for name, upd, mrg, dg in (("fp64 DMMA everywhere", -3, -5, -8), ("integer updates (7 digits), DMMA merges", -4, -5, -8),
                           ("integer updates (7 digits) and merges (7 digits)", -4, -6, -7),
                           ("integer updates (7 digits) and merges (8 digits): shipped", -4, -6, -8)):
    mt.set_factorization_engine("fp64" if upd == -3 else "int8")  # also selects the engine of W = M^T M (socks_lauum[_i8])
    lib.socks_potrf_inv_set_block(upd)
    lib.socks_potrf_inv_set_block(mrg)
    lib.socks_potrf_inv_set_block(dg)
    fac = mt.Factorization(Xd, (-50.0, 0), 1.0 / n)
    fac.check()
    w = fac.diag().clone()
    Wd = fac.full()
    G = mt.gram_dev(Xd, Xd[:64], (-50.0, 0))[:, :64]
    G[torch.arange(64), torch.arange(64)] += 1.0 / n
    R = Wd[:n, :n] @ G  # torch matmul only as a checker
    R[torch.arange(64), torch.arange(64)] -= 1.0
    if ref_w is None:
        ref_w = w
    rec = {"N": n, "engine": name + ("; W = M^T M on FP64 DMMA" if upd == -3 else "; W = M^T M as an 8-digit integer GEMM"),
           "residual_max_WG_minus_I": float(R.abs().max()),
           "diag_W_max_rel_diff_vs_dmma": float(((w - ref_w).abs() / ref_w.abs()).max())}
    print(json.dumps(rec), flush=True)
    res.append(rec)
    del fac, Wd, R, G
    torch.cuda.empty_cache()
mt.set_factorization_engine("int8")
lib.socks_potrf_inv_set_block(-8)
if len(sys.argv) > 2:
    json.dump(res, open(sys.argv[2], "w"), indent=1)
