"""Accuracy of the factorisation engines on ill-conditioned Gram matrices: residual max|W G - I| with the FP64 engine, the
integer engine without its conditioning guard, and the shipped behaviour (metrics.COND_LIMIT_INT8).  Usage: python tools/illcond_check.py"""
import os
import sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from socks_b200 import _device as dv, _lib
from socks_b200.kernel import metrics as mt
rng = np.random.default_rng(3)
for n, sigma, lam in ((3000, 0.5, 1e-9), (3000, 1.0, 1e-11), (6000, 0.3, 1e-8), (6000, 2.0, 1e-12), (6000, 0.1, 1.0 / 6000 ** 2),
                      (5000, 0.2, 1e-6)):
    X = rng.uniform(-1.1, 1.1, (n, 2))
    G = np.exp(-((X[:, None, :] - X[None, :, :]) ** 2).sum(-1) / (2 * sigma ** 2))
    G[np.diag_indices(n)] += lam * n
    ev = np.linalg.eigvalsh(G)
    cond = ev[-1] / ev[0]
    out = {"n": n, "sigma": sigma, "lambda": lam, "cond": float(cond)}
    for engine in ("fp64", "int8 (guard off)", "int8 (shipped: inverse redone on fp64 above COND_LIMIT_INT8)"):
        mt.set_factorization_engine(engine.split()[0])
        limit, mt.COND_LIMIT_INT8 = mt.COND_LIMIT_INT8, (float("inf") if "guard off" in engine else mt.COND_LIMIT_INT8)
        try:
            fac = mt.factorize(X, None, mt.RBFSpec(sigma=sigma), lam)
            fac.check()  # as every algorithm does right after the factorisation: this is where the guard acts
            W = fac.full()[:n, :n].cpu().numpy()
            R = W @ G - np.eye(n)
            out[engine + ": residual_max"] = float(np.abs(R).max())
            out[engine + ": cond_estimate"] = float(fac.cond_estimate)
        except Exception as e:
            out[engine + ": error"] = repr(e)[:120]
        finally:
            mt.COND_LIMIT_INT8 = limit
    mt.set_factorization_engine("int8")
    print(json.dumps(out), flush=True)
