"""diag(L)-based conditioning estimate (Factorization.cond_estimate) of the BASELINE configs: all below
metrics.COND_LIMIT_INT8, i.e. on the integer tensor-core engine.  Usage: python tools/cond_estimates.py"""
import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from socks_b200 import _device as dv, synthetic as syn
from socks_b200.kernel import metrics as mt
mt.COND_LIMIT_INT8 = float("inf")
def est(name, X, U, spec, lam):
    fac = mt.factorize(X, U, spec, lam); fac.check(); print(name, "cond_estimate %.3e" % fac.cond_estimate, flush=True); del fac; torch.cuda.empty_cache()
X,U,Y = syn.integrator_sample(1000, seed=0); est("C1 N=1000", X, None, mt.RBFSpec(sigma=0.1), 1/1000**2)
X,U,Y = syn.integrator_sample(20000, seed=0); est("C2 N=20000", X, None, mt.RBFSpec(sigma=0.1), 1/20000**2)
est("C2 N=20000 lambda=1 (reference default)", X, None, mt.RBFSpec(sigma=0.1), 1.0)
est("C2max product Gram lambda=1", X, U, mt.RBFSpec(sigma=0.1), 1.0)
X,U,Y = syn.unicycle_sample(5000, seed=12345); est("C3 N=5000 sigma=3 lambda=1e-7 product", X, U, mt.RBFSpec(sigma=3.0), 1e-7)
X,U,Y = syn.integrator_sample(40000, seed=0); est("C4 N=40000", X, None, mt.RBFSpec(sigma=0.1), 1/40000**2)
X,U,Y = syn.integrator_sample(40000, 4, seed=0); est("C5 N=40000 d=4", X, None, mt.RBFSpec(sigma=0.1), 1/40000**2)
