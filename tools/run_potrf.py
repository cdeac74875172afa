"""Factorise one synthetic regularised Gram matrix (for ncu launch lists). Usage: python tools/run_potrf.py [N] [what]
what: potrf | inv (potrf + trtri + diag) | full (+ lauum)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from socks_b200 import _device as dv, synthetic as syn
from socks_b200.kernel import metrics as mt

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
what = sys.argv[2] if len(sys.argv) > 2 else "potrf"
X, U, Y = syn.integrator_sample(n, seed=0)
fac = mt.Factorization(dv.to_dev(X), (-50.0, 0), 1.0 / n)
if what in ("inv", "full"):
    fac.diag()
if what == "full":
    fac.full()
torch.cuda.synchronize()
fac.check()
print("ok", n, what)
