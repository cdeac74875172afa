"""Timeline of socks_potrf_inv (CUDA events on the panel stream and on the caller's stream).
Usage: python tools/potrf_trace.py [N] [out.json]"""
import ctypes
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
for _ in range(2):  # warm-up
    fac = mt.Factorization(Xd, (-50.0, 0), 1.0 / n)
    fac.check()
    del fac
lib.socks_dev_potrf_trace(None, None, 0)
fac = mt.Factorization(Xd, (-50.0, 0), 1.0 / n)
torch.cuda.synchronize()
cap = 4096
codes = (ctypes.c_int * cap)()
ms = (ctypes.c_double * cap)()
cnt = lib.socks_dev_potrf_trace(codes, ms, cap)
KIND = {0: "start", 1: "block", 2: "panel", 3: "small_merges", 4: "col_update", 5: "rest_update", 6: "join", 7: "large_merge"}
ev = [{"kind": KIND[codes[i] // 1000], "step": codes[i] % 1000, "ms": ms[i]} for i in range(cnt)]
by = {}
for e in ev:
    by.setdefault(e["kind"], {})[e["step"]] = e["ms"]
print(f"{'k':>3} {'block':>8} {'panel':>8} {'merges':>8} | {'col':>8} {'rest':>8}   (ms since start; s2 | s1)")
for k in sorted(by.get("block", {})):
    g = lambda kind: by.get(kind, {}).get(k, float("nan"))
    print(f"{k:3d} {g('block'):8.2f} {g('panel'):8.2f} {g('small_merges'):8.2f} | {g('col_update'):8.2f} {g('rest_update'):8.2f}")
print("join", by["join"][0])
print("large-merge products done at (product, span in block columns, ms):",
      [(e["step"] // 100, e["step"] % 100, round(e["ms"], 2)) for e in ev if e["kind"] == "large_merge"])
tot = max(e["ms"] for e in ev)
print("total ms", tot)
if len(sys.argv) > 2:
    json.dump({"N": n, "events": ev, "total_ms": tot}, open(sys.argv[2], "w"), indent=1)
