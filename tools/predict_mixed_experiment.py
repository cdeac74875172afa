"""Experiment: the fp64 predict kernel and the integer tensor-core engine on two streams at once, each on its own share of
the points (are the two kernels complementary on an SM?).  Timing only.  Usage: python tools/predict_mixed_experiment.py"""
import copy
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from socks_b200 import _device as dv, synthetic as syn
from socks_b200.algorithms.reach.kernel_sr import KernelSR

n, d, T, m = 20000, 2, 15, 250000
X, U, Y = syn.integrator_sample(n, d, seed=0)
ct, tt = syn.integrator_tubes(T, d)
alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", regularization_param=1 / n ** 2)
alg._keep_factors = False
alg.fit_arrays(X, Y)
pts = dv.to_dev(syn.uniform_points(m, d, seed=1))
ref = alg.predict_dev(pts).clone()
b = copy.copy(alg)
b.__dict__.pop("_buf_cache", None)
b.__dict__.pop("_work", None)
b.predict_variant = 6
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(frac_i8, fp64_variant):
    alg.predict_variant = fp64_variant
    m8 = int(m * frac_i8) // 128 * 128
    pa, pb = pts[: m - m8], pts[m - m8:]
    outs = [None, None]

    def once():
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        sa.wait_stream(torch.cuda.current_stream())
        sb.wait_stream(torch.cuda.current_stream())
        if m - m8 > 0:
            with torch.cuda.stream(sa):
                outs[0] = alg.predict_dev(pa)
        if m8 > 0:
            with torch.cuda.stream(sb):
                outs[1] = b.predict_dev(pb)
        torch.cuda.current_stream().wait_stream(sa)
        torch.cuda.current_stream().wait_stream(sb)
        e1.record()
        return e0, e1

    for _ in range(3):
        once()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = once()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    got = torch.cat([o for o in outs if o is not None], dim=1)
    err = float(torch.nan_to_num(got - ref).abs().max())
    rec = {"fraction_on_integer_engine": frac_i8, "fp64_variant": fp64_variant, "ms_median": float(np.median(ts)),
           "Mpts_per_s": m / float(np.median(ts)) / 1e3, "max_abs_diff_vs_fp64_only": err}
    print(json.dumps(rec), flush=True)
    return rec


res = []
for fv in (0, 4):
    for fr in (0.0, 0.25, 0.35, 0.45, 0.55, 1.0):
        res.append(run(fr, fv))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
