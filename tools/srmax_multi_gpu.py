"""KernelMaximalSR (BASELINE configs[1]) across GPUs: fit on rank 0, broadcast (X, w, value functions, CUA) over
NCCL, contiguous blocks of the test points per rank, all-gather of the result.
    python -m torch.distributed.run --nproc-per-node R tools/srmax_multi_gpu.py [N] [M] [P]"""
import json
import os
import sys
import time
from functools import partial

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from socks_b200 import _device as dv, synthetic as syn
from socks_b200.algorithms.reach.kernel_sr_max import KernelMaximalSR
from socks_b200.distributed import broadcast_state, gather_columns, init_from_env, shard_bounds
from socks_b200.kernel.metrics import rbf_kernel

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 250000
P = int(sys.argv[3]) if len(sys.argv) > 3 else 100
T, d = 15, 2
rank, world, local = init_from_env()
ct, tt = syn.integrator_tubes(T, d)
alg = KernelMaximalSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT",
                      kernel_fn=partial(rbf_kernel, sigma=0.1), regularization_param=1.0)
res = {"N": n, "M": m, "P": P, "T": T, "world": world}
if rank == 0:
    X, U, Y = syn.integrator_sample(n, d, seed=0)
    A = np.linspace(-1, 1, P).reshape(-1, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    alg.fit_arrays(X, U, Y, A)
    torch.cuda.synchronize()
    res["fit_s"] = time.perf_counter() - t0
else:
    alg.allocate_state(n, d, P)
res["broadcast_bytes"] = broadcast_state(alg.state_tensors(), src=0)
pts = dv.to_dev(syn.uniform_points(m, d, seed=1))
s, e = shard_bounds(m)
alg.predict_dev(pts[s:s + 4096])
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
local_out = alg.predict_dev(pts[s:e])
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
full = gather_columns(local_out.contiguous(), m)
res["predict_s_max_over_ranks"] = float(t)
res["points_per_s"] = m / float(t)
res["checksum"] = float(full.sum())
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
