"""2..8-rank check of the library's own NCCL path (socks_ctx_* / socks_comm_init / socks_bcast / socks_allgather):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/comm_check.py [out.json]
Broadcasts a fitted KernelSR state from rank 0, predicts a shard per rank, all-gathers the shards and compares with the
single-GPU result bit for bit; times the broadcast of a W-sized buffer (BASELINE.json configs[3])."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from socks_b200 import _device as dv, synthetic as syn
from socks_b200.algorithms.reach.kernel_sr import KernelSR
from socks_b200.distributed import SocksComm, init_from_env, shard_bounds

rank, world, local = init_from_env()
comm = SocksComm()
n, m, T, d = 4000, 40000, 15, 2
ct, tt = syn.integrator_tubes(T, d)
alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", regularization_param=1 / n ** 2)
alg._validate_params(None)
pts = dv.to_dev(syn.uniform_points(m, d, seed=1))
if rank == 0:
    X, U, Y = syn.integrator_sample(n, d, seed=0)
    alg.fit_arrays(X, Y)
    want = alg.predict_dev(pts)
else:
    alg.allocate_state(n, d)
nbytes = comm.broadcast_state(alg.state_tensors(), src=0)
lo, hi = shard_bounds(m, world, rank)
width = -(-m // world)
local_out = torch.zeros(T, width, dtype=torch.float64, device="cuda")
local_out[:, : hi - lo] = alg.predict_dev(pts[lo:hi])
gathered = comm.allgather(local_out.contiguous()).view(world, T, width)
ok = True
if rank == 0:
    full = torch.empty(T, m, dtype=torch.float64, device="cuda")
    for r in range(world):
        s, e = shard_bounds(m, world, r)
        full[:, s:e] = gathered[r][:, : e - s]
    ok = bool(torch.equal(full, want))
# broadcast bandwidth on a large buffer (3.2 GB = W at N = 20 000)
big = torch.empty(3_200_000_000 // 8, dtype=torch.float64, device="cuda")
if rank == 0:
    big.fill_(1.5)
comm.broadcast_state({"w": big})
torch.cuda.synchronize()
t0 = time.perf_counter()
comm.broadcast_state({"w": big})
dt = time.perf_counter() - t0
good = bool((big[::100003] == 1.5).all())
if rank == 0:
    res = {"ranks": world, "state_bytes": nbytes, "sharded_equals_single_gpu_bitwise": ok, "bcast_3p2GB_s": dt,
           "bcast_GBps": 3.2 / dt, "bcast_payload_ok": good}
    print(json.dumps(res))
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)
comm.close()
torch.distributed.destroy_process_group()
