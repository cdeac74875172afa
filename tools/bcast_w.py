"""BASELINE.json configs[3]: regularized_inverse at N = 40 000 on ONE B200 (rank 0), then W (12.8 GB) broadcast to
the other ranks over NCCL/NVLink.  Run under torchrun:  python -m torch.distributed.run --nproc-per-node R tools/bcast_w.py [N]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from socks_b200 import _device as dv, synthetic as syn
from socks_b200.distributed import init_from_env
from socks_b200.kernel.metrics import RBFSpec, factorize

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
rank, world, local = init_from_env()
npad = dv.padded(n)
res = {"N": n, "world": world}
if rank == 0:
    X, U, Y = syn.integrator_sample(n, seed=0)
    factorize(X[:1024], None, RBFSpec(sigma=0.1), 1e-3).full()  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fac = factorize(X, None, RBFSpec(sigma=0.1), 1 / n ** 2)
    W = fac.full()
    torch.cuda.synchronize()
    res["gram_potrf_trtri_lauum_s"] = time.perf_counter() - t0
    res["tflops"] = float(npad) ** 3 / res["gram_potrf_trtri_lauum_s"] / 1e12
    fac.check()
    fac.release_factors()
else:
    W = dv.empty(npad, npad)
if world > 1:
    small = dv.empty(1024)
    dist.broadcast(small, src=0)  # communicator warm-up
    torch.cuda.synchronize()
    dist.barrier()
    times = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(W, src=0)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t))
    res["bcast_bytes"] = 8.0 * npad * npad
    res["bcast_s"] = min(times)
    res["bcast_gbs"] = res["bcast_bytes"] / res["bcast_s"] / 1e9
    chk = W[:64, :64].sum()
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    res["replicas_identical"] = bool(all(float(x) == float(lst[0]) for x in lst))
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
