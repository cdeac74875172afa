"""ncu target: the integer tensor-core GEMM on one trailing-update shape of socks_potrf_inv (N = 20k: rows x 1024 panel).
Usage: python tools/oz_gemm_profile_target.py [m] [K]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from socks_b200 import _device as dv, _lib

lib = _lib.load()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 10112
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(m, K, dtype=torch.float64, device="cuda", generator=g)
C = torch.randn(m, m, dtype=torch.float64, device="cuda", generator=g)
wk = dv.work(lib.socks_dev_ozaki_gemm_work_bytes(m, m, K) + 2048)
base = (wk.data_ptr() + 1023) // 1024 * 1024
for _ in range(3):
    _lib.call("socks_dev_ozaki_gemm_nt", A.data_ptr(), K, A.data_ptr(), K, C.data_ptr(), m, m, m, K, 1, 1, base, dv.stream())
torch.cuda.synchronize()
print("ok")
