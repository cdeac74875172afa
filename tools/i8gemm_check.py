"""Exactness check of the tcgen05 int8 GEMM building block against an integer reference.  Usage: python tools/i8gemm_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from socks_b200 import _device as dv, _lib

torch.manual_seed(0)
for (M, N, K) in ((128, 64, 32), (128, 64, 128), (256, 128, 512), (1024, 512, 4096)):
    A = torch.randint(-128, 128, (M, K), dtype=torch.int8, device="cuda")
    B = torch.randint(-128, 128, (N, K), dtype=torch.int8, device="cuda")
    want = (A.double() @ B.double().T).to(torch.int32)
    for variant in (0, 1, 2):  # 2: A operand written to tensor memory by tcgen05.st (TS form)
        C = torch.full((M, N), -7, dtype=torch.int32, device="cuda")
        _lib.call("socks_dev_i8gemm_check", A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, variant, dv.stream())
        torch.cuda.synchronize()
        bad = int((C != want).sum())
        print(f"M={M} N={N} K={K} variant={variant}: mismatches {bad} of {M * N}", flush=True)
        if bad and M == 128 and K == 32:
            print(" got ", C[:2, :8].tolist(), "\n want", want[:2, :8].tolist())
