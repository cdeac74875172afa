"""Measure the FP64 roofline denominators MEASURED_PEAKS.json does not have:
FP64 FMA pipe, FP64 DMMA tensor pipe, both interleaved, and exp() throughput.
Register-resident loops (socks_probe_fp64), CUDA-event timing, best of 5.
Usage (on a B200): python tools/probe_fp64.py [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from socks_b200 import _device as dv, _lib


def probe(mode, iters=20000, blocks=148 * 8):
    out = dv.empty(blocks * 256)
    st = dv.stream()
    _lib.call("socks_probe_fp64", mode, 200, dv.ptr(out), blocks, st)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("socks_probe_fp64", mode, iters, dv.ptr(out), blocks, st)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    threads, warps = blocks * 256, blocks * 8
    if mode == 0:
        return {"fp64_fma_tflops": threads * iters * 8 * 2 / best / 1e12}
    if mode == 1:
        return {"fp64_dmma_tflops": warps * iters * 8 * 512 / best / 1e12}
    if mode == 2:
        return {"mixed_fma_tflops": threads * iters * 4 * 2 / best / 1e12,
                "mixed_dmma_tflops": warps * iters * 4 * 512 / best / 1e12}
    if mode == 3:
        return {"exp_per_s": threads * iters * 4 / best}
    if mode == 4:
        return {"fp64_fma_distinct_operands_tflops": threads * iters * 8 * 2 / best / 1e12}
    if mode == 6:  # m16n8k32: 16*8*32*2 ops per warp instruction
        return {"legacy_mma_sync_int8_tops": warps * iters * 8 * 16 * 8 * 32 * 2 / best / 1e12}
    return {"fp64_fma_mul_crossfed_instr_per_s": threads * iters * 16 / best}


def measure():
    res = {}
    for mode in range(7):
        res.update(probe(mode, iters=4000 if mode == 3 else 20000))
    return res


if __name__ == "__main__":
    r = measure()
    r["gpu"] = torch.cuda.get_device_name(0)
    print(json.dumps(r))
    if len(sys.argv) > 1:
        json.dump(r, open(sys.argv[1], "w"), indent=1)
