// Register-resident FP64 throughput probes: the roofline denominators MEASURED_PEAKS.json lacks
// (FP64 FMA pipe, FP64 DMMA tensor pipe, both together, and the fp64 exp).  Used by tools/ and bench.py.
#include "common.cuh"

namespace socks {

// mode 0: 8 independent DFMA chains per thread              -> 16 flop per inner iteration per thread
// mode 1: 8 independent DMMA.8x8x4 accumulators per warp    -> 8 * 512 flop per inner iteration per warp
// mode 2: 4 DFMA chains + 4 DMMA accumulators interleaved   -> 8 flop per thread + 4 * 512 per warp
// mode 3: 4 independent exp() per thread per iteration
__global__ void __launch_bounds__(256) probe_fp64_kernel(int mode, int iters, double* __restrict__ out) {
    const double seed = 1.0 + 1e-9 * (threadIdx.x + 1);
    double a[8], c0[8], c1[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { a[u] = seed + u * 1e-3; c0[u] = 0.0; c1[u] = 0.0; }
    const double m = 0.999999, b = 1e-7;
    if (mode == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] = fma(a[u], m, b);
        }
    } else if (mode == 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) dmma884(c0[u], c1[u], m, b);
        }
    } else if (mode == 2) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a[u] = fma(a[u], m, b);
                dmma884(c0[u], c1[u], m, b);
            }
        }
    } else if (mode == 3) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = exp(-a[u]) + 0.5;
        }
    } else if (mode == 4) {  // 8 DFMA chains with three DISTINCT register operands each (operand-bandwidth probe)
        double bb[8], cc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { bb[u] = m + u * 1e-9 * seed; cc[u] = b * (u + 1) * seed; }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] = fma(a[u], bb[u], cc[u]);
        }
    } else {  // mode 5: chains cross-feeding each other: a[u] = fma(a[u], a[(u+1)%8], a[(u+3)%8]) (all operands vary)
        for (int it = 0; it < iters; ++it) {
            double n8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) n8[u] = fma(a[u], a[(u + 1) & 7], a[(u + 3) & 7]) * 0.25;
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] = n8[u];
        }
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += a[u] + c0[u] + c1[u];
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 6: legacy warp-level int8 tensor op (mma.sync.m16n8k32.s8 -> SASS IMMA): 8 independent accumulators.
// Feasibility probe for an Ozaki-scheme fp64 GEMM on the integer tensor cores (round-2 idea).
__global__ void __launch_bounds__(256) probe_imma_kernel(int iters, int* __restrict__ out) {
    int acc[8][4];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[u][e] = 0;
    const unsigned a0 = 0x01020304u * (threadIdx.x + 1), a1 = a0 ^ 0x11111111u, a2 = a0 + 7, a3 = a1 + 3;
    const unsigned b0 = 0x01010101u + threadIdx.x, b1 = b0 * 3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            asm volatile(
                "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                : "+r"(acc[u][0]), "+r"(acc[u][1]), "+r"(acc[u][2]), "+r"(acc[u][3])
                : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int s = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += acc[u][0] + acc[u][1] + acc[u][2] + acc[u][3];
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace socks

extern "C" int socks_probe_fp64(int mode, int iters, double* out_dev, int blocks, socks_stream_t stream) {
    SOCKS_CHECK_ARG(mode >= 0 && mode <= 6, 1);
    SOCKS_CHECK_ARG(iters >= 1, 2);
    SOCKS_CHECK_ARG(out_dev != nullptr, 3);
    SOCKS_CHECK_ARG(blocks >= 1, 4);
    if (mode == 6) {
        socks::probe_imma_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, reinterpret_cast<int*>(out_dev));
        SOCKS_CHECK_LAUNCH();
        return SOCKS_OK;
    }
    socks::probe_fp64_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(mode, iters, out_dev);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}
