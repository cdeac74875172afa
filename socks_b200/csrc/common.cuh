// Shared helpers for libsocksb200 (sm_100a, fp64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/socks_b200.h"
#include "../../include/socks_b200_dev.h"

namespace socks {

constexpr int TILE = SOCKS_TILE;
constexpr int SOCKS_MAX_DEVICES = 64;  // per-device one-time kernel attribute flags

void set_error(const char* fmt, ...);

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

#define SOCKS_CHECK_ARG(cond, k)                                                    \
    do {                                                                            \
        if (!(cond)) {                                                              \
            socks::set_error("%s: invalid argument %d (%s)", __func__, (k), #cond); \
            return SOCKS_EINVAL(k);                                                 \
        }                                                                           \
    } while (0)

#define SOCKS_CHECK_LAUNCH()                                                                   \
    do {                                                                                       \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess) {                                                              \
            socks::set_error("%s: CUDA launch failed: %s", __func__, cudaGetErrorString(e__)); \
            return SOCKS_ECUDA;                                                                \
        }                                                                                      \
    } while (0)

#define SOCKS_CUDA(call)                                                                  \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            socks::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__)); \
            return SOCKS_ECUDA;                                                           \
        }                                                                                 \
    } while (0)

// D(8x8) += A(8x4, row) * B(4x8, col), fp64 tensor core (SASS: DMMA.8x8x4).
// lane l holds a = A[l/4][l%4], b = B[l%4][l/4], c0,c1 = C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// Box-membership of one point against tubes[t][which] (which: 0 = constraint, 1 = target).
// Closed box, `>= low` and `<= high` (gym_socks/utils/__init__.py:42-47).  NaN coordinates fail.
__device__ __forceinline__ bool in_box(const double* __restrict__ p, const double* __restrict__ tube_t,
                                       int which, int d) {
    const double* lo = tube_t + (2 * which) * d;
    const double* hi = lo + d;
    bool ok = true;
    for (int k = 0; k < d; ++k) ok = ok && (p[k] >= lo[k]) && (p[k] <= hi[k]);
    return ok;
}

// One backward-recursion step (gym_socks/algorithms/reach/common.py:36-39 and :67-69), written
// with the same multiply-by-indicator arithmetic so NaN propagates exactly like numpy's
// `in_t + (in_k & ~in_t) * V` and `in_k * V`.
__device__ __forceinline__ double sr_step(const double* __restrict__ p, const double* __restrict__ tube_t,
                                          int d, int problem, double V) {
    bool in_k = in_box(p, tube_t, 0, d);
    if (problem == SOCKS_PROBLEM_FHT) {
        bool in_t = in_box(p, tube_t, 1, d);
        return (in_t ? 1.0 : 0.0) + ((in_k && !in_t) ? 1.0 : 0.0) * V;
    }
    return (in_k ? 1.0 : 0.0) * V;
}

}  // namespace socks
