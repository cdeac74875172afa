// Column reductions over a row-major matrix: the HBM-bound GEMV passes of the hot path.
//   part[z][s][j] = sum_{i in chunk s} V[z][i] * f(Bm[i][j]),   f = identity or square
// `np.einsum("i,ij->j", vf[t+1], beta)` (gym_socks/algorithms/reach/kernel_sr.py:62) is one such
// pass per time step; `normalize` (gym_socks/utils/__init__.py:22) is the pass with V = 1.
//
// Layout/roofline: each thread owns 2 adjacent columns (16-byte loads; a warp reads 512 contiguous
// bytes of a row), rows unrolled by 8 so every thread keeps 128 B in flight; the row range is split
// over gridDim.y CTAs and the partials are summed in a fixed order by the caller's finish kernel
// (deterministic: no atomics).  Algorithmic bytes = 8 n m per pass.
#pragma once
#include "common.cuh"

namespace socks {

constexpr int CR_THREADS = 128;
constexpr int CR_COLS = CR_THREADS * 2;
constexpr int CR_MAX_SPLITS = 64;

// tri != 0: Bm is lower triangular by 128-tiles; column j only reads rows >= 128*floor(j/128).
// WITH_SUM: the same pass also accumulates the UNWEIGHTED column sums (V = 1) into the partials of "row vector" z = 1
// (gridDim.z must be 1): the normaliser of kernel_sr.py:45 rides on the first GEMV pass instead of costing its own.
template <bool SQUARE, bool WITH_SUM = false>
__global__ void __launch_bounds__(CR_THREADS)
    colreduce_kernel(const double* __restrict__ Bm, int64_t n, int64_t m, int64_t ldb, const double* __restrict__ V,
                     int64_t ldv, double* __restrict__ part, int64_t ldpart, int tri, int vec_ok) {
    const int64_t j0 = ((int64_t)blockIdx.x * CR_THREADS + threadIdx.x) * 2;
    if (j0 >= m) return;
    const int splits = gridDim.y;
    const int64_t row_begin = tri ? (j0 / TILE) * TILE : 0;
    const int64_t len = (n - row_begin + splits - 1) / splits;
    const int64_t r0 = row_begin + (int64_t)blockIdx.y * len;
    const int64_t r1 = min(n, r0 + len);
    const double* v = V ? V + (int64_t)blockIdx.z * ldv : nullptr;
    const bool has1 = (j0 + 1 < m);
    const bool pair = vec_ok && has1;
    const double* p = Bm + j0;
    double a0 = 0.0, a1 = 0.0, s0 = 0.0, s1 = 0.0;
    int64_t i = r0;
    for (; i + 8 <= r1; i += 8) {
        double2 x[8];
        double vv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (pair) {
                x[u] = *reinterpret_cast<const double2*>(p + (i + u) * ldb);
            } else {
                x[u].x = p[(i + u) * ldb];
                x[u].y = has1 ? p[(i + u) * ldb + 1] : 0.0;
            }
            vv[u] = v ? __ldg(v + i + u) : 1.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (SQUARE) {
                a0 = fma(x[u].x * x[u].x, vv[u], a0);
                a1 = fma(x[u].y * x[u].y, vv[u], a1);
            } else {
                a0 = fma(x[u].x, vv[u], a0);
                a1 = fma(x[u].y, vv[u], a1);
            }
            if (WITH_SUM) {  // same association as the V = 1 pass (fma(x, 1, s) == s + x)
                s0 += x[u].x;
                s1 += x[u].y;
            }
        }
    }
    for (; i < r1; ++i) {
        const double x0 = p[i * ldb], x1 = has1 ? p[i * ldb + 1] : 0.0;
        const double vv = v ? __ldg(v + i) : 1.0;
        if (SQUARE) {
            a0 = fma(x0 * x0, vv, a0);
            a1 = fma(x1 * x1, vv, a1);
        } else {
            a0 = fma(x0, vv, a0);
            a1 = fma(x1, vv, a1);
        }
        if (WITH_SUM) {
            s0 += x0;
            s1 += x1;
        }
    }
    double* q = part + ((int64_t)blockIdx.z * splits + blockIdx.y) * ldpart + j0;
    q[0] = a0;
    if (has1) q[1] = a1;
    if (WITH_SUM) {
        double* qs = part + ((int64_t)splits + blockIdx.y) * ldpart + j0;
        qs[0] = s0;
        if (has1) qs[1] = s1;
    }
}

// y[z][j] = sum_s part[z][s][j]
static __global__ void colreduce_finish_kernel(const double* __restrict__ part, int64_t ldpart, int splits, int64_t m,
                                        double* __restrict__ y, int64_t ldy) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double* p = part + (int64_t)blockIdx.y * splits * ldpart + j;
    double a = 0.0;
    for (int s = 0; s < splits; ++s) a += p[(int64_t)s * ldpart];
    y[(int64_t)blockIdx.y * ldy + j] = a;
}

inline int colreduce_splits(int64_t n, int64_t m) {
    int64_t groups = ceil_div(m, CR_COLS);
    int64_t want = ceil_div(148 * 8, groups);
    int64_t cap = n / 64 > 1 ? n / 64 : 1;  // at least 64 rows per chunk
    if (want > cap) want = cap;
    if (want > CR_MAX_SPLITS) want = CR_MAX_SPLITS;
    if (want < 1) want = 1;
    return (int)want;
}

inline int64_t colreduce_ldpart(int64_t m) { return round_up(m, 2); }

// Launch part = reduce(...). part must hold R * CR_MAX_SPLITS * ldpart doubles. Returns splits used.
template <bool SQUARE, bool WITH_SUM = false>
inline int colreduce_launch(const double* Bm, int64_t n, int64_t m, int64_t ldb, const double* V, int64_t ldv, int R,
                            double* part, int tri, cudaStream_t st) {
    const int splits = colreduce_splits(n, m);
    const int vec_ok = (ldb % 2 == 0) && ((reinterpret_cast<uintptr_t>(Bm) & 15) == 0);
    dim3 grid((unsigned)ceil_div(m, CR_COLS), (unsigned)splits, (unsigned)R);
    colreduce_kernel<SQUARE, WITH_SUM><<<grid, CR_THREADS, 0, st>>>(Bm, n, m, ldb, V, ldv, part, colreduce_ldpart(m), tri,
                                                                   vec_ok);
    return splits;
}

}  // namespace socks
