// Integer tensor-core (tcgen05.mma kind::i8, TMEM accumulators) building blocks.
// Phase-1 check kernel: C (M x N, int32) = A (M x K, int8, K contiguous) * B (N x K, int8, K contiguous)^T, one CTA per
// 128 x 64 tile, one MMA (K = 32) per staged K-block.  Validates the descriptor / layout conventions of umma.cuh against
// an exact integer reference (tools/i8gemm_check.py) before anything is built on them: variant 0 = the convention used
// everywhere (LBO 128, SBO 256), 1 = the swapped one (negative control: every entry wrong), 2 = A operand written to
// tensor memory by tcgen05.st and read by the TS form of the MMA (what sr_predict_i8_kernel does).
#include <algorithm>
#include <atomic>
#include <cmath>

#include "umma.cuh"

namespace socks {

constexpr int I8_BM = 128, I8_BN = 64, I8_KB = 32;

// Core-matrix layout of an (R rows x 32 bytes) K-major operand tile: [r / 8][k-chunk c (16 B)][r % 8][16 B].
__device__ __forceinline__ int core_offset(int r, int c) { return (r >> 3) * 256 + c * 128 + (r & 7) * 16; }

__global__ void __launch_bounds__(128)
    i8gemm_check_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ C, int M, int N,
                        int K, int variant) {
    __shared__ __align__(128) uint8_t sA[I8_BM * I8_KB];
    __shared__ __align__(128) uint8_t sB[I8_BN * I8_KB];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.y * I8_BM, n0 = blockIdx.x * I8_BN;
    if (warp == 0) umma::tmem_alloc(&tmem_base, 128);
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::mbar_fence_init();
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_base;
    const uint32_t idesc = umma::idesc_i8(I8_BM, I8_BN, 1, 1);
    const uint32_t lbo = variant == 1 ? 256 : 128, sbo = variant == 1 ? 128 : 256;
    uint32_t phase = 0;
    for (int k0 = 0; k0 < K; k0 += I8_KB) {
        if (variant == 2) {  // A operand through tensor memory: thread = row, its 32 K-bytes -> 8 columns at column 64
            const int4* src = reinterpret_cast<const int4*>(A + (int64_t)(m0 + tid) * K + k0);
            const int4 lo = src[0], hi = src[1];
            const uint32_t w[8] = {(uint32_t)lo.x, (uint32_t)lo.y, (uint32_t)lo.z, (uint32_t)lo.w,
                                   (uint32_t)hi.x, (uint32_t)hi.y, (uint32_t)hi.z, (uint32_t)hi.w};
            umma::tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 64, w);
            umma::tmem_st_wait();
            umma::tc_fence_before();
        } else {
            for (int ch = tid; ch < I8_BM * 2; ch += 128) {
                const int r = ch >> 1, c = ch & 1;
                *reinterpret_cast<int4*>(sA + core_offset(r, c)) =
                    *reinterpret_cast<const int4*>(A + (int64_t)(m0 + r) * K + k0 + c * 16);
            }
        }
        for (int ch = tid; ch < I8_BN * 2; ch += 128) {
            const int r = ch >> 1, c = ch & 1;
            *reinterpret_cast<int4*>(sB + core_offset(r, c)) =
                *reinterpret_cast<const int4*>(B + (int64_t)(n0 + r) * K + k0 + c * 16);
        }
        umma::fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
        __syncthreads();
        if (tid == 0) {
            umma::tc_fence_after();
            const uint64_t da = umma::smem_desc_kmajor_noswizzle(umma::smem_u32(sA), lbo, sbo);
            const uint64_t db = umma::smem_desc_kmajor_noswizzle(umma::smem_u32(sB), lbo, sbo);
            if (variant == 2) umma::mma_i8_ts(tmem, tmem + 64, db, idesc, k0 > 0 ? 1u : 0u);
            else umma::mma_i8(tmem, da, db, idesc, k0 > 0 ? 1u : 0u);
            umma::mma_commit(&bar);
        }
        umma::mbar_wait(&bar, phase);  // the MMA has consumed this K-block: shared memory may be overwritten
        phase ^= 1;
    }
    umma::tc_fence_after();
    for (int c0 = 0; c0 < I8_BN; c0 += 16) {
        uint32_t v[16];
        umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) C[(int64_t)(m0 + tid) * N + n0 + c0 + i] = (int32_t)v[i];
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 128);
}

}  // namespace socks

using namespace socks;

extern "C" int socks_dev_i8gemm_check(const void* A, const void* B, void* C, int M, int N, int K, int variant,
                                      socks_stream_t stream) {
    SOCKS_CHECK_ARG(A && B && C, 1);
    SOCKS_CHECK_ARG(M > 0 && M % I8_BM == 0, 4);
    SOCKS_CHECK_ARG(N > 0 && N % I8_BN == 0, 5);
    SOCKS_CHECK_ARG(K > 0 && K % I8_KB == 0, 6);
    i8gemm_check_kernel<<<dim3(N / I8_BN, M / I8_BM), 128, 0, (cudaStream_t)stream>>>(
        (const int8_t*)A, (const int8_t*)B, (int32_t*)C, M, N, K, variant);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

// =====================================================================================================================
// KernelMaximalSR.predict on the integer tensor cores (Ozaki splitting; gym_socks/algorithms/reach/kernel_sr_max.py:65-72).
//
// The contraction  Cm[j][c] = sum_i K(x_i, p_j) * Bm[i][c],  Bm[i][(r,a)] = coef_r[i] CUA[i][a]  (1.5e13 flop at C2) only
// feeds ratios num/den of sums of NON-NEGATIVE terms that must be good to ~1e-9, so it does not need the FP64 pipe: both
// operands are split error-free into S unsigned 7-bit digits,
//     K_ij = sum_s a_s 2^(-7(s+1)),   Bm_ic / 2^(e_r) = sum_t b_t 2^(-7(t+1))      (e_r: exponent of max_i coef_r[i]),
// and every digit product A_s^T B_t is an EXACT int8 x int8 -> int32 GEMM on tcgen05.mma kind::i8.  Products with
// s + t = g share one TMEM accumulator; groups g >= S are dropped (below 2^(-7(S+2)) K 127^2 of the scale); the
// epilogue merges the group sums in exact 64-bit integer arithmetic, converts once and applies 2^(e_r).  The contraction
// runs on the two-pass 128 x 128 kernel of the next section (oz_gemm_nt_kernel<S, 7>; this section holds the operand
// producers, which write both operands directly in the shared-memory stage layout, and the host side).
// Points whose normaliser is too small for the fixed-point operand (far from every sample) are listed for the FP64 path.
namespace socks {

constexpr int OZ_BM = 128, OZ_KB = 32;
constexpr int OZ_A_BYTES = OZ_BM * OZ_KB;  // one digit of a 128-row tile, one K-block
constexpr int OZ_THREADS = 192;

template <int S>
__device__ __forceinline__ void oz_digits(double v01, uint32_t (&dig)[S]) {
    // v01 in [0, 1]: q = round(v 2^(7S)); leading digit unmasked (v = 1 gives 128), the others 7 bits
    const long long q = __double2ll_rn(v01 * (double)(1ll << (7 * S)));
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const unsigned d = (unsigned)(q >> (7 * (S - 1 - s)));
        dig[s] = (s == 0) ? (d & 0xFFu) : (d & 0x7Fu);
    }
}

// ---- operand producers: both write the shared-memory stage layout directly -----------------------------------------
// Aop[ptile][kb][s][core layout of 128 points x 32 samples]: digits of K(x_i, p_j) (direct form, CUDA exp).
template <int D, int S>
__global__ void __launch_bounds__(128)
    oz_slice_points_kernel(const double* __restrict__ X, int64_t n, const double* __restrict__ pts, int64_t m, double kmul,
                           uint8_t* __restrict__ Aop, int kblocks) {
    __shared__ double sx[512 * D];
    const int r = threadIdx.x;
    const int64_t ptile = blockIdx.y, j = ptile * OZ_BM + r;
    const int kbase = blockIdx.x * 16;
    for (int e = threadIdx.x; e < 512 * D; e += 128) {
        const int64_t i = (int64_t)kbase * 32 + e / D;
        sx[e] = (i < n) ? X[i * D + e % D] : 0.0;
    }
    __syncthreads();
    double p[D];
#pragma unroll
    for (int k = 0; k < D; ++k) p[k] = pts[min(j, m - 1) * D + k];
    const bool live_j = j < m;
    for (int kbl = 0; kbl < 16; ++kbl) {
        const int kb = kbase + kbl;
        if (kb >= kblocks) break;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t words[S][4];
#pragma unroll
            for (int s = 0; s < S; ++s)
#pragma unroll
                for (int q = 0; q < 4; ++q) words[s][q] = 0;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int li = kbl * 32 + c * 16 + e;
                const int64_t i = (int64_t)kbase * 32 + li;
                double d2 = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double t = p[k] - sx[li * D + k];
                    d2 = fma(t, t, d2);
                }
                const double kv = (live_j && i < n) ? exp(d2 * kmul) : 0.0;
                uint32_t dig[S];
                oz_digits<S>(kv, dig);
#pragma unroll
                for (int s = 0; s < S; ++s) words[s][e >> 2] |= dig[s] << (8 * (e & 3));
            }
            uint8_t* dst = Aop + (((int64_t)ptile * kblocks + kb) * S) * OZ_A_BYTES + core_offset(r, c);
#pragma unroll
            for (int s = 0; s < S; ++s)
                *reinterpret_cast<uint4*>(dst + (int64_t)s * OZ_A_BYTES) =
                    make_uint4(words[s][0], words[s][1], words[s][2], words[s][3]);
        }
    }
}

// rowscale[r] = 2^(e_r) >= max_i coef_r[i], inv[r] = 2^(-e_r); coef_0 = w, coef_r = vf[r] * w (srmax_pack's rows)
__global__ void oz_rowscale_kernel(const double* __restrict__ w, const double* __restrict__ vf, int64_t ldvf, int64_t n,
                                   double* __restrict__ rowscale, double* __restrict__ inv) {
    __shared__ double sm[256];
    const int r = blockIdx.x;
    double mx = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) mx = fmax(mx, r == 0 ? w[i] : vf[(int64_t)r * ldvf + i] * w[i]);
    sm[threadIdx.x] = mx;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + off]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double m0 = sm[0];
        const int e = (m0 > 0.0 && m0 < INFINITY) ? ilogb(m0) + 1 : 0;
        rowscale[r] = ldexp(1.0, e);
        inv[r] = ldexp(1.0, -e);
    }
}

__global__ void oz_colscale_kernel(const double* __restrict__ rowscale, int P, int T, int64_t ld, double* __restrict__ colscale) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ld) return;
    colscale[c] = (c < (int64_t)T * P) ? rowscale[c / P] : 0.0;
}

// Bop[ctile][kb][t][core layout of 128 columns x 32 samples]: digits of coef_r[i] CUA[i][a] / 2^(e_r).
template <int S>
__global__ void __launch_bounds__(128)
    oz_slice_coef_kernel(const double* __restrict__ CUA, int64_t n, int P, const double* __restrict__ w,
                         const double* __restrict__ vf, int64_t ldvf, int T, const double* __restrict__ inv,
                         uint8_t* __restrict__ Bop, int kblocks) {
    const int col = threadIdx.x;
    const int64_t ctile = blockIdx.y, c = ctile * OZ_BM + col;
    const bool live_c = c < (int64_t)T * P;
    const int r = live_c ? (int)(c / P) : 0, a = live_c ? (int)(c % P) : 0;
    const double sc = inv[r];
    const int kbase = blockIdx.x * 16;
    for (int kbl = 0; kbl < 16; ++kbl) {
        const int kb = kbase + kbl;
        if (kb >= kblocks) break;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            uint32_t words[S][4];
#pragma unroll
            for (int s = 0; s < S; ++s)
#pragma unroll
                for (int q = 0; q < 4; ++q) words[s][q] = 0;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int64_t i = (int64_t)kb * 32 + ch * 16 + e;
                double v = 0.0;
                if (live_c && i < n) {
                    const double coef = (r == 0) ? w[i] : vf[(int64_t)r * ldvf + i] * w[i];
                    v = fmin(coef * CUA[i * P + a] * sc, 1.0);
                }
                uint32_t dig[S];
                oz_digits<S>(v, dig);
#pragma unroll
                for (int s = 0; s < S; ++s) words[s][e >> 2] |= dig[s] << (8 * (e & 3));
            }
            uint8_t* dst = Bop + (((int64_t)ctile * kblocks + kb) * S) * OZ_A_BYTES + core_offset(col, ch);
#pragma unroll
            for (int s = 0; s < S; ++s)
                *reinterpret_cast<uint4*>(dst + (int64_t)s * OZ_A_BYTES) =
                    make_uint4(words[s][0], words[s][1], words[s][2], words[s][3]);
        }
    }
}

// Points whose smallest normaliser den[a] = Cm[j][a] is below `tau` of the operand scale are queued for the FP64 path
// (list[0] = count, list[1..] = global point indices): there the fixed-point digits no longer carry ~1e-9 relative.
__global__ void oz_flag_kernel(const double* __restrict__ Cm, int64_t ldc, int64_t m, int P, const double* __restrict__ rowscale,
                               double tau, int64_t j0, int* __restrict__ list) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= m) return;
    double mn = INFINITY;
    for (int a = lane; a < P; a += 32) mn = fmin(mn, Cm[j * ldc + a]);
    for (int off = 16; off > 0; off >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    if (lane == 0 && !(mn >= tau * rowscale[0])) list[1 + atomicAdd(list, 1)] = (int)(j0 + j);
}

static inline int64_t pad_to(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

}  // namespace socks

// =====================================================================================================================
// FP64-equivalent GEMM on the integer tensor cores (Ozaki scheme) for the factorisation's trailing updates:
//     C (m x n) -= A (m x K) * B (n x K)^T,   A, B, C fp64 row-major.
// Each row of an operand is scaled by 2^(-e_row) (e_row = exponent of its largest magnitude) and written as S = 7
// BALANCED base-256 digits d in [-128, 127] (int8): v 2^(-e) 2^54 = sum_s d_s 256^(6-s), exactly.  Digit products are
// exact int8 x int8 -> int32 GEMMs (|sum| <= 7 K 2^14 < 2^31 for K <= 16384); products with s + t = g share a TMEM
// accumulator and carry the weight 2^(-12 - 8g); groups g >= 7 (< 2^(-68) K 2^14 of |a|_max |b|_max per entry, i.e.
// below the rounding error K u |a||b| of an fp64 GEMM for K <= 2^14) are dropped.  The result is normwise as accurate
// as DGEMM; 28 int8 GEMMs replace one fp64 GEMM.  Operand digits live in 128-row tiles of the shared-memory stage
// layout: Dop[tile][kb][s][128 rows x 32 B core-matrix layout].
namespace socks {

// (digit counts: 7 in the trailing updates, 8 in the merges of inv(L) and in W = M^T M -- see chol.cu)

__device__ __forceinline__ double og_pow2(int e) {  // 2^e, e in the normal range (folds to a constant)
    return __longlong_as_double((long long)(1023 + e) << 52);
}


// Structure masks (128-tile granularity, as everywhere in chol.cu: tiles beyond the block diagonal hold garbage):
//   OZ_TRI_K_LE_ROW  operand row r only has entries for k <  128 (floor(r/128) + 1)   (lower-triangular A: k <= row)
//   OZ_TRI_K_GE_ROW  operand row r only has entries for k >= 128 floor(r/128)         (B(k, j) = 0 for k < j)
enum OzTri : int { OZ_TRI_NONE = 0, OZ_TRI_K_LE_ROW = 1, OZ_TRI_K_GE_ROW = 2 };

__device__ __forceinline__ bool oz_in_range(int tri, int64_t r, int64_t k) {
    if (tri == OZ_TRI_K_LE_ROW) return k < (r / 128 + 1) * 128;
    if (tri == OZ_TRI_K_GE_ROW) return k >= (r / 128) * 128;
    return true;
}

// Operand element (r, k): P[r][k] (TRANS = false) or P[k][r] (TRANS = true: the operand is the transpose of a row-major
// K x rows matrix, i.e. the B of an "NN" product).
// scale[r] = 2^(e) with 2^(e-1) <= max_k |op(r, k)| < 2^e (1 for an all-zero row).
__global__ void oz_rowscale_rows_kernel(const double* __restrict__ P, int64_t rows, int64_t K, int64_t ld, int tri,
                                        const double* __restrict__ kscale, double* __restrict__ scale) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const double* p = P + r * ld;
    double mx = 0.0;
    for (int64_t k = lane; k < K; k += 32)
        if (oz_in_range(tri, r, k)) mx = fmax(mx, fabs(p[k]) * (kscale ? kscale[k] : 1.0));
    for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if (lane == 0) scale[r] = (mx > 0.0 && mx < INFINITY) ? ldexp(1.0, ilogb(mx) + 1) : 1.0;
}
// transposed operand: thread = operand row (= column of P, coalesced across threads), blockIdx.y = 256-sample K chunk;
// the maxima are combined as integers (non-negative doubles order like their bit patterns), then turned into scales.
__global__ void oz_colmax_kernel(const double* __restrict__ P, int64_t rows, int64_t K, int64_t ld, int tri,
                                 const double* __restrict__ kscale, unsigned long long* __restrict__ mxbits) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int64_t k0 = (int64_t)blockIdx.y * 256, k1 = min(K, k0 + 256);
    double mx = 0.0;
    for (int64_t k = k0; k < k1; ++k)
        if (oz_in_range(tri, r, k)) mx = fmax(mx, fabs(P[k * ld + r]) * (kscale ? kscale[k] : 1.0));
    if (mx > 0.0) atomicMax(mxbits + r, (unsigned long long)__double_as_longlong(mx));
}
__global__ void oz_scale_from_max_kernel(double* __restrict__ scale, int64_t rows) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double mx = scale[r];
    scale[r] = (mx > 0.0 && mx < INFINITY) ? ldexp(1.0, ilogb(mx) + 1) : 1.0;
}
// ---- balancing along K -------------------------------------------------------------------------------------------------
// The digit split truncates relative to an operand ROW's maximum.  In the merges of inv(L) the two operands have opposite
// magnitude profiles along K (columns of L decay like l_kk, rows of inv(L) grow like 1 / l_kk): terms a_ik b_jk of order one
// are products of a tiny and a huge factor, and the tiny factor is truncated away.  A power-of-two scale per k,
//     A B^T = (A D^-1) (B D)^T,   D_k = 2^round((log2 max_i |a_ik| - log2 max_j |b_jk|) / 2),
// equalises the K-profiles of the two operands exactly (the scales are exponent shifts) before their rows are scaled and
// split: on Gram matrices with cond up to 8e11 the 7-digit product is then as accurate as FP64.
// kmax[k] = max over operand rows r of |op(r, k)| (mask as in oz_in_range), combined as integer bit patterns.
template <bool TRANS>
__global__ void oz_kmax_kernel(const double* __restrict__ P, int64_t rows, int64_t K, int64_t ld, int tri,
                               unsigned long long* __restrict__ kmax) {
    if (TRANS) {  // op(r, k) = P[k][r]: one warp per k, lanes over r (coalesced)
        const int lane = threadIdx.x & 31;
        const int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        if (k >= K) return;
        double mx = 0.0;
        for (int64_t r = lane; r < rows; r += 32)
            if (oz_in_range(tri, r, k)) mx = fmax(mx, fabs(P[k * ld + r]));
        for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        if (lane == 0 && mx > 0.0) kmax[k] = (unsigned long long)__double_as_longlong(mx);
    } else {  // op(r, k) = P[r][k]: thread = k (coalesced), blockIdx.y = chunk of 256 operand rows
        const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (k >= K) return;
        const int64_t r0 = (int64_t)blockIdx.y * 256, r1 = min(rows, r0 + 256);
        double mx = 0.0;
        for (int64_t r = r0; r < r1; ++r)
            if (oz_in_range(tri, r, k)) mx = fmax(mx, fabs(P[r * ld + k]));
        if (mx > 0.0) atomicMax(kmax + k, (unsigned long long)__double_as_longlong(mx));
    }
}
// sa[k] = 2^-e_k (applied to A), sb[k] = 2^e_k (applied to B); in place over the two kmax arrays
__global__ void oz_kbalance_kernel(double* __restrict__ ka, double* __restrict__ kb, int64_t K) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const double a = ka[k], b = kb[k];
    int e = 0;
    if (a > 0.0 && b > 0.0 && a < INFINITY && b < INFINITY) e = (int)rint(0.5 * (double)(ilogb(a) - ilogb(b)));
    e = max(-500, min(500, e));
    ka[k] = ldexp(1.0, -e);
    kb[k] = ldexp(1.0, e);
}

// K-blocks of operand tile `rtile` that a clipped GEMM reads (the others are neither written nor read)
__device__ __forceinline__ bool oz_kb_live(int tri, int64_t rtile, int kb) {
    if (tri == OZ_TRI_K_LE_ROW) return kb < (rtile + 1) * (OZ_BM / OZ_KB);
    if (tri == OZ_TRI_K_GE_ROW) return kb >= rtile * (OZ_BM / OZ_KB);
    return true;
}

// Balanced base-256 digits of q (|q| < 2^(8S-2)): q = sum_s d_s 256^(S-1-s), d_s in [-128, 127] (d_0 in [-64, 64]).
template <int S>
__device__ __forceinline__ void oz_signed_digits(long long q, uint32_t (&dig)[S]) {
#pragma unroll
    for (int s = S - 1; s > 0; --s) {
        const int d = (int)(signed char)(q & 0xFF);
        dig[s] = (uint32_t)(d & 0xFF);
        q = (q - d) >> 8;  // exact: q - d is a multiple of 256
    }
    dig[0] = (uint32_t)((int)q & 0xFF);
}

// Dop[rtile][kb][s][core layout]: digits of op(row, k) / scale[row]; thread = operand row of the tile, OZ_SLICE_KB
// K-blocks per CTA (few, so that a 1024-wide panel still fills the machine).
constexpr int OZ_SLICE_KB = 2;
template <bool TRANS, int S>
__global__ void __launch_bounds__(128)
    oz_slice_rows_kernel(const double* __restrict__ P, int64_t rows, int64_t K, int64_t ld, int tri,
                         const double* __restrict__ scale, const double* __restrict__ kscale, uint8_t* __restrict__ Dop,
                         int kblocks) {
    const int r = threadIdx.x;
    const int64_t rtile = blockIdx.y, row = rtile * OZ_BM + r;
    const bool live = row < rows;
    const double inv = live ? (og_pow2(8 * S - 2) / scale[row]) : 0.0;  // exact power of two; |q| < 2^(8S-2): top digit in [-64, 64]
    const int kb0 = blockIdx.x * OZ_SLICE_KB;
#pragma unroll 1
    for (int kbl = 0; kbl < OZ_SLICE_KB; ++kbl) {
        const int kb = kb0 + kbl;
        if (kb >= kblocks) break;
        if (!oz_kb_live(tri, rtile, kb)) continue;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t words[S][4];
#pragma unroll
            for (int s = 0; s < S; ++s)
#pragma unroll
                for (int q = 0; q < 4; ++q) words[s][q] = 0;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int64_t k = (int64_t)kb * 32 + c * 16 + e;
                double v = 0.0;
                if (live && k < K && oz_in_range(tri, row, k))
                    v = (TRANS ? P[k * ld + row] : P[row * ld + k]) * (kscale ? kscale[k] : 1.0);  // exact: power of two
                uint32_t d[S];
                oz_signed_digits<S>(__double2ll_rn(v * inv), d);
#pragma unroll
                for (int s = 0; s < S; ++s) words[s][e >> 2] |= d[s] << (8 * (e & 3));
            }
            uint8_t* dst = Dop + (((int64_t)rtile * kblocks + kb) * S) * OZ_A_BYTES + core_offset(r, c);
#pragma unroll
            for (int s = 0; s < S; ++s)
                *reinterpret_cast<uint4*>(dst + (int64_t)s * OZ_A_BYTES) =
                    make_uint4(words[s][0], words[s][1], words[s][2], words[s][3]);
        }
    }
}

// Tile = 128 x 128 of C, computed in TWO passes over K.  The seven digit groups g = s + t need one int32 accumulator each,
// and 7 x 128 columns do not fit the 512 columns of tensor memory; 7 x 64 would, but a 128 x 64 x 32 MMA is bound by the
// shared-memory reads of its operands (48 clocks for 32 clocks of math; tools/probes/umma_rate.cu), while 128 x 128 x 32
// runs at the full rate.  So pass A accumulates groups 0..3 (10 MMAs per K-block, digit slices 0..3 of both operands) and
// pass B groups 4..6 (18 MMAs, all slices); each pass ends in its own epilogue.  Both passes are bound by L2 -> SM
// traffic at ~45 B/clk/SM (tools/probes/umma_pipe.cu): 88 KB per K-block and tile.
template <int S>
struct OgCfg {
    static constexpr int BN = 128;
    static constexpr int SLICE = OZ_A_BYTES;                                 // 128 rows x 32 B of one digit
    static constexpr int NS_A = 4, NS_B = S;                                 // digit slices a pass needs
    static constexpr int STAGE_A = 2 * NS_A * SLICE, STAGES_A = 5;           // 32 KB x 5
    static constexpr int STAGE_B = 2 * NS_B * SLICE;                         // 48 / 56 / 64 KB for S = 6 / 7 / 8
    static constexpr int STAGES_B = (192 * 1024) / STAGE_B;                  // 4 / 3 / 3
    static constexpr int RING_A = STAGES_A * STAGE_A, RING_B = STAGES_B * STAGE_B;
    static constexpr int RING_BYTES = RING_A > RING_B ? RING_A : RING_B;     // 192 / 168 / 192 KB
    static constexpr int NBAR = 2 * (STAGES_A + STAGES_B) + 3;
    static constexpr int SMEM_BYTES = RING_BYTES + NBAR * 8 + 16;
};

// int32 stays exact for og_kchunk K-blocks; longer products hand the accumulators over chunk by chunk (fp64 sum in C).
// Signed 8-bit digits: |d d'| <= 2^14, at most 7 pairs per group: 7 * 16384 * 2^14 < 2^31.  Unsigned 7-bit digits (top
// digit up to 128): at most 8 pairs: 8 * 10240 * 2^14 < 2^31.
template <int BITS, int S>
struct OgChunk {  // 8 signed digits: the last group has 8 pairs, 8 * 2^14 * 16384 would reach 2^31 exactly
    static constexpr int KBLOCKS = BITS == 8 ? (S == 8 ? 448 : 512) : 320;
};

// exact int64 -> double for |x| < 2^51 (one integer add, one DADD; I2F.F64 is several times slower)
__device__ __forceinline__ double og_i2d(long long x) {
    return __longlong_as_double(x + 0x4338000000000000LL) - 0x1.8p52;
}

// C[128 ptile + i][128 ctile + j] (op)= sa[i] sb[j] sum_g 2^(-12 - 8g) G_g[i][j];  mode: 0 C -= AB^T, 1 C = AB^T,
// 2 C = -AB^T.  tri_a / tri_b clip the K range per tile exactly like GEMM_A_K_LE_M / GEMM_B_K_GE_N of gemm_f64.cuh.
// Warps 0-3: epilogue (thread = row), warp 4: loader (bulk copies), warp 5: MMA issue.
//
// Template: S digit slices of BITS bits.  BITS = 8: balanced signed digits, group weight 2^(-12 - 8g) (the GEMM of the
// factorisation, S = 7).  BITS = 7: unsigned digits of values in [0, 1], group weight 2^(-14 - 7g), sa == nullptr (no row
// scale): the KernelMaximalSR contraction, S = 6..8.
template <int S, int BITS>
__global__ void __launch_bounds__(OZ_THREADS, 1)
    oz_gemm_nt_kernel(const uint8_t* __restrict__ Aop, const double* __restrict__ sa, const uint8_t* __restrict__ Bop,
                      const double* __restrict__ sb, double* __restrict__ C, int64_t ldc, int kblocks, int lower_only,
                      int mode, int tri_a, int tri_b) {
    using Cfg = OgCfg<S>;
    constexpr int SL = Cfg::SLICE, BN = Cfg::BN, OG_KCHUNK = OgChunk<BITS, S>::KBLOCKS;
    constexpr int W0 = (BITS == 8) ? 12 : 14, NGB = S - 4;  // weight of group 0 is 2^-W0; groups of pass B
    static_assert(S >= 6 && S <= 8 && (BITS == 7 || BITS == 8), "digit scheme");
    const int ctile = blockIdx.x, ptile = blockIdx.y;
    if (lower_only && ctile > ptile) return;
    int kb_lo = 0, kb_hi = kblocks;
    if (tri_a == 1) kb_hi = min(kb_hi, (ptile + 1) * (OZ_BM / OZ_KB));  // A(i, k) = 0 for k beyond row tile i
    if (tri_a == 2) kb_lo = max(kb_lo, ptile * (OZ_BM / OZ_KB));        // A(i, k) = 0 for k before row tile i
    if (tri_b) kb_lo = max(kb_lo, ctile * (BN / OZ_KB));           // B(k, j) = 0 for k before column tile j
    const int nk = max(kb_hi - kb_lo, 0);
    const int nchunks = (nk + OG_KCHUNK - 1) / OG_KCHUNK;
    extern __shared__ __align__(1024) uint8_t oz_smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(oz_smem + Cfg::RING_BYTES);
    uint64_t* full_a = bars;
    uint64_t* empty_a = full_a + Cfg::STAGES_A;
    uint64_t* full_b = empty_a + Cfg::STAGES_A;
    uint64_t* empty_b = full_b + Cfg::STAGES_B;
    uint64_t* acc_full = empty_b + Cfg::STAGES_B;
    uint64_t* acc_empty = acc_full + 1;
    uint64_t* ring_free = acc_empty + 1;  // all MMAs of pass A have read their stages: pass B may overwrite the ring
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);
    __shared__ double col_scale[BN];
    __shared__ double turn[4 * 32 * 17];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (nk == 0) {  // empty K range: the product is zero
        if (mode != 0 && tid < OZ_BM) {
            double* crow = C + ((int64_t)ptile * OZ_BM + tid) * ldc + (int64_t)ctile * BN;
            for (int i = 0; i < BN; ++i) crow[i] = 0.0;
        }
        return;
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    if (tid == 32) {
        for (int i = 0; i < 2 * (Cfg::STAGES_A + Cfg::STAGES_B); ++i) umma::mbar_init(bars + i, 1);
        umma::mbar_init(acc_full, 1);
        umma::mbar_init(acc_empty, 128);
        umma::mbar_init(ring_free, 1);
        umma::mbar_fence_init();
    }
    if (tid < BN) col_scale[tid] = sb[(int64_t)ctile * BN + tid];
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 4) {
        if (lane == 0) {
            const uint8_t* a_src = Aop + ((int64_t)ptile * kblocks + kb_lo) * (S * SL);
            const uint8_t* b_src = Bop + ((int64_t)ctile * kblocks + kb_lo) * (S * SL);
            for (int i = 0; i < nk; ++i) {  // pass A: slices 0..3
                const int st = i % Cfg::STAGES_A, round = i / Cfg::STAGES_A;
                if (round > 0) umma::mbar_wait(empty_a + st, (uint32_t)(round - 1) & 1u);
                uint8_t* dst = oz_smem + st * Cfg::STAGE_A;
                umma::mbar_arrive_expect_tx(full_a + st, Cfg::STAGE_A);
                umma::bulk_g2s(dst, a_src + (int64_t)i * (S * SL), Cfg::NS_A * SL, full_a + st);
                umma::bulk_g2s(dst + Cfg::NS_A * SL, b_src + (int64_t)i * (S * SL), Cfg::NS_A * SL, full_a + st);
            }
            umma::mbar_wait(ring_free, 0);
            for (int i = 0; i < nk; ++i) {  // pass B: all slices
                const int st = i % Cfg::STAGES_B, round = i / Cfg::STAGES_B;
                if (round > 0) umma::mbar_wait(empty_b + st, (uint32_t)(round - 1) & 1u);
                uint8_t* dst = oz_smem + st * Cfg::STAGE_B;
                umma::mbar_arrive_expect_tx(full_b + st, Cfg::STAGE_B);
                umma::bulk_g2s(dst, a_src + (int64_t)i * (S * SL), S * SL, full_b + st);
                umma::bulk_g2s(dst + S * SL, b_src + (int64_t)i * (S * SL), S * SL, full_b + st);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            const uint32_t idesc = umma::idesc_i8(OZ_BM, BN, BITS == 8, BITS == 8);  // signed 8-bit or unsigned 7-bit digits
            uint32_t seg = 0;                                        // accumulator hand-overs so far
            for (int i0 = 0; i0 < nk; i0 += OG_KCHUNK, ++seg) {      // pass A: groups 0..3 -> columns 128 g
                if (seg > 0) {
                    umma::mbar_wait(acc_empty, (seg - 1) & 1u);
                    umma::tc_fence_after();
                }
                const int i1 = min(nk, i0 + OG_KCHUNK);
                for (int i = i0; i < i1; ++i) {
                    const int st = i % Cfg::STAGES_A, round = i / Cfg::STAGES_A;
                    umma::mbar_wait(full_a + st, (uint32_t)round & 1u);
                    umma::tc_fence_after();
                    const uint32_t sA = umma::smem_u32(oz_smem + st * Cfg::STAGE_A), sB = sA + Cfg::NS_A * SL;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const uint64_t da = umma::smem_desc_kmajor_noswizzle(sA + a * SL, 128, 256);
#pragma unroll
                        for (int b = 0; a + b < 4; ++b)
                            umma::mma_i8(tmem + (uint32_t)((a + b) * BN), da,
                                         umma::smem_desc_kmajor_noswizzle(sB + b * SL, 128, 256), idesc,
                                         (i > i0 || a > 0) ? 1u : 0u);
                    }
                    umma::mma_commit(empty_a + st);
                }
                if (i1 == nk) umma::mma_commit(ring_free);
                umma::mma_commit(acc_full);
            }
            for (int i0 = 0; i0 < nk; i0 += OG_KCHUNK, ++seg) {  // pass B: groups 4..6 -> columns 128 (g - 4)
                umma::mbar_wait(acc_empty, (seg - 1) & 1u);
                umma::tc_fence_after();
                const int i1 = min(nk, i0 + OG_KCHUNK);
                for (int i = i0; i < i1; ++i) {
                    const int st = i % Cfg::STAGES_B, round = i / Cfg::STAGES_B;
                    umma::mbar_wait(full_b + st, (uint32_t)round & 1u);
                    umma::tc_fence_after();
                    const uint32_t sA = umma::smem_u32(oz_smem + st * Cfg::STAGE_B), sB = sA + S * SL;
#pragma unroll
                    for (int a = 0; a < S; ++a) {
                        const uint64_t da = umma::smem_desc_kmajor_noswizzle(sA + a * SL, 128, 256);
#pragma unroll
                        for (int b = 0; b < S; ++b)
                            if (a + b >= 4 && a + b < S)
                                umma::mma_i8(tmem + (uint32_t)((a + b - 4) * BN), da,
                                             umma::smem_desc_kmajor_noswizzle(sB + b * SL, 128, 256), idesc,
                                             (i > i0 || a > 0) ? 1u : 0u);
                    }
                    umma::mma_commit(empty_b + st);
                }
                umma::mma_commit(acc_full);
            }
        }
    } else {
        const int64_t row = (int64_t)ptile * OZ_BM + tid;
        double* crow = C + row * ldc + (int64_t)ctile * BN;
        const double sr = ((mode == 1) ? 1.0 : -1.0) * (sa ? sa[row] : 1.0);
        if (mode == 0) {  // pull this thread's 1 KB of C towards the SM while the tensor cores work
#pragma unroll
            for (int i = 0; i < BN; i += 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(crow + i));
        }
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        double* tb = turn + warp * (32 * 17);
        double* cwarp = C + ((int64_t)ptile * OZ_BM + warp * 32) * ldc + (int64_t)ctile * BN;
        double nxt[16];
        if (mode == 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k) nxt[k] = __ldcg(cwarp + (lane & 15) + (int64_t)(2 * k + (lane >> 4)) * ldc);
        }
        for (uint32_t seg = 0; seg < 2u * (uint32_t)nchunks; ++seg) {
            const bool pass_b = seg >= (uint32_t)nchunks;
            const bool overwrite = (mode != 0) && seg == 0;
            const bool last_seg = seg + 1 == 2u * (uint32_t)nchunks;
            umma::mbar_wait(acc_full, seg & 1u);
            umma::tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                // this chunk's C values were requested one chunk ago (or before the pass); request the next ones now, so
                // that their latency hides behind the tensor-memory reads and the arithmetic of this chunk.  A lane only
                // ever reads elements it wrote itself, so a request may run ahead into the next pass.
                double cur[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) cur[k] = overwrite ? 0.0 : nxt[k];
                {
                    const int cn = (c0 + 16) & (BN - 1);
                    if (c0 + 16 < BN ? !overwrite : !last_seg) {
#pragma unroll
                        for (int k = 0; k < 16; ++k)
                            nxt[k] = __ldcg(cwarp + cn + (lane & 15) + (int64_t)(2 * k + (lane >> 4)) * ldc);
                    }
                }
                // sum_g 2^(-W0 - BITS g) G_g over the pass's groups, merged in exact 64-bit integer arithmetic first
                // (every partial stays below 2^51: |G_g| < 2^31, |G_0| <= K 2^14) and converted once
                double v[16];
                if (!pass_b) {
                    uint32_t u[4][16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) umma::tmem_ld16(trow + (uint32_t)(g * BN + c0), u[g]);
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const long long hi = ((long long)(int)u[0][i] << BITS) + (int)u[1][i];
                        const long long lo = ((long long)(int)u[2][i] << BITS) + (int)u[3][i];
                        v[i] = fma(og_i2d(hi), og_pow2(-W0 - BITS), og_i2d(lo) * og_pow2(-W0 - 3 * BITS));
                    }
                } else {
                    uint32_t u[NGB][16];
#pragma unroll
                    for (int g = 0; g < NGB; ++g) umma::tmem_ld16(trow + (uint32_t)(g * BN + c0), u[g]);
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if constexpr (NGB == 4) {  // two halves: one merged integer could pass 2^51
                            const long long hi = ((long long)(int)u[0][i] << BITS) + (int)u[1][i];
                            const long long lo = ((long long)(int)u[2][i] << BITS) + (int)u[3][i];
                            v[i] = fma(og_i2d(hi), og_pow2(-W0 - 5 * BITS), og_i2d(lo) * og_pow2(-W0 - 7 * BITS));
                        } else {
                            long long q = (int)u[0][i];
#pragma unroll
                            for (int g = 1; g < NGB; ++g) q = (q << BITS) + (int)u[g][i];
                            v[i] = og_i2d(q) * og_pow2(-W0 - BITS * (S - 1));
                        }
                    }
                }
                // tensor memory hands a thread one ROW; C wants 128 contiguous bytes per row and instruction, so the
                // 32 x 16 block of the warp is turned through shared memory (two rows of 16 doubles per instruction)
#pragma unroll
                for (int i = 0; i < 16; ++i) tb[lane * 17 + i] = sr * col_scale[c0 + i] * v[i];
                __syncwarp();
                double* cbase = cwarp + c0 + (lane & 15);
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int r = 2 * k + (lane >> 4);
                    cbase[(int64_t)r * ldc] = cur[k] + tb[r * 17 + (lane & 15)];
                }
                __syncwarp();
            }
            umma::tc_fence_before();
            umma::mbar_arrive(acc_empty);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

template <int S, int BITS>
static int og_configure() {
    static std::atomic<bool> done[SOCKS_MAX_DEVICES];
    int dev = 0;
    SOCKS_CUDA(cudaGetDevice(&dev));
    const bool tracked = dev >= 0 && dev < SOCKS_MAX_DEVICES;
    if (tracked && done[dev].load(std::memory_order_acquire)) return SOCKS_OK;
    SOCKS_CUDA(cudaFuncSetAttribute((oz_gemm_nt_kernel<S, BITS>), cudaFuncAttributeMaxDynamicSharedMemorySize, OgCfg<S>::SMEM_BYTES));
    if (tracked) done[dev].store(true, std::memory_order_release);
    return SOCKS_OK;
}

// Digits + row scales of a (rows x K) fp64 operand; rows are padded to whole 128-row tiles with zeros.
int64_t ozaki_operand_bytes(int64_t rows, int64_t K, int S) {
    const int64_t rt = pad_to(rows, OZ_BM) / OZ_BM, kb = pad_to(K, OZ_KB) / OZ_KB;
    return pad_to(rt * kb * S * OZ_A_BYTES, 1024) + pad_to(rt * OZ_BM * 8, 1024);
}
// trans == 0: operand(r, k) = P[r][k];  trans != 0: operand(r, k) = P[k][r].  tri: OzTri mask of the operand.
int ozaki_slice_operand(const double* P, int64_t rows, int64_t K, int64_t ld, int trans, int tri, uint8_t* buf, int S,
                        const double* kscale, cudaStream_t st) {
    const int64_t rt = pad_to(rows, OZ_BM) / OZ_BM;
    const int kb = (int)(pad_to(K, OZ_KB) / OZ_KB);
    double* scale = reinterpret_cast<double*>(buf + pad_to(rt * kb * S * OZ_A_BYTES, 1024));
    const dim3 gs((unsigned)ceil_div(kb, OZ_SLICE_KB), (unsigned)rt);
    if (trans) {
        cudaMemsetAsync(scale, 0, (size_t)rt * OZ_BM * 8, st);
        oz_colmax_kernel<<<dim3((unsigned)ceil_div(rows, 128), (unsigned)ceil_div(K, 256)), 128, 0, st>>>(
            P, rows, K, ld, tri, kscale, reinterpret_cast<unsigned long long*>(scale));
        oz_scale_from_max_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, st>>>(scale, rows);
        if (S == 8) oz_slice_rows_kernel<true, 8><<<gs, 128, 0, st>>>(P, rows, K, ld, tri, scale, kscale, buf, kb);
        else oz_slice_rows_kernel<true, 7><<<gs, 128, 0, st>>>(P, rows, K, ld, tri, scale, kscale, buf, kb);
    } else {
        oz_rowscale_rows_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, st>>>(P, rows, K, ld, tri, kscale, scale);
        if (S == 8) oz_slice_rows_kernel<false, 8><<<gs, 128, 0, st>>>(P, rows, K, ld, tri, scale, kscale, buf, kb);
        else oz_slice_rows_kernel<false, 7><<<gs, 128, 0, st>>>(P, rows, K, ld, tri, scale, kscale, buf, kb);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("ozaki_slice_operand: %s", cudaGetErrorString(e));
        return SOCKS_ECUDA;
    }
    return SOCKS_OK;
}

// Power-of-two scales along K that balance the two operands of one product (see oz_kmax_kernel): on return
// kwork[0 .. K) is the scale for operand A, kwork[K .. 2K) the one for operand B (pass them to ozaki_slice_operand).
int ozaki_kbalance(const double* PA, int64_t rows_a, int64_t lda, int trans_a, int tri_a, const double* PB, int64_t rows_b,
                   int64_t ldb, int trans_b, int tri_b, int64_t K, double* kwork, cudaStream_t st) {
    cudaMemsetAsync(kwork, 0, (size_t)(2 * K) * 8, st);
    unsigned long long* ka = reinterpret_cast<unsigned long long*>(kwork);
    unsigned long long* kb = ka + K;
    if (trans_a) oz_kmax_kernel<true><<<(unsigned)ceil_div(K, 8), 256, 0, st>>>(PA, rows_a, K, lda, tri_a, ka);
    else oz_kmax_kernel<false><<<dim3((unsigned)ceil_div(K, 128), (unsigned)ceil_div(rows_a, 256)), 128, 0, st>>>(PA, rows_a, K, lda, tri_a, ka);
    if (trans_b) oz_kmax_kernel<true><<<(unsigned)ceil_div(K, 8), 256, 0, st>>>(PB, rows_b, K, ldb, tri_b, kb);
    else oz_kmax_kernel<false><<<dim3((unsigned)ceil_div(K, 128), (unsigned)ceil_div(rows_b, 256)), 128, 0, st>>>(PB, rows_b, K, ldb, tri_b, kb);
    oz_kbalance_kernel<<<(unsigned)ceil_div(K, 256), 256, 0, st>>>(kwork, kwork + K, K);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("ozaki_kbalance: %s", cudaGetErrorString(e));
        return SOCKS_ECUDA;
    }
    return SOCKS_OK;
}

// C (m x n) (op)= A B^T from sliced operands (tile offsets in units of 128 operand rows); mode 0: -=, 1: =, 2: = -.
int ozaki_gemm_nt_sliced(const uint8_t* abuf, int64_t a_rows_total, int64_t a_tile0, const uint8_t* bbuf, int64_t b_rows_total,
                         int64_t b_tile0, int64_t K, double* C, int64_t ldc, int64_t m, int64_t n, int lower_only, int mode,
                         int tri_a, int tri_b, int S, cudaStream_t st) {
    int rc = (S == 8) ? og_configure<8, 8>() : og_configure<7, 8>();
    if (rc) return rc;
    const int kb = (int)(pad_to(K, OZ_KB) / OZ_KB);
    const int64_t art = pad_to(a_rows_total, OZ_BM) / OZ_BM, brt = pad_to(b_rows_total, OZ_BM) / OZ_BM;
    const double* sa = reinterpret_cast<const double*>(abuf + pad_to(art * kb * S * OZ_A_BYTES, 1024)) + a_tile0 * OZ_BM;
    const double* sb = reinterpret_cast<const double*>(bbuf + pad_to(brt * kb * S * OZ_A_BYTES, 1024)) + b_tile0 * OZ_BM;
    const uint8_t* aop = abuf + a_tile0 * (int64_t)kb * S * OZ_A_BYTES;
    const uint8_t* bop = bbuf + b_tile0 * (int64_t)kb * S * OZ_A_BYTES;
    const dim3 grid((unsigned)(n / 128), (unsigned)(m / OZ_BM));
    if (S == 8)
        oz_gemm_nt_kernel<8, 8><<<grid, OZ_THREADS, OgCfg<8>::SMEM_BYTES, st>>>(aop, sa, bop, sb, C, ldc, kb, lower_only, mode,
                                                                                tri_a, tri_b);
    else
        oz_gemm_nt_kernel<7, 8><<<grid, OZ_THREADS, OgCfg<7>::SMEM_BYTES, st>>>(aop, sa, bop, sb, C, ldc, kb, lower_only, mode,
                                                                                tri_a, tri_b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("ozaki_gemm_nt: %s", cudaGetErrorString(e));
        return SOCKS_ECUDA;
    }
    return SOCKS_OK;
}


// ---- KernelMaximalSR predict on the same kernel ------------------------------------------------------------------
template <int S>
static int srmax_predict_i8_impl(const double* X, const double* w, const double* vf, int64_t ldvf, const double* CUA,
                                 int64_t n, int d, int P, double kmul, const double* pts, int64_t m, const double* tubes,
                                 int problem, int T, double* out, int64_t ldout, uint8_t* work, int64_t chunk, double tau,
                                 int* list, cudaStream_t st) {
    int rc = og_configure<S, 7>();
    if (rc) return rc;
    const int kblocks = (int)(pad_to(n, OZ_KB) / OZ_KB);
    const int64_t ldbm = pad_to((int64_t)T * P, 128), ctiles = ldbm / OZ_BM;
    const int64_t mc = pad_to(std::min(chunk, m), OZ_BM), ptiles = mc / OZ_BM;
    uint8_t* Bop = work;
    uint8_t* Aop = Bop + pad_to(ctiles * kblocks * (int64_t)S * OZ_A_BYTES, 1024);
    double* Cm = reinterpret_cast<double*>(Aop + pad_to(ptiles * kblocks * (int64_t)S * OZ_A_BYTES, 1024));
    double* rowscale = Cm + mc * ldbm;
    double* inv = rowscale + pad_to(T, 2);
    double* colscale = inv + pad_to(T, 2);
    oz_rowscale_kernel<<<T, 256, 0, st>>>(w, vf, ldvf, n, rowscale, inv);
    oz_colscale_kernel<<<(unsigned)ceil_div(ldbm, 256), 256, 0, st>>>(rowscale, P, T, ldbm, colscale);
    oz_slice_coef_kernel<S><<<dim3((unsigned)ceil_div(kblocks, 16), (unsigned)ctiles), 128, 0, st>>>(CUA, n, P, w, vf, ldvf, T, inv,
                                                                                              Bop, kblocks);
    SOCKS_CHECK_LAUNCH();
    SOCKS_CUDA(cudaMemsetAsync(list, 0, sizeof(int), st));
    for (int64_t s0 = 0; s0 < m; s0 += mc) {
        const int64_t cnt = std::min(mc, m - s0), pt = ceil_div(cnt, OZ_BM);
        const double* p = pts + s0 * d;
        const dim3 gs((unsigned)ceil_div(kblocks, 16), (unsigned)pt);
#define SOCKS_OZ_CASE(DD)                                                                              \
    case DD:                                                                                           \
        oz_slice_points_kernel<DD, S><<<gs, 128, 0, st>>>(X, n, p, cnt, kmul, Aop, kblocks);           \
        break;
        switch (d) {
            SOCKS_OZ_CASE(1) SOCKS_OZ_CASE(2) SOCKS_OZ_CASE(3) SOCKS_OZ_CASE(4)
            SOCKS_OZ_CASE(5) SOCKS_OZ_CASE(6) SOCKS_OZ_CASE(7) SOCKS_OZ_CASE(8)
        }
#undef SOCKS_OZ_CASE
        SOCKS_CHECK_LAUNCH();
        // Cm (points x T P) = K(pts, X) . coef: unsigned 7-bit digits, no row scale (kernel values lie in [0, 1])
        oz_gemm_nt_kernel<S, 7><<<dim3((unsigned)ctiles, (unsigned)pt), OZ_THREADS, OgCfg<S>::SMEM_BYTES, st>>>(
            Aop, nullptr, Bop, colscale, Cm, ldbm, kblocks, 0, 1, 0, 0);
        SOCKS_CHECK_LAUNCH();
        oz_flag_kernel<<<(unsigned)ceil_div(cnt, 8), 256, 0, st>>>(Cm, ldbm, cnt, P, rowscale, tau, s0, list);
        SOCKS_CHECK_LAUNCH();
        rc = socks_srmax_step(Cm, ldbm, nullptr, 0, cnt, P, T, 0, p, d, tubes, problem, T, out + s0, ldout, 1,
                              (socks_stream_t)st);
        if (rc) return rc;
    }
    return SOCKS_OK;
}

}  // namespace socks

// ---- KernelMaximalSR fit on the same engine (kernel_sr_max.py:323-366) -----------------------------------------------
// The backward recursion is the predict contraction with the NEXT STATES Y as the points, one time step at a time
// (vf[t] needs vf[t+1]): the digits of K(x_i, y_j) are produced once, each step slices the two coefficient rows
// {w, vf[t+1] w} x CUA (2 P columns), runs the integer GEMM and applies the max / step.  The normaliser is the same at
// every step, so the points whose normaliser is too small for the fixed-point range are counted once: if any, the caller
// redoes the fit with the FP64 engine (socks_srmax_fit).
namespace socks {

template <int S>
static int srmax_fit_i8_impl(const double* X, const double* Y, const double* CUA, const double* w, int64_t n, int d, int P,
                             double kmul, const double* tubes, int problem, int T, double* vf, int64_t ldvf, uint8_t* work,
                             double tau, int* flagged, cudaStream_t st) {
    int rc = og_configure<S, 7>();
    if (rc) return rc;
    const int kblocks = (int)(pad_to(n, OZ_KB) / OZ_KB);
    const int64_t ldbm = pad_to(2 * (int64_t)P, 128), ctiles = ldbm / OZ_BM;
    const int64_t mc = pad_to(n, OZ_BM), ptiles = mc / OZ_BM;
    uint8_t* Bop = work;
    uint8_t* Aop = Bop + pad_to(ctiles * kblocks * (int64_t)S * OZ_A_BYTES, 1024);
    double* Cm = reinterpret_cast<double*>(Aop + pad_to(ptiles * kblocks * (int64_t)S * OZ_A_BYTES, 1024));
    double* rowscale = Cm + mc * ldbm;
    double* inv = rowscale + 2;
    double* colscale = inv + 2;
    int* list = reinterpret_cast<int*>(colscale + ldbm);  // [0] = count, [1..] = flagged points (only the count is returned)
    const dim3 gs((unsigned)ceil_div(kblocks, 16), (unsigned)ptiles);
#define SOCKS_OZ_CASE(DD)                                                                 \
    case DD:                                                                              \
        oz_slice_points_kernel<DD, S><<<gs, 128, 0, st>>>(X, n, Y, n, kmul, Aop, kblocks); \
        break;
    switch (d) {
        SOCKS_OZ_CASE(1) SOCKS_OZ_CASE(2) SOCKS_OZ_CASE(3) SOCKS_OZ_CASE(4)
        SOCKS_OZ_CASE(5) SOCKS_OZ_CASE(6) SOCKS_OZ_CASE(7) SOCKS_OZ_CASE(8)
    }
#undef SOCKS_OZ_CASE
    SOCKS_CHECK_LAUNCH();
    SOCKS_CUDA(cudaMemsetAsync(list, 0, sizeof(int), st));
    for (int t = T - 2; t >= 0; --t) {
        const double* vfp = vf + (int64_t)t * ldvf;  // row 1 of the pair = vf[t + 1]
        oz_rowscale_kernel<<<2, 256, 0, st>>>(w, vfp, ldvf, n, rowscale, inv);
        oz_colscale_kernel<<<(unsigned)ceil_div(ldbm, 256), 256, 0, st>>>(rowscale, P, 2, ldbm, colscale);
        oz_slice_coef_kernel<S><<<dim3((unsigned)ceil_div(kblocks, 16), (unsigned)ctiles), 128, 0, st>>>(CUA, n, P, w, vfp, ldvf, 2,
                                                                                                       inv, Bop, kblocks);
        SOCKS_CHECK_LAUNCH();
        oz_gemm_nt_kernel<S, 7><<<dim3((unsigned)ctiles, (unsigned)ptiles), OZ_THREADS, OgCfg<S>::SMEM_BYTES, st>>>(
            Aop, nullptr, Bop, colscale, Cm, ldbm, kblocks, 0, 1, 0, 0);
        SOCKS_CHECK_LAUNCH();
        if (t == T - 2) {
            oz_flag_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(Cm, ldbm, n, P, rowscale, tau, 0, list);
            SOCKS_CHECK_LAUNCH();
        }
        rc = socks_srmax_step(Cm, ldbm, nullptr, 0, n, P, 2, t, Y, d, tubes, problem, T, vf, ldvf, 0, (socks_stream_t)st);
        if (rc) return rc;
    }
    SOCKS_CUDA(cudaMemcpyAsync(flagged, list, sizeof(int), cudaMemcpyDeviceToDevice, st));
    return SOCKS_OK;
}

}  // namespace socks

extern "C" int64_t socks_srmax_fit_i8_work_bytes(int64_t n, int P, int slices) {
    const int64_t kblocks = pad_to(n, OZ_KB) / OZ_KB, ldbm = pad_to(2 * (int64_t)P, 128), mc = pad_to(n, OZ_BM);
    return pad_to((ldbm / OZ_BM) * kblocks * slices * OZ_A_BYTES, 1024) + pad_to((mc / OZ_BM) * kblocks * slices * OZ_A_BYTES, 1024) +
           (mc * ldbm + 4 + ldbm) * 8 + (n + 2) * 4 + 2048;
}

extern "C" int socks_srmax_fit_i8(const double* X, const double* Y, const double* U, const double* A, const double* w, int64_t n,
                                  int d, int du, int P, double scale, int divide, const double* tubes, int problem, int T,
                                  double* CUA, double* vf, int64_t ldvf, void* work, int slices, double tau, int* flagged,
                                  socks_stream_t stream) {
    SOCKS_CHECK_ARG(X && Y && U && A && w, 1);
    SOCKS_CHECK_ARG(n >= 1, 6);
    SOCKS_CHECK_ARG(d >= 1 && d <= 8, 7);
    SOCKS_CHECK_ARG(P >= 1, 9);
    SOCKS_CHECK_ARG(scale == scale && (!divide || scale != 0.0), 10);
    SOCKS_CHECK_ARG(T >= 1, 14);
    SOCKS_CHECK_ARG(CUA != nullptr, 15);
    SOCKS_CHECK_ARG(vf != nullptr && ldvf >= n, 16);
    SOCKS_CHECK_ARG(work != nullptr && (reinterpret_cast<uintptr_t>(work) & 1023) == 0, 18);
    SOCKS_CHECK_ARG(slices >= 6 && slices <= 8, 19);
    SOCKS_CHECK_ARG(tau >= 0.0, 20);
    SOCKS_CHECK_ARG(flagged != nullptr, 21);
    const double kmul = divide ? 1.0 / scale : scale;
    SOCKS_CHECK_ARG(kmul < 0.0, 10);
    int rc = socks_gram_rbf(U, n, A, P, du, scale, divide, nullptr, CUA, P, stream);  // kernel_sr_max.py:332
    if (rc) return rc;
    uint8_t* wk = reinterpret_cast<uint8_t*>(work);
    // vf[T-1] = 1_target(Y) (:59); Cm is not read for R = 1
    rc = socks_srmax_step(reinterpret_cast<const double*>(wk), pad_to(2 * (int64_t)P, 128), nullptr, 0, n, P, 1, 0, Y, d, tubes,
                          problem, T, vf, ldvf, 1, stream);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (T == 1) return cudaMemsetAsync(flagged, 0, sizeof(int), st) == cudaSuccess ? SOCKS_OK : SOCKS_ECUDA;
    switch (slices) {
        case 6: return srmax_fit_i8_impl<6>(X, Y, CUA, w, n, d, P, kmul, tubes, problem, T, vf, ldvf, wk, tau, flagged, st);
        case 7: return srmax_fit_i8_impl<7>(X, Y, CUA, w, n, d, P, kmul, tubes, problem, T, vf, ldvf, wk, tau, flagged, st);
        default: return srmax_fit_i8_impl<8>(X, Y, CUA, w, n, d, P, kmul, tubes, problem, T, vf, ldvf, wk, tau, flagged, st);
    }
}

extern "C" int64_t socks_srmax_predict_i8_work_bytes(int64_t n, int P, int T, int64_t chunk, int slices) {
    const int64_t kblocks = pad_to(n, OZ_KB) / OZ_KB, ldbm = pad_to((int64_t)T * P, 128), mc = pad_to(chunk, OZ_BM);
    return pad_to((ldbm / OZ_BM) * kblocks * slices * OZ_A_BYTES, 1024) + pad_to((mc / OZ_BM) * kblocks * slices * OZ_A_BYTES, 1024) +
           (mc * ldbm + 2 * pad_to(T, 2) + ldbm) * 8 + 1024;
}

extern "C" int socks_srmax_predict_i8(const double* X, const double* w, const double* vf, int64_t ldvf, const double* CUA,
                                      int64_t n, int d, int P, double scale, int divide, const double* pts, int64_t m,
                                      const double* tubes, int problem, int T, double* out, int64_t ldout, void* work,
                                      int64_t chunk, int slices, double tau, int* flagged, socks_stream_t stream) {
    SOCKS_CHECK_ARG(X && w && vf && CUA, 1);
    SOCKS_CHECK_ARG(n >= 1, 6);
    SOCKS_CHECK_ARG(d >= 1 && d <= 8, 7);
    SOCKS_CHECK_ARG(P >= 1, 8);
    SOCKS_CHECK_ARG(scale == scale && (!divide || scale != 0.0), 9);
    SOCKS_CHECK_ARG(m >= 0 && m < ((int64_t)1 << 31), 12);
    SOCKS_CHECK_ARG(T >= 2, 15);
    SOCKS_CHECK_ARG(ldout >= m, 17);
    SOCKS_CHECK_ARG(work != nullptr && (reinterpret_cast<uintptr_t>(work) & 1023) == 0, 18);
    SOCKS_CHECK_ARG(chunk >= 1, 19);
    SOCKS_CHECK_ARG(slices >= 6 && slices <= 8, 20);
    SOCKS_CHECK_ARG(tau >= 0.0, 21);
    SOCKS_CHECK_ARG(flagged != nullptr, 22);
    if (m == 0) return SOCKS_OK;
    SOCKS_CHECK_ARG(pts != nullptr && out != nullptr, 11);
    const double kmul = divide ? 1.0 / scale : scale;
    SOCKS_CHECK_ARG(kmul < 0.0, 9);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* wk = reinterpret_cast<uint8_t*>(work);
    switch (slices) {
        case 6: return srmax_predict_i8_impl<6>(X, w, vf, ldvf, CUA, n, d, P, kmul, pts, m, tubes, problem, T, out, ldout, wk, chunk, tau, flagged, st);
        case 7: return srmax_predict_i8_impl<7>(X, w, vf, ldvf, CUA, n, d, P, kmul, pts, m, tubes, problem, T, out, ldout, wk, chunk, tau, flagged, st);
        default: return srmax_predict_i8_impl<8>(X, w, vf, ldvf, CUA, n, d, P, kmul, pts, m, tubes, problem, T, out, ldout, wk, chunk, tau, flagged, st);
    }
}


extern "C" int64_t socks_dev_ozaki_gemm_work_bytes(int64_t m, int64_t n, int64_t K) {
    return ozaki_operand_bytes(m, K, 8) + ozaki_operand_bytes(n, K, 8) + 2048 + 2 * K * 8;  // + balancing scales
}

// C (m x n, ldc) -= A (m x K, lda) B (n x K, ldb)^T through the integer tensor cores; m, n multiples of 128, K of 32,
// `syrk` is a bit set: 1 = B is A (sliced once), 2 = 8 digits instead of 7, 4 = balance the operands along K (ozaki_kbalance;
// ignored with bit 0); lower_only as in socks_gemm_f64.  work: 1024-byte aligned.
extern "C" int socks_dev_ozaki_gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                                       int64_t m, int64_t n, int64_t K, int syrk, int lower_only, void* work,
                                       socks_stream_t stream) {
    SOCKS_CHECK_ARG(A != nullptr && lda >= K && lda % 2 == 0, 1);
    SOCKS_CHECK_ARG(syrk || (B != nullptr && ldb >= K && ldb % 2 == 0), 3);
    SOCKS_CHECK_ARG(C != nullptr && ldc >= n && ldc % 2 == 0, 5);
    SOCKS_CHECK_ARG(m > 0 && m % 128 == 0, 7);
    SOCKS_CHECK_ARG(n > 0 && n % 128 == 0 && (!syrk || n == m), 8);
    SOCKS_CHECK_ARG(K > 0 && K % 32 == 0, 9);
    SOCKS_CHECK_ARG(work != nullptr && (reinterpret_cast<uintptr_t>(work) & 1023) == 0, 12);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* abuf = reinterpret_cast<uint8_t*>(work);
    const int S = (syrk & 2) ? 8 : 7;
    const bool balance = (syrk & 4) && !(syrk & 1);
    syrk &= 1;
    uint8_t* bbuf = syrk ? abuf : abuf + ozaki_operand_bytes(m, K, S);
    int rc = SOCKS_OK;
    const double *ksa = nullptr, *ksb = nullptr;
    if (balance) {
        double* kw = reinterpret_cast<double*>(abuf + ozaki_operand_bytes(m, K, 8) + ozaki_operand_bytes(n, K, 8) + 1024);
        if ((rc = ozaki_kbalance(A, m, lda, 0, 0, B, n, ldb, 0, 0, K, kw, st))) return rc;
        ksa = kw;
        ksb = kw + K;
    }
    rc = ozaki_slice_operand(A, m, K, lda, 0, 0, abuf, S, ksa, st);
    if (rc) return rc;
    if (!syrk && (rc = ozaki_slice_operand(B, n, K, ldb, 0, 0, bbuf, S, ksb, st))) return rc;
    return ozaki_gemm_nt_sliced(abuf, m, 0, bbuf, syrk ? m : n, 0, K, C, ldc, m, n, lower_only, 0, 0, 0, S, st);
}
