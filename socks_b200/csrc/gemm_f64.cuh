// fp64 tensor-core GEMM for the dense contractions of the hot path:
//   C(m x n) = alpha * op(A)(m x K) * op(B)(K x n) + beta * C
// 128 x 128 x 16 CTA tiles, 8 warps of 64 x 32, mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 — tcgen05
// has no f64 kind, SURVEY §2 hardware note), 4-stage cp.async pipeline, operands kept in their
// native row-major orientation in shared memory so every global read is a 16-byte coalesced chunk.
//
// Operand forms (all row-major):
//   TA = false: A is m x K (lda), element (i,k) at A[i*lda + k]
//   TA = true : A is K x m (lda), element (i,k) at A[k*lda + i]
//   TB = true : B is n x K (ldb), element (k,j) at B[j*ldb + k]      ("NT")
//   TB = false: B is K x n (ldb), element (k,j) at B[k*ldb + j]      ("NN")
// Requirements: m, n multiples of 128; K multiple of 16; lda, ldb, ldc even; 16-byte aligned bases.
// Triangular structure is exploited at tile level through `flags` (k-range clipping per output tile)
// and `lower_only` (skip output tiles strictly above the block diagonal).
#pragma once
#include "common.cuh"

namespace socks {

constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 16, GM_STAGES = 4, GM_THREADS = 256;
constexpr int GM_LDK = GM_BK + 4;    // 20: row stride of a [128][16] tile; 20 mod 16 = 4 -> conflict-free frags
constexpr int GM_LDM = GM_BM + 4;    // 132: row stride of a [16][128] tile; 132 mod 16 = 4
constexpr int GM_TILE_DBL = 128 * GM_LDK;  // 2560 doubles >= 16 * 132 = 2112
constexpr int GM_SMEM_BYTES = GM_STAGES * 2 * GM_TILE_DBL * 8;  // 163840

enum GemmFlags : int {
    GEMM_A_K_LE_M = 1,  // A(i,k) == 0 for k >  i  (A lower-triangular in (m,k))  -> k_hi = m0 + 128
    GEMM_A_K_GE_M = 2,  // A(i,k) == 0 for k <  i                                  -> k_lo = m0
    GEMM_B_K_GE_N = 4,  // B(k,j) == 0 for k <  j                                  -> k_lo = n0
    GEMM_B_K_LE_N = 8,  // B(k,j) == 0 for k >  j                                  -> k_hi = n0 + 128
};

template <bool TRANS>
__device__ __forceinline__ void gemm_load_tile(double* __restrict__ s, const double* __restrict__ g, int64_t ld,
                                               int64_t mn0, int64_t k0, int tid) {
    // TRANS == false: tile [128 (m or n)][16 k] from rows mn0.., cols k0..   (global row-major "mn x K")
    // TRANS == true : tile [16 k][128 (m or n)] from rows k0.., cols mn0..   (global row-major "K x mn")
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        int c = tid + it * GM_THREADS;  // 1024 chunks of 2 doubles
        if (!TRANS) {
            int r = c >> 3, kc = (c & 7) * 2;
            cp_async16(s + r * GM_LDK + kc, g + (mn0 + r) * ld + k0 + kc);
        } else {
            int r = c >> 6, mc = (c & 63) * 2;
            cp_async16(s + r * GM_LDM + mc, g + (k0 + r) * ld + mn0 + mc);
        }
    }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(GM_THREADS, 1)
    gemm_f64_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                    double* __restrict__ C, int64_t ldc, int64_t K, double alpha, double beta, int flags,
                    int lower_only) {
    extern __shared__ __align__(16) double gm_smem[];
    const int64_t m0 = (int64_t)blockIdx.y * GM_BM;
    const int64_t n0 = (int64_t)blockIdx.x * GM_BN;
    if (lower_only && n0 > m0) return;

    int64_t k_lo = 0, k_hi = K;
    if (flags & GEMM_A_K_LE_M) k_hi = min(k_hi, m0 + GM_BM);
    if (flags & GEMM_A_K_GE_M) k_lo = max(k_lo, m0);
    if (flags & GEMM_B_K_GE_N) k_lo = max(k_lo, n0);
    if (flags & GEMM_B_K_LE_N) k_hi = min(k_hi, n0 + GM_BN);
    const int nkt = (k_hi > k_lo) ? (int)((k_hi - k_lo) / GM_BK) : 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
    const int lq = lane >> 2, lr = lane & 3;  // lq: row-in-8 (A) / col-in-8 (B); lr: k-in-4

    double acc[8][4][2];
    const bool use_c = (beta != 0.0);
    const double cscale = use_c ? beta / alpha : 0.0;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            if (use_c) {
                const double2 c = *reinterpret_cast<const double2*>(
                    C + (m0 + wm + mi * 8 + lq) * ldc + n0 + wn + ni * 8 + 2 * lr);
                acc[mi][ni][0] = c.x * cscale;
                acc[mi][ni][1] = c.y * cscale;
            } else {
                acc[mi][ni][0] = 0.0;
                acc[mi][ni][1] = 0.0;
            }
        }

    double* sA = gm_smem;
    double* sB = gm_smem + GM_STAGES * GM_TILE_DBL;

    // prologue
#pragma unroll
    for (int s = 0; s < GM_STAGES - 1; ++s) {
        if (s < nkt) {
            gemm_load_tile<TA>(sA + s * GM_TILE_DBL, A, lda, m0, k_lo + (int64_t)s * GM_BK, tid);
            gemm_load_tile<!TB>(sB + s * GM_TILE_DBL, B, ldb, n0, k_lo + (int64_t)s * GM_BK, tid);
        }
        cp_async_commit();
    }

    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<GM_STAGES - 2>();
        __syncthreads();
        {
            const int nx = kt + GM_STAGES - 1;
            if (nx < nkt) {
                const int s = nx % GM_STAGES;
                gemm_load_tile<TA>(sA + s * GM_TILE_DBL, A, lda, m0, k_lo + (int64_t)nx * GM_BK, tid);
                gemm_load_tile<!TB>(sB + s * GM_TILE_DBL, B, ldb, n0, k_lo + (int64_t)nx * GM_BK, tid);
            }
            cp_async_commit();
        }
        const double* a_s = sA + (kt % GM_STAGES) * GM_TILE_DBL;
        const double* b_s = sB + (kt % GM_STAGES) * GM_TILE_DBL;
#pragma unroll
        for (int ks = 0; ks < GM_BK / 4; ++ks) {
            double af[8], bf[4];
            const int kk = ks * 4 + lr;
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) {
                const int r = wm + mi * 8 + lq;
                af[mi] = TA ? a_s[kk * GM_LDM + r] : a_s[r * GM_LDK + kk];
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int c = wn + ni * 8 + lq;
                bf[ni] = TB ? b_s[c * GM_LDK + kk] : b_s[kk * GM_LDM + c];
            }
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            double2 c = make_double2(alpha * acc[mi][ni][0], alpha * acc[mi][ni][1]);
            *reinterpret_cast<double2*>(C + (m0 + wm + mi * 8 + lq) * ldc + n0 + wn + ni * 8 + 2 * lr) = c;
        }
}

// Host launcher. Returns cudaGetLastError() of the launch.
template <bool TA, bool TB>
inline cudaError_t gemm_f64(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                            int64_t m, int64_t n, int64_t K, double alpha, double beta, int flags, int lower_only,
                            cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_f64_kernel<TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             GM_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (m == 0 || n == 0) return cudaSuccess;
    dim3 grid((unsigned)(n / GM_BN), (unsigned)(m / GM_BM));
    gemm_f64_kernel<TA, TB><<<grid, GM_THREADS, GM_SMEM_BYTES, st>>>(A, lda, B, ldb, C, ldc, K, alpha, beta, flags,
                                                                     lower_only);
    return cudaGetLastError();
}

}  // namespace socks
