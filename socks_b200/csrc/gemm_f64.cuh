// fp64 tensor-core GEMM for the dense contractions of the hot path:
//   C(m x n) = alpha * op(A)(m x K) * op(B)(K x n) + beta * C
// BM x BN x 16 CTA tiles, WM x WN warp tiles of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 -- tcgen05 has no
// f64 kind, SURVEY §2 hardware note), 4-stage cp.async pipeline, operands kept in their native row-major
// orientation in shared memory so every global read is a 16-byte coalesced chunk.
//
// Operand forms (all row-major):
//   TA = false: A is m x K (lda), element (i,k) at A[i*lda + k]
//   TA = true : A is K x m (lda), element (i,k) at A[k*lda + i]
//   TB = true : B is n x K (ldb), element (k,j) at B[j*ldb + k]      ("NT")
//   TB = false: B is K x n (ldb), element (k,j) at B[k*ldb + j]      ("NN")
// Requirements: m, n multiples of 128; K multiple of 16; lda, ldb, ldc even; 16-byte aligned bases.
// Triangular structure is exploited at tile level through `flags` (k-range clipping per output tile, in
// units of 128) and `lower_only` (skip output tiles strictly above the block diagonal).
// Default tile configuration: 64 x 64 CTA tiles (4 warps of 32 x 32), 3 stages, three CTAs per SM.
#pragma once
#include <atomic>

#include "common.cuh"

namespace socks {

constexpr int GM_BK = 16;
constexpr int GM_LDK = GM_BK + 4;  // 20: row stride of a [BM][16] tile; 20 mod 16 = 4 -> conflict-free frags

enum GemmFlags : int {
    GEMM_A_K_LE_M = 1,  // A(i,k) == 0 for k >  i  (A lower-triangular in (m,k))  -> k_hi = m0 + 128-block end
    GEMM_A_K_GE_M = 2,  // A(i,k) == 0 for k <  i                                  -> k_lo = 128-block start of m0
    GEMM_B_K_GE_N = 4,  // B(k,j) == 0 for k <  j                                  -> k_lo = 128-block start of n0
    GEMM_B_K_LE_N = 8,  // B(k,j) == 0 for k >  j                                  -> k_hi = 128-block end of n0
};

template <int BM, int BN, int WM, int WN, int STAGES>
struct GemmCfg {
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
    static constexpr int THREADS = WARPS_M * WARPS_N * 32;
    static constexpr int MI = WM / 8, NI = WN / 8;
    static constexpr int LDA_T = BM + 4, LDB_T = BN + 4;  // strides of [16][BM] / [16][BN] tiles (= 4 mod 16)
    static constexpr int A_DBL = (BM * GM_LDK > GM_BK * LDA_T) ? BM * GM_LDK : GM_BK * LDA_T;
    static constexpr int B_DBL = (BN * GM_LDK > GM_BK * LDB_T) ? BN * GM_LDK : GM_BK * LDB_T;
    static constexpr int SMEM_BYTES = STAGES * (A_DBL + B_DBL) * 8;
};

// ROWS x 16 doubles of a row-major "rows x K" operand (TRANS=false) or 16 x ROWS of a "K x rows" operand.
template <bool TRANS, int ROWS, int THREADS>
__device__ __forceinline__ void gemm_load_tile(double* __restrict__ s, const double* __restrict__ g, int64_t ld,
                                               int64_t mn0, int64_t k0, int tid) {
    constexpr int CHUNKS = ROWS * GM_BK / 2;  // 16-byte chunks
#pragma unroll
    for (int it = 0; it < (CHUNKS + THREADS - 1) / THREADS; ++it) {
        const int c = tid + it * THREADS;
        if (CHUNKS % THREADS != 0 && c >= CHUNKS) break;
        if (!TRANS) {
            const int r = c >> 3, kc = (c & 7) * 2;
            cp_async16(s + r * GM_LDK + kc, g + (mn0 + r) * ld + k0 + kc);
        } else {
            const int r = c / (ROWS / 2), mc = (c % (ROWS / 2)) * 2;
            cp_async16(s + r * (ROWS + 4) + mc, g + (k0 + r) * ld + mn0 + mc);
        }
    }
}

template <bool TA, bool TB, int BM, int BN, int WM, int WN, int STAGES, int MINB>
__global__ void __launch_bounds__(GemmCfg<BM, BN, WM, WN, STAGES>::THREADS, MINB)
    gemm_f64_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                    double* __restrict__ C, int64_t ldc, int64_t K, double alpha, double beta, int flags,
                    int lower_only) {
    using Cfg = GemmCfg<BM, BN, WM, WN, STAGES>;
    constexpr int GM_STAGES = STAGES;
    constexpr int THREADS = Cfg::THREADS, MI = Cfg::MI, NI = Cfg::NI;
    extern __shared__ __align__(16) double gm_smem[];
    // k_hi grows with m0 for a lower-triangular A: launch the long tiles first (no tail of heavy CTAs)
    const int by = (flags & GEMM_A_K_LE_M) ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
    const int64_t m0 = (int64_t)by * BM;
    const int64_t n0 = (int64_t)blockIdx.x * BN;
    if (lower_only && (n0 / TILE) > (m0 / TILE)) return;  // 128-tile strictly above the block diagonal

    int64_t k_lo = 0, k_hi = K;
    if (flags & GEMM_A_K_LE_M) k_hi = min(k_hi, (m0 / TILE + 1) * TILE);
    if (flags & GEMM_A_K_GE_M) k_lo = max(k_lo, (m0 / TILE) * TILE);
    if (flags & GEMM_B_K_GE_N) k_lo = max(k_lo, (n0 / TILE) * TILE);
    if (flags & GEMM_B_K_LE_N) k_hi = min(k_hi, (n0 / TILE + 1) * TILE);
    const int nkt = (k_hi > k_lo) ? (int)((k_hi - k_lo) / GM_BK) : 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp / Cfg::WARPS_N) * WM, wn = (warp % Cfg::WARPS_N) * WN;
    const int lq = lane >> 2, lr = lane & 3;  // lq: row-in-8 (A) / col-in-8 (B); lr: k-in-4

    double acc[MI][NI][2];
    const bool use_c = (beta != 0.0);
    const double cscale = use_c ? beta / alpha : 0.0;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
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
    double* sB = gm_smem + GM_STAGES * Cfg::A_DBL;

#pragma unroll
    for (int s = 0; s < GM_STAGES - 1; ++s) {
        if (s < nkt) {
            gemm_load_tile<TA, BM, THREADS>(sA + s * Cfg::A_DBL, A, lda, m0, k_lo + (int64_t)s * GM_BK, tid);
            gemm_load_tile<!TB, BN, THREADS>(sB + s * Cfg::B_DBL, B, ldb, n0, k_lo + (int64_t)s * GM_BK, tid);
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
                gemm_load_tile<TA, BM, THREADS>(sA + s * Cfg::A_DBL, A, lda, m0, k_lo + (int64_t)nx * GM_BK, tid);
                gemm_load_tile<!TB, BN, THREADS>(sB + s * Cfg::B_DBL, B, ldb, n0, k_lo + (int64_t)nx * GM_BK, tid);
            }
            cp_async_commit();
        }
        const double* a_s = sA + (kt % GM_STAGES) * Cfg::A_DBL;
        const double* b_s = sB + (kt % GM_STAGES) * Cfg::B_DBL;
#pragma unroll
        for (int ks = 0; ks < GM_BK / 4; ++ks) {
            double af[MI], bf[NI];
            const int kk = ks * 4 + lr;
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                const int r = wm + mi * 8 + lq;
                af[mi] = TA ? a_s[kk * Cfg::LDA_T + r] : a_s[r * GM_LDK + kk];
            }
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                const int c = wn + ni * 8 + lq;
                bf[ni] = TB ? b_s[c * GM_LDK + kk] : b_s[kk * Cfg::LDB_T + c];
            }
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            double2 c = make_double2(alpha * acc[mi][ni][0], alpha * acc[mi][ni][1]);
            *reinterpret_cast<double2*>(C + (m0 + wm + mi * 8 + lq) * ldc + n0 + wn + ni * 8 + 2 * lr) = c;
        }
}

template <bool TA, bool TB, int BM, int BN, int WM, int WN, int STAGES, int MINB>
inline cudaError_t gemm_f64_launch(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                                   int64_t m, int64_t n, int64_t K, double alpha, double beta, int flags,
                                   int lower_only, cudaStream_t st) {
    using Cfg = GemmCfg<BM, BN, WM, WN, STAGES>;
    auto kern = gemm_f64_kernel<TA, TB, BM, BN, WM, WN, STAGES, MINB>;
    // the opt-in to > 48 KB of dynamic shared memory is a per-DEVICE attribute: track it per device ordinal
    static std::atomic<bool> configured[SOCKS_MAX_DEVICES];
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess) return e0;
    if (dev < 0 || dev >= SOCKS_MAX_DEVICES || !configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < SOCKS_MAX_DEVICES) configured[dev].store(true, std::memory_order_release);
    }
    dim3 grid((unsigned)(n / BN), (unsigned)(m / BM));
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(A, lda, B, ldb, C, ldc, K, alpha, beta, flags, lower_only);
    return cudaGetLastError();
}

// Host launcher: picks the tile configuration from the number of 128 x 128 output tiles.
// tile_cfg: 0 = default: 64 x 64 tiles, 4 warps of 32 x 32, 3-stage pipeline, three CTAs per SM -- measured
// 35.0-35.4 TFLOP/s (95 % of the 37.1 TFLOP/s DMMA peak) on the factorisation's large shapes, against
// 30.8 for 1 = 128 x 128 tiles / 8 warps of 64 x 32 / one CTA per SM (profiles/r1_gemm_tile_sweep.txt):
// with one DMMA pipe per SM sub-partition, three independent CTAs hide the fragment-load and barrier
// latency that two resident warps per scheduler cannot.
template <bool TA, bool TB>
inline cudaError_t gemm_f64(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                            int64_t m, int64_t n, int64_t K, double alpha, double beta, int flags, int lower_only,
                            cudaStream_t st, int tile_cfg = 0) {
    if (m == 0 || n == 0) return cudaSuccess;
    const int64_t tm = m / TILE, tn = n / TILE;
    const int64_t tiles = lower_only ? (tm * (tm + 1)) / 2 : tm * tn;
    (void)tiles;
    if (tile_cfg == 1)
        return gemm_f64_launch<TA, TB, 128, 128, 64, 32, 4, 1>(A, lda, B, ldb, C, ldc, m, n, K, alpha, beta, flags,
                                                               lower_only, st);
    return gemm_f64_launch<TA, TB, 64, 64, 32, 32, 3, 3>(A, lda, B, ldb, C, ldc, m, n, K, alpha, beta, flags,
                                                         lower_only, st);
}

}  // namespace socks
