// Regularised solve: blocked recursive Cholesky, triangular inverse, diag(inv(G)) and inv(G).
// Replaces `inv(K + lambda N I)` (LAPACK dgesv through numpy.linalg.inv,
// gym_socks/kernel/metrics.py:278-280).  G is symmetric positive definite (Gram + lambda N I), so
// the LU inverse of the reference becomes G = L L^T, M = inv(L), W = M^T M; the SR algorithms only
// ever read diag(W) (kernel_sr.py:45), i.e. the column sums of squares of M.
//
// Structure: recursion on the host over 128-aligned splits; every flop outside the 128 x 128
// diagonal leaves is a DMMA GEMM (gemm_f64.cuh):
//   potrf(A)  = potrf(A11); A21 <- A21 inv(L11)^T (recursive TRSM); A22 -= A21 A21^T; potrf(A22)
//   trsm(B,L) = trsm(B1,L11); B2 -= B1 L21^T; trsm(B2,L22)      leaf: B <- B inv(L_kk)^T (GEMM)
//   trtri     = M11 = inv(L11), M22 = inv(L22), M21 = -(M22 L21) M11
// Roofline: N^3/3 flops each for potrf, trtri and lauum; FP64 DMMA (tensor pipe) bound.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "gemm_f64.cuh"
#include "ozaki.cuh"
#include "reduce.cuh"

namespace socks {

constexpr int LF = TILE;        // leaf size 128
constexpr int LF_LD = LF + 1;   // odd smem stride: row and column accesses are both conflict-free
constexpr int LF_THREADS = 512;
constexpr int LF_NB = 32;       // inner block: one warp factors a 32 x 32 diagonal block in registers
constexpr int LF_SC_LD = 97;    // scratch [32][97] for the inverse assembly
// S (L, then inv L in place), Dinv[4][32][33], scratch 32*97 (+ slack), rinv[128]  (~190 KB)
constexpr int LF_SMEM = (LF * LF_LD + 4 * LF_NB * 33 + 96 * 33 + LF) * 8;

// 1/sqrt(d) for the pivot recurrence, branch-free so the whole unrolled factorisation stays one basic block
// and the scheduler can interleave a column's trailing updates with the next pivot's chain: hardware seed
// (rsqrt.approx.ftz.f64 -> MUFU.RSQ64H, relative error ~2^-22) and ONE third-order step
// y (1 + r/2 + 3 r^2/8), r = 1 - d y^2 (error ~ r^3 < 2^-60: the result is within ~1 ulp).
// Valid for d in the normal range; the caller treats pivots outside [1e-290, 1e290] as failures.
__device__ __forceinline__ double rsqrt_pivot(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double t = d * y;
    const double r = fma(-t, y, 1.0);
    const double u = y * r;
    const double s2 = fma(r, 0.375, 0.5);
    return fma(u, s2, y);
}

// (a) Cholesky of the 32 x 32 block at (o, o) of S by ONE warp, entirely in registers: lane i owns row i,
// pivots and multipliers travel by shuffle, the pivot is inverted with rsqrt (rinv[o+j] = 1 / L_jj).
__device__ __forceinline__ void chol32_warp(double* __restrict__ S, int o, double* __restrict__ rinv,
                                            int* __restrict__ info, int col_offset, int lane) {
    double a[LF_NB];
    double* row = S + (o + lane) * LF_LD + o;
#pragma unroll
    for (int k = 0; k < LF_NB; ++k) a[k] = row[k];  // strictly-upper entries of S are zero
    double dnext = a[0];  // lane j's own updated diagonal, formed without waiting for the shuffle of l
    int first_bad = -1;
#pragma unroll
    for (int j = 0; j < LF_NB; ++j) {
        double d = __shfl_sync(0xffffffffu, dnext, j);
        const bool bad = !(d > 1e-290 && d < 1e290);  // non-positive, NaN or out-of-range pivot: continue with 1
        first_bad = (bad && first_bad < 0) ? j : first_bad;
        d = bad ? 1.0 : d;
        const double ri = rsqrt_pivot(d);
        const double l = ((lane == j) ? d : a[j]) * ri;  // lane j: sqrt(d); lanes > j: L[i][j]
        a[j] = l;
        if (lane == j) rinv[o + j] = ri;
        if (j + 1 < LF_NB) dnext = fma(-l, l, a[j + 1]);  // valid in lane j+1 (its a[j+1] is the diagonal)
#pragma unroll
        for (int k = j + 1; k < LF_NB; ++k) {
            const double lk = __shfl_sync(0xffffffffu, l, k);
            a[k] = fma(-l, lk, a[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < LF_NB; ++k)
        if (k <= lane) row[k] = a[k];
    if (first_bad >= 0 && lane == 0) atomicCAS(info, 0, col_offset + o + first_bad + 1);  // report the first one
}

// (b) D = inv(L_oo) (32 x 32 lower) by one warp: lane c owns column c (forward substitution in registers,
// L entries are warp-uniform shared-memory broadcasts).
__device__ __forceinline__ void inv32_warp(const double* __restrict__ S, int o, const double* __restrict__ rinv,
                                           double* __restrict__ D, int lane) {
    double m[LF_NB];
#pragma unroll
    for (int i = 0; i < LF_NB; ++i) {
        const double* lrow = S + (o + i) * LF_LD + o;
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) {
            if (k & 1) acc1 = fma(lrow[k], m[k], acc1);
            else acc0 = fma(lrow[k], m[k], acc0);
        }
        const double ri = rinv[o + i];
        m[i] = (i < lane) ? 0.0 : ((i == lane) ? ri : -(acc0 + acc1) * ri);
    }
#pragma unroll
    for (int i = 0; i < LF_NB; ++i) D[i * 33 + lane] = m[i];
}

// One 8 x 8 output tile of A(8 x 4*ksteps) * B on the fp64 tensor core, operands in shared memory.
//   a(i,k) = Ap[i*lda + k];  TB: b(k,j) = Bp[j*ldb + k];  !TB: b(k,j) = Bp[k*ldb + j].
// Returns this lane's two accumulators C[lq][2*lr], C[lq][2*lr+1].
template <bool TB>
__device__ __forceinline__ double2 smem_tile8(const double* __restrict__ Ap, int lda, const double* __restrict__ Bp,
                                              int ldb, int ksteps, int lq, int lr) {
    double c0 = 0.0, c1 = 0.0;
#pragma unroll 4
    for (int ks = 0; ks < ksteps; ++ks) {
        const double a = Ap[lq * lda + ks * 4 + lr];
        const double b = TB ? Bp[lq * ldb + ks * 4 + lr] : Bp[(ks * 4 + lr) * ldb + lq];
        dmma884(c0, c1, a, b);
    }
    return make_double2(c0, c1);
}

// 16 x 16 output block = 2 x 2 tiles with four independent accumulator chains (DMMA latency is ~4x its issue
// interval, so one 8 x 8 tile per warp leaves the pipe idle).  c[a][b] is the tile at rows 8a, cols 8b.
template <bool TB>
__device__ __forceinline__ void smem_tile16(const double* __restrict__ Ap, int lda, const double* __restrict__ Bp,
                                            int ldb, int ksteps, int lq, int lr, double2 (&c)[2][2]) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) c[a][b] = make_double2(0.0, 0.0);
#pragma unroll 2
    for (int ks = 0; ks < ksteps; ++ks) {
        const int k = ks * 4 + lr;
        const double a0 = Ap[lq * lda + k], a1 = Ap[(lq + 8) * lda + k];
        const double b0 = TB ? Bp[lq * ldb + k] : Bp[k * ldb + lq];
        const double b1 = TB ? Bp[(lq + 8) * ldb + k] : Bp[k * ldb + lq + 8];
        dmma884(c[0][0].x, c[0][0].y, a0, b0);
        dmma884(c[0][1].x, c[0][1].y, a0, b1);
        dmma884(c[1][0].x, c[1][0].y, a1, b0);
        dmma884(c[1][1].x, c[1][1].y, a1, b1);
    }
}

// Factor one 128 x 128 diagonal block in shared memory and invert the factor.
//   A (global, lower part) <- L ;  dinv (dense 128 x 128) <- inv(L) with explicit zeros above.
// Right-looking over four 32-column panels: (a) one warp factors the 32 x 32 pivot block in registers,
// (c) one thread per row solves the panel below it against L_kk^T, (d) all threads update the trailing
// lower triangle; then the four pivot-block inverses are formed in parallel (one warp each) and inv(L) is
// assembled block row by block row, in place.
template <bool TIMING>
__global__ void __launch_bounds__(LF_THREADS, 1)
    potf2_inv_kernel(double* __restrict__ A, int64_t lda, double* __restrict__ dinv, int64_t ldd,
                     int* __restrict__ info, int col_offset, long long* __restrict__ cycles) {
#define LF_TICK(slot)                                       \
    if (TIMING && threadIdx.x == 0) cycles[slot] = clock64();
    extern __shared__ __align__(16) double lf_smem[];
    LF_TICK(0)
    double* S = lf_smem;                     // [128][129]  block -> L (lower) -> inv(L), block row by block row
    double* Di = S + LF * LF_LD;             // [4][32][33] inverses of the diagonal 32-blocks
    double* Sc = Di + 4 * LF_NB * 33;        // scratch
    double* rinv = Sc + 96 * 33;             // [128] 1 / L_jj
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

#pragma unroll 4
    for (int e = tid; e < LF * LF / 2; e += LF_THREADS) {  // 16-byte loads, all in flight before first use
        const int r = e >> 6, c = (e & 63) * 2;
        const double2 v = *reinterpret_cast<const double2*>(A + (int64_t)r * lda + c);
        S[r * LF_LD + c] = (c <= r) ? v.x : 0.0;
        S[r * LF_LD + c + 1] = (c + 1 <= r) ? v.y : 0.0;
    }
    __syncthreads();
    LF_TICK(1)

    for (int kb = 0; kb < 4; ++kb) {
        const int o = kb * LF_NB;
        if (warp == 0) chol32_warp(S, o, rinv, info, col_offset, lane);
        __syncthreads();
        LF_TICK(2 + 3 * kb)
        const int rem = LF - o - LF_NB;  // rows below the pivot block
        if (rem > 0) {
            // ---- (c) panel solve X L_kk^T = P, one thread per row, row in registers
            if (tid < rem) {
                double* prow = S + (o + LF_NB + tid) * LF_LD + o;
                double x[LF_NB];
#pragma unroll
                for (int k = 0; k < LF_NB; ++k) x[k] = prow[k];
#pragma unroll
                for (int c = 0; c < LF_NB; ++c) {
                    const double* lrow = S + (o + c) * LF_LD + o;  // warp-uniform: broadcast reads
                    double acc0 = x[c], acc1 = 0.0;
#pragma unroll
                    for (int k = 0; k < c; ++k) {
                        if (k & 1) acc1 = fma(-x[k], lrow[k], acc1);
                        else acc0 = fma(-x[k], lrow[k], acc0);
                    }
                    x[c] = (acc0 + acc1) * rinv[o + c];
                }
#pragma unroll
                for (int k = 0; k < LF_NB; ++k) prow[k] = x[k];
            }
            __syncthreads();
            LF_TICK(3 + 3 * kb)
            // ---- (d) trailing lower triangle -= X X^T: 16 x 16 DMMA blocks, one warp per block (K = 32)
            {
                const int nt = rem >> 4, lq = lane >> 2, lr = lane & 3;
                const double* Xp = S + (o + LF_NB) * LF_LD + o;
                for (int e = warp; e < nt * nt; e += LF_THREADS / 32) {
                    const int ti = e / nt, tj = e % nt;
                    if (tj > ti) continue;
                    double2 c[2][2];
                    smem_tile16<true>(Xp + ti * 16 * LF_LD, LF_LD, Xp + tj * 16 * LF_LD, LF_LD, LF_NB / 4, lq, lr, c);
#pragma unroll
                    for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                        for (int b2 = 0; b2 < 2; ++b2) {
                            double* dst = S + (o + LF_NB + ti * 16 + a2 * 8 + lq) * LF_LD + o + LF_NB + tj * 16 + b2 * 8 + 2 * lr;
                            dst[0] -= c[a2][b2].x;
                            dst[1] -= c[a2][b2].y;
                        }
                }
            }
            __syncthreads();
            LF_TICK(4 + 3 * kb)
        }
    }

    // ---- L is final.  Warps 0-3 invert the four pivot blocks; everyone writes L out.
    if (warp < 4) inv32_warp(S, warp * LF_NB, rinv, Di + warp * LF_NB * 33, lane);
    for (int e = tid; e < LF * LF; e += LF_THREADS) {
        const int r = e >> 7, c = e & 127;
        if (c <= r) A[(int64_t)r * lda + c] = S[r * LF_LD + c];
    }
    __syncthreads();
    LF_TICK(12)
    // ---- turn S into inv(L) in place.  Block row i of inv(L) only needs block row i of L and the finished
    //      rows above it:  M_ii = D_i,  M_ij = -D_i (sum_{k=j}^{i-1} L_ik M_kj).
    for (int e = tid; e < 4 * LF_NB * LF_NB; e += LF_THREADS) {
        const int b = e >> 10, r = (e >> 5) & 31, c = e & 31;
        S[(b * LF_NB + r) * LF_LD + b * LF_NB + c] = Di[b * LF_NB * 33 + r * 33 + c];  // zeros above the diagonal
    }
    __syncthreads();
    for (int i = 1; i < 4; ++i) {
        const int w = i * LF_NB;  // columns 0 .. w-1
        const int lq = lane >> 2, lr = lane & 3;
        // stage 1: T (32 x w) = L_i,: (32 x w) * M (w x w lower), 16 x 16 DMMA blocks; k starts at the column's
        // 32-block (explicit zeros above the diagonal inside it)
        const int ntc16 = w >> 4;
        for (int e = warp; e < 2 * ntc16; e += LF_THREADS / 32) {
            const int tr = e / ntc16, tc = e % ntc16;
            const int ks0 = (tc * 16) & ~31;
            double2 c[2][2];
            smem_tile16<false>(S + (w + tr * 16) * LF_LD + ks0, LF_LD, S + ks0 * LF_LD + tc * 16, LF_LD, (w - ks0) >> 2,
                               lq, lr, c);
#pragma unroll
            for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2) {
                    double* dst = Sc + (tr * 16 + a2 * 8 + lq) * LF_SC_LD + tc * 16 + b2 * 8 + 2 * lr;
                    dst[0] = c[a2][b2].x;
                    dst[1] = c[a2][b2].y;
                }
        }
        __syncthreads();
        // stage 2: M_i,: = -D_i * T  (D_i lower with explicit zeros above: k runs to the end of the row block)
        const double* D = Di + i * LF_NB * 33;
        for (int e = warp; e < 2 * ntc16; e += LF_THREADS / 32) {
            const int tr = e / ntc16, tc = e % ntc16;
            double2 c[2][2];
            smem_tile16<false>(D + tr * 16 * 33, 33, Sc + tc * 16, LF_SC_LD, (tr * 16 + 16) >> 2, lq, lr, c);
#pragma unroll
            for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2) {
                    double* dst = S + (w + tr * 16 + a2 * 8 + lq) * LF_LD + tc * 16 + b2 * 8 + 2 * lr;
                    dst[0] = -c[a2][b2].x;
                    dst[1] = -c[a2][b2].y;
                }
        }
        __syncthreads();
    }
    LF_TICK(13)
    for (int e = tid; e < LF * LF; e += LF_THREADS) {
        const int r = e >> 7, c = e & 127;
        dinv[(int64_t)r * ldd + c] = (c <= r) ? S[r * LF_LD + c] : 0.0;
    }
    LF_TICK(14)
#undef LF_TICK
}

// Leaf TRSM: B (rows x 128, in place) <- B * Minv^T, Minv = inv(L_kk) lower triangular (dense 128x128).
// One CTA per 64 rows; both operands fully resident in shared memory before any write, so the
// in-place update is race-free.  out[i][j] = sum_{k <= j} B[i][k] Minv[j][k].
constexpr int TR_ROWS = 64, TR_LD = LF + 4;  // 132
constexpr int TR_SMEM = (TR_ROWS + LF) * TR_LD * 8;

__global__ void __launch_bounds__(256, 1)
    trsm_leaf_kernel(double* __restrict__ Bp, int64_t ldb, const double* __restrict__ Minv) {
    extern __shared__ __align__(16) double tr_smem[];
    double* sB = tr_smem;                   // [64][132]
    double* sM = tr_smem + TR_ROWS * TR_LD;  // [128][132]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* Bg = Bp + (int64_t)blockIdx.x * TR_ROWS * ldb;

    for (int c = tid; c < TR_ROWS * (LF / 2); c += 256) {
        int r = c >> 6, kc = (c & 63) * 2;
        cp_async16(sB + r * TR_LD + kc, Bg + (int64_t)r * ldb + kc);
    }
    for (int c = tid; c < LF * (LF / 2); c += 256) {
        int r = c >> 6, kc = (c & 63) * 2;
        cp_async16(sM + r * TR_LD + kc, Minv + r * LF + kc);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    const int wm = (warp >> 2) * 32, wn = (warp & 3) * 32;
    const int lq = lane >> 2, lr = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    const int ksteps = (wn + 32) / 4;  // Minv[j][k] == 0 for k > j, j < wn + 32
    for (int ks = 0; ks < ksteps; ++ks) {
        const int kk = ks * 4 + lr;
        double af[4], bf[4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) af[mi] = sB[(wm + mi * 8 + lq) * TR_LD + kk];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) bf[ni] = sM[(wn + ni * 8 + lq) * TR_LD + kk];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
            *reinterpret_cast<double2*>(Bg + (int64_t)(wm + mi * 8 + lq) * ldb + wn + ni * 8 + 2 * lr) =
                make_double2(acc[mi][ni][0], acc[mi][ni][1]);
}

__global__ void copy_block_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int64_t ldd,
                                  int64_t rows, int64_t cols) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < rows * cols;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / cols, c = e % cols;
        dst[r * ldd + c] = src[r * lds + c];
    }
}


// Symmetrise: copy the lower triangle of W into the upper one, 32 x 32 tiles through shared memory.
__global__ void symmetrize_kernel(double* __restrict__ W, int64_t ldw) {
    __shared__ double t[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int r = ty; r < 32; r += 8) t[r][tx] = W[((int64_t)bi * 32 + r) * ldw + bj * 32 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t row = (int64_t)bj * 32 + r, col = (int64_t)bi * 32 + tx;  // transposed position
        if (bi != bj || col > row) W[row * ldw + col] = t[tx][r];
    }
}

// Gout (np x np, padded) = Gin (n x n, arbitrary ld) + diag_add * I, identity in the padding; only lower
// 128-tiles are written (what the factorisation reads).  For callers that bring their own Gram matrix
// (gym_socks/algorithms/kernel.py:43-66 ConditionalEmbedding.fit).
__global__ void pad_spd_kernel(const double* __restrict__ Gin, int64_t n, int64_t ldin, double diag_add,
                               double* __restrict__ Gout, int64_t ldout, int64_t np) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j >= np || j / TILE > i / TILE) return;
    double v = (i == j) ? 1.0 : 0.0;
    if (i < n && j < n) v = Gin[i * ldin + j] + (i == j ? diag_add : 0.0);
    Gout[i * ldout + j] = v;
}

struct CholCtx {
    cudaStream_t st;
    cudaError_t err = cudaSuccess;
    int rc = SOCKS_OK;  // first failing status of a library call made on this stream (its message is already set)
    void note(cudaError_t e) {
        if (err == cudaSuccess && e != cudaSuccess) err = e;
    }
    void note_rc(int r) {
        if (rc == SOCKS_OK && r != SOCKS_OK) rc = r;
    }
};

static inline int64_t split_point(int64_t n) { return ((n / LF) / 2) * LF; }  // n > 128, multiple of 128

// B (m x n, ld) <- B inv(L)^T, L = n x n lower at `L` (ld ldl); dinv indexes inv(L_kk) by diagonal block.
static void trsm_rec(CholCtx& cx, double* B, int64_t ldb, int64_t m, const double* L, int64_t ldl, int64_t n,
                     const double* dinv) {
    if (m == 0) return;
    if (n == LF) {
        trsm_leaf_kernel<<<(unsigned)(m / TR_ROWS), 256, TR_SMEM, cx.st>>>(B, ldb, dinv);
        cx.note(cudaGetLastError());
        return;
    }
    const int64_t n1 = split_point(n), n2 = n - n1;
    trsm_rec(cx, B, ldb, m, L, ldl, n1, dinv);
    // B2 -= B1 * L21^T
    cx.note(gemm_f64<false, true>(B, ldb, L + n1 * ldl, ldl, B + n1, ldb, m, n2, n1, -1.0, 1.0, 0, 0, cx.st));
    trsm_rec(cx, B + n1, ldb, m, L + n1 * ldl + n1, ldl, n2, dinv + (n1 / LF) * LF * LF);
}

static void potrf_rec(CholCtx& cx, double* A, int64_t lda, int64_t n, double* dinv, int* info, int64_t off) {
    if (n == LF) {
        potf2_inv_kernel<false><<<1, LF_THREADS, LF_SMEM, cx.st>>>(A, lda, dinv, LF, info, (int)off, nullptr);
        cx.note(cudaGetLastError());
        return;
    }
    const int64_t n1 = split_point(n), n2 = n - n1;
    double* A21 = A + n1 * lda;
    double* A22 = A21 + n1;
    potrf_rec(cx, A, lda, n1, dinv, info, off);
    trsm_rec(cx, A21, lda, n2, A, lda, n1, dinv);
    cx.note(gemm_f64<false, true>(A21, lda, A21, lda, A22, lda, n2, n2, n1, -1.0, 1.0, 0, 1, cx.st));
    potrf_rec(cx, A22, lda, n2, dinv + (n1 / LF) * LF * LF, info, off + n1);
}

static void trtri_rec(CholCtx& cx, const double* L, int64_t ldl, const double* dinv, double* M, int64_t ldm,
                      int64_t n, double* work) {
    if (n == LF) {
        copy_block_kernel<<<16, 256, 0, cx.st>>>(dinv, LF, M, ldm, LF, LF);
        cx.note(cudaGetLastError());
        return;
    }
    const int64_t n1 = split_point(n), n2 = n - n1;
    trtri_rec(cx, L, ldl, dinv, M, ldm, n1, work);
    trtri_rec(cx, L + n1 * ldl + n1, ldl, dinv + (n1 / LF) * LF * LF, M + n1 * ldm + n1, ldm, n2, work);
    // T (n2 x n1, ld n1) = M22 * L21 ; M21 = -T * M11
    cx.note(gemm_f64<false, false>(M + n1 * ldm + n1, ldm, L + n1 * ldl, ldl, work, n1, n2, n1, n2, 1.0, 0.0,
                                   GEMM_A_K_LE_M, 0, cx.st));
    cx.note(gemm_f64<false, false>(work, n1, M, ldm, M + n1 * ldm, ldm, n2, n1, n1, -1.0, 0.0, GEMM_B_K_GE_N, 0,
                                   cx.st));
}

// Factor AND invert in one recursion:  A -> L (lower, optional),  M <- inv(L).
//   (L11, M11) = rec(A11);  L21 = A21 M11^T (one triangular GEMM: the TRSM against an already inverted factor);
//   A22 -= L21 L21^T;  (L22, M22) = rec(A22);  M21 = -(M22 L21) M11.
// L21 lives in M's (2,1) block until M21 overwrites it, so the panel solve needs no leaf TRSM kernels, no
// recursive splitting of the solve and no copies: every flop outside the 128 x 128 leaves is a large GEMM.
static void potrf_inv_rec(CholCtx& cx, double* A, int64_t lda, double* M, int64_t ldm, int64_t n, double* work,
                          int* info, int64_t off, bool keep_l) {
    if (n == LF) {
        potf2_inv_kernel<false><<<1, LF_THREADS, LF_SMEM, cx.st>>>(A, lda, M, ldm, info, (int)off, nullptr);
        cx.note(cudaGetLastError());
        return;
    }
    const int64_t n1 = split_point(n), n2 = n - n1;
    double* A21 = A + n1 * lda;
    double* A22 = A21 + n1;
    double* M21 = M + n1 * ldm;  // holds L21 until the last step
    double* M22 = M21 + n1;
    potrf_inv_rec(cx, A, lda, M, ldm, n1, work, info, off, keep_l);
    // L21 = A21 * M11^T:  B(k, j) = M11[j][k], zero for k > j
    cx.note(gemm_f64<false, true>(A21, lda, M, ldm, M21, ldm, n2, n1, n1, 1.0, 0.0, GEMM_B_K_LE_N, 0, cx.st));
    cx.note(gemm_f64<false, true>(M21, ldm, M21, ldm, A22, lda, n2, n2, n1, -1.0, 1.0, 0, 1, cx.st));
    if (keep_l) {
        copy_block_kernel<<<(unsigned)std::min<int64_t>(4096, ceil_div(n2 * n1, 256)), 256, 0, cx.st>>>(M21, ldm, A21, lda,
                                                                                                      n2, n1);
        cx.note(cudaGetLastError());
    }
    potrf_inv_rec(cx, A22, lda, M22, ldm, n2, work, info, off + n1, keep_l);
    // T (n2 x n1, ld n1) = M22 * L21 ;  M21 = -T * M11
    cx.note(gemm_f64<false, false>(M22, ldm, M21, ldm, work, n1, n2, n1, n2, 1.0, 0.0, GEMM_A_K_LE_M, 0, cx.st));
    cx.note(gemm_f64<false, false>(work, n1, M, ldm, M21, ldm, n2, n1, n1, -1.0, 0.0, GEMM_B_K_GE_N, 0, cx.st));
}

// ---- look-ahead variant (what socks_potrf_inv runs) -------------------------------------------------------
// The pure recursion above serialises ~N/128 single-CTA leaves (64 us each, 147 SMs idle) and the tiny GEMMs of its
// bottom levels behind the large updates: 82 % of the DMMA peak at N = 20 000 while the GEMMs alone reach 95 %
// (profiles/r1_bench_launch_summary.txt).  Here the factorisation is right-looking over NB-wide block columns with a
// look-ahead of one: the trailing update of step k is split into the next block column (done first) and the rest, and
// the whole critical path of step k+1 -- factor+invert of the NB x NB diagonal block (the recursion above, on that
// block only) and the panel solve L21 = A21 inv(L_kk)^T -- runs on a second, high-priority stream while the rest of
// step k's update keeps the SMs busy.  inv(L) is then completed by the same block recursion as before
// (M21 = -(M22 L21) M11, large GEMMs only).  Same arithmetic per block as the pure recursion, different association
// across blocks, i.e. results differ at the rounding level (~cond * eps; tests compare both).
struct Lookahead {
    // panel stream (greatest priority), update stream (middle), merge stream (least): the block scheduler serves pending
    // CTAs in that order, so the merges of inv(L) only fill what the factorisation leaves idle
    cudaStream_t s2 = nullptr, su = nullptr, s3 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_panel = nullptr, ev_col = nullptr, ev_su = nullptr, ev_s3 = nullptr;
    std::vector<cudaEvent_t> ev_step;  // panel stream after step k (block, panel, small merges)
    bool ready = false;
};

static Lookahead* lookahead_for_device(int* rc_out) {
    static Lookahead la[SOCKS_MAX_DEVICES];
    static std::mutex mu;
    *rc_out = SOCKS_OK;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= SOCKS_MAX_DEVICES) {
        *rc_out = SOCKS_ECUDA;
        return nullptr;
    }
    std::lock_guard<std::mutex> lock(mu);
    Lookahead& l = la[dev];
    if (!l.ready) {
        int lo = 0, hi = 0;
        cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = greatest priority (numerically lowest)
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&l.s2, cudaStreamNonBlocking, hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&l.su, cudaStreamNonBlocking, (lo + hi) / 2);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&l.s3, cudaStreamNonBlocking, lo);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&l.ev_su, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&l.ev_s3, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&l.ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&l.ev_panel, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&l.ev_col, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            set_error("socks_potrf_inv: creating the look-ahead stream failed: %s", cudaGetErrorString(e));
            *rc_out = SOCKS_ECUDA;
            return nullptr;
        }
        l.ready = true;
    }
    return &l;
}

// ---- optional timeline of one socks_potrf_inv call (tools/potrf_trace.py): timed events on both streams ----------------
struct PotrfTrace {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> code;  // 1000 * kind + step; kinds: 0 start, 1 block done (s2), 2 panel done (s2), 3 small merges done (s2),
                            // 4 column update done (s1), 5 rest update done (s1), 6 join, 7 large merge done (s1)
    void mark(int kind, int step, cudaStream_t st) {
        if (!on) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, st);
        ev.push_back(e);
        code.push_back(1000 * kind + step);
    }
};
static PotrfTrace g_trace;

static int g_potrf_ozaki = 1;  // 1: trailing updates on the integer tensor cores (csrc/ozaki.cu), 0: FP64 DMMA GEMM
// Digits of the merges: the entries of inv(L) span many orders of magnitude WITHIN a row, and the digit scheme truncates
// relative to the row maximum; with 7 digits the residual |W G - I| rose from 3e-9 (DMMA) to 4e-7 at N = 20k, with 8 it is
// back at the DMMA level (tools/inverse_residual.py).  The trailing updates (rows of L21: narrow range) keep 7.
static int g_potrf_kbalance = 1;  // balance the two operands of a merge product along K (csrc/ozaki.cu: oz_kmax_kernel)
static int g_potrf_merge_digits = 8;
static int g_potrf_update_digits = 7;  // digits of the trailing updates (7 or 8)
static int g_potrf_ozaki_merge = 1;  // 1: the large merges of inv(L) on the integer tensor cores as well
static int g_potrf_tail = 1;   // 1: half-width block columns over the last 4 NB columns
static int g_potrf_nb = 1024;  // block-column width of the look-ahead factorisation (multiple of 128; 0 = pure recursion)

// inv(L) from its diagonal-block inverses: the merges M21 = -(M22 L21) M11 of the block recursion, in post-order
// (children before parents).  A merge over blocks [lo, hi) is ready as soon as block hi-1 has been factored, reads L
// only in rows [mid, hi) x columns [lo, mid) -- which no later trailing update reads -- and overwrites exactly that.
struct Merge {
    int lo, mid, hi;
};
static void plan_merges(int lo, int hi, std::vector<Merge>& out) {
    if (hi - lo <= 1) return;
    const int mid = lo + (hi - lo) / 2;
    plan_merges(lo, mid, out);
    plan_merges(mid, hi, out);
    out.push_back(Merge{lo, mid, hi});
}
// One merge M21 = -M22 L21 M11 as two products, associated so that the first needs nothing from the right half:
//   product 1:  T (n2 x n1, ld n1) = L21 M11    -- ready once blocks [lo, mid) are factored and M11 is complete
//   product 2:  M21 = -M22 T                    -- ready once M22 is complete as well; overwrites L21
// digits == nullptr: FP64 DMMA GEMMs.  Otherwise the product runs on the integer tensor cores (csrc/ozaki.cu) when its
// digit operands fit in `digit_bytes`; the triangular structure is clipped exactly as the DMMA GEMM clips it.
static void merge_product(CholCtx& cx, int which, double* M, int64_t ldm, const int64_t* bnd, const Merge& g, double* T,
                          uint8_t* digits, int64_t digit_bytes) {
    const int64_t o1 = bnd[g.lo], o2 = bnd[g.mid], n1 = bnd[g.mid] - bnd[g.lo], n2 = bnd[g.hi] - bnd[g.mid];
    double* M11 = M + o1 * ldm + o1;
    double* M21 = M + o2 * ldm + o1;  // holds L21 until product 2
    double* M22 = M + o2 * ldm + o2;
    const int S = g_potrf_merge_digits;
    const int64_t K = (which == 1) ? n1 : n2;
    if (digits && ozaki_operand_bytes(n2, K, S) + ozaki_operand_bytes(n1, K, S) <= digit_bytes) {
        int rc = SOCKS_OK;
        uint8_t* da = digits;
        uint8_t* db = digits + ozaki_operand_bytes(n2, K, S);
        // per-k balancing scales, after the digit operands (2 K doubles; digit_bytes leaves 1 KB + this much spare)
        double* kw = g_potrf_kbalance ? reinterpret_cast<double*>(digits + digit_bytes) : nullptr;
        const double *ksa = nullptr, *ksb = nullptr;
        if (which == 1) {  // A = rows of L21, B = columns of M11 (zero above the diagonal: k >= j), K = n1
            if (kw) {
                rc = ozaki_kbalance(M21, n2, ldm, 0, 0, M11, n1, ldm, 1, 2, n1, kw, cx.st);
                ksa = kw;
                ksb = kw + n1;
            }
            if (!rc) rc = ozaki_slice_operand(M21, n2, n1, ldm, 0, 0, da, S, ksa, cx.st);
            if (!rc) rc = ozaki_slice_operand(M11, n1, n1, ldm, 1, 2, db, S, ksb, cx.st);
            if (!rc) rc = ozaki_gemm_nt_sliced(da, n2, 0, db, n1, 0, n1, T, n1, n2, n1, 0, 1, 0, 1, S, cx.st);
        } else {  // A = rows of M22 (lower triangular: k <= i), B = columns of T, K = n2
            if (kw) {
                rc = ozaki_kbalance(M22, n2, ldm, 0, 1, T, n1, n1, 1, 0, n2, kw, cx.st);
                ksa = kw;
                ksb = kw + n2;
            }
            if (!rc) rc = ozaki_slice_operand(M22, n2, n2, ldm, 0, 1, da, S, ksa, cx.st);
            if (!rc) rc = ozaki_slice_operand(T, n1, n2, n1, 1, 0, db, S, ksb, cx.st);
            if (!rc) rc = ozaki_gemm_nt_sliced(da, n2, 0, db, n1, 0, n2, M21, ldm, n2, n1, 0, 2, 1, 0, S, cx.st);
        }
        cx.note_rc(rc);
        return;
    }
    if (which == 1)
        cx.note(gemm_f64<false, false>(M21, ldm, M11, ldm, T, n1, n2, n1, n1, 1.0, 0.0, GEMM_B_K_GE_N, 0, cx.st));
    else
        cx.note(gemm_f64<false, false>(M22, ldm, T, n1, M21, ldm, n2, n1, n2, -1.0, 0.0, GEMM_A_K_LE_M, 0, cx.st));
}

static inline int64_t trtri_scratch_doubles(int64_t np) {
    if (np <= LF) return 2;
    const int64_t n1 = split_point(np);
    return n1 * (np - n1);
}
// digits of one block-column panel (np x NB) for the integer tensor-core trailing updates, after the merge scratch
// Workspace of the look-ahead factorisation, in this order:
//   [scratch of the diagonal-block recursion and of the small merges | one T per large merge | panel digits | merge digits]
// T arena: every pair of distinct block columns meets in exactly one merge, so sum n1 n2 < np^2 / 2.
constexpr int64_t kRecScratchDoubles = (int64_t)4096 * 4096;  // merges spanning <= 4 block columns of <= 2048
static inline int64_t potrf_arena_doubles(int64_t np) { return np * np / 2; }
static inline int64_t potrf_panel_digit_bytes(int64_t np) { return ozaki_operand_bytes(np, 2048, 8) + 1024; }
// the two digit operands of the largest product (rows n1 + n2 <= np, K <= max(n1, n2) <= np/2 + 2 block columns)
static inline int64_t potrf_merge_digit_bytes(int64_t np) {
    const int64_t kmax = std::min<int64_t>(np, np / 2 + 2048);
    return 2 * ozaki_operand_bytes((np + 1) / 2 + 2048, kmax, 8) + 1024 + 2 * np * 8;  // + the K-balancing scales
}
static inline int64_t potrf_lookahead_bytes(int64_t np) {
    return (kRecScratchDoubles + potrf_arena_doubles(np)) * 8 + potrf_panel_digit_bytes(np) + potrf_merge_digit_bytes(np) + 2048;
}

static int potrf_inv_lookahead(double* A, int64_t lda, double* M, int64_t ldm, int64_t np, double* work, int* info,
                               bool keep_l, cudaStream_t s1, int64_t NB) {
    int rc;
    Lookahead* la = lookahead_for_device(&rc);
    if (!la) return rc;
    // workspace (socks_trtri_work_bytes): recursion scratch | T arena | panel digits | merge digits
    double* arena = work + kRecScratchDoubles;
    uint8_t* digits = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(arena + potrf_arena_doubles(np)) + 1023) & ~(uintptr_t)1023);
    uint8_t* mdigits = digits + ((potrf_panel_digit_bytes(np) + 1023) & ~(int64_t)1023);
    const int64_t mdigit_bytes = potrf_merge_digit_bytes(np) - 1024 - 2 * np * 8;
    const bool use_ozaki = g_potrf_ozaki && NB <= 2048;
    cudaStream_t su = la->su, s3 = la->s3;
    CholCtx c1{su}, c2{la->s2}, c3{s3};
    // block-column boundaries: NB wide, but half as wide over the last 4 NB columns -- there the trailing update is
    // shorter than the critical path of a block (leaves + small GEMMs), so shorter blocks shorten what is exposed
    std::vector<int64_t> bnd;
    const int64_t tail = g_potrf_tail ? std::max<int64_t>(0, np - 4 * NB) : np;
    for (int64_t o = 0; o < np;) {
        bnd.push_back(o);
        o += (o >= tail && NB >= 2 * LF) ? NB / 2 : NB;
    }
    bnd.push_back(np);
    const int nb = (int)bnd.size() - 1;
    std::vector<Merge> merges;
    plan_merges(0, nb, merges);
    constexpr int kSmallMerge = 4;  // merges spanning <= 4 block columns ride on the panel stream
    // one T per large merge
    std::vector<int64_t> t_off(merges.size(), -1);
    int64_t used = 0;
    for (size_t i = 0; i < merges.size(); ++i) {
        const Merge& g = merges[i];
        if (g.hi - g.lo <= kSmallMerge) continue;
        t_off[i] = used;
        used += (bnd[g.mid] - bnd[g.lo]) * (bnd[g.hi] - bnd[g.mid]);
    }
    if (used > potrf_arena_doubles(np)) {
        set_error("socks_potrf_inv: merge arena too small (%lld > %lld doubles)", (long long)used,
                  (long long)potrf_arena_doubles(np));
        return SOCKS_EINVAL(6);
    }
    while ((int)la->ev_step.size() < nb) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return SOCKS_ECUDA;
        la->ev_step.push_back(e);
    }
    // fork: everything already enqueued on the caller's stream (the Gram build) precedes the work of all three streams
    g_trace.mark(0, 0, s1);
    c1.note(cudaEventRecord(la->ev_fork, s1));
    c1.note(cudaStreamWaitEvent(su, la->ev_fork, 0));
    c2.note(cudaStreamWaitEvent(la->s2, la->ev_fork, 0));
    c3.note(cudaStreamWaitEvent(s3, la->ev_fork, 0));
    for (int k = 0; k < nb; ++k) {
        const int64_t k0 = bnd[k], bk = bnd[k + 1] - k0, rem = np - k0 - bk;
        double* Akk = A + k0 * lda + k0;
        double* Mkk = M + k0 * ldm + k0;
        if (k > 0) c2.note(cudaStreamWaitEvent(la->s2, la->ev_col, 0));
        potrf_inv_rec(c2, Akk, lda, Mkk, ldm, bk, work, info, k0, keep_l);
        g_trace.mark(1, k, la->s2);
        if (rem > 0) {  // panel: L21 = A21 * inv(L_kk)^T, stored in M's (2,1) block until the inverse overwrites it
            double* A21 = A + (k0 + bk) * lda + k0;
            double* L21 = M + (k0 + bk) * ldm + k0;
            c2.note(gemm_f64<false, true>(A21, lda, Mkk, ldm, L21, ldm, rem, bk, bk, 1.0, 0.0, GEMM_B_K_LE_N, 0, c2.st));
            if (keep_l) {
                copy_block_kernel<<<(unsigned)std::min<int64_t>(4096, ceil_div(rem * bk, 256)), 256, 0, c2.st>>>(
                    L21, ldm, A21, lda, rem, bk);
                c2.note(cudaGetLastError());
            }
        }
        g_trace.mark(2, k, la->s2);
        c2.note(cudaEventRecord(la->ev_panel, la->s2));
        c1.note(cudaStreamWaitEvent(su, la->ev_panel, 0));
        // small merges of the inverse that end at this block: on the panel stream, in the shadow of the column update
        for (const Merge& g : merges)
            if (g.hi == k + 1 && g.hi - g.lo <= kSmallMerge) {
                merge_product(c2, 1, M, ldm, bnd.data(), g, work, nullptr, 0);
                merge_product(c2, 2, M, ldm, bnd.data(), g, work, nullptr, 0);
            }
        g_trace.mark(3, k, la->s2);
        // large merges: whatever became ready with this step goes to the merge stream (children before parents: post-order)
        c2.note(cudaEventRecord(la->ev_step[k], la->s2));
        bool waited = false;
        for (size_t i = 0; i < merges.size(); ++i) {
            const Merge& g = merges[i];
            if (t_off[i] < 0) continue;
            for (int which = 1; which <= 2; ++which) {
                if ((which == 1 ? g.mid : g.hi) != k + 1) continue;
                if (!waited) {
                    c3.note(cudaStreamWaitEvent(s3, la->ev_step[k], 0));
                    waited = true;
                }
                merge_product(c3, which, M, ldm, bnd.data(), g, arena + t_off[i],
                              (use_ozaki && g_potrf_ozaki_merge) ? mdigits : nullptr, mdigit_bytes);
                g_trace.mark(7, 100 * which + (g.hi - g.lo), s3);
            }
        }
        if (rem == 0) break;
        const int64_t b1 = bnd[k + 2] - bnd[k + 1];
        const double* L21 = M + (k0 + bk) * ldm + k0;
        double* A22 = A + (k0 + bk) * lda + (k0 + bk);
        // next block column first (rows rem x cols b1), then the rest of the trailing lower triangle
        if (use_ozaki) {
            // A22 -= L21 L21^T on the integer tensor cores: the panel is split into digits once, both updates read them
            const int SU = g_potrf_update_digits;
            if ((rc = ozaki_slice_operand(L21, rem, bk, ldm, 0, 0, digits, SU, nullptr, su))) return rc;
            if ((rc = ozaki_gemm_nt_sliced(digits, rem, 0, digits, rem, 0, bk, A22, lda, rem, b1, 1, 0, 0, 0, SU, su))) return rc;
            g_trace.mark(4, k, su);
            c1.note(cudaEventRecord(la->ev_col, su));
            if (rem > b1 && (rc = ozaki_gemm_nt_sliced(digits, rem, b1 / LF, digits, rem, b1 / LF, bk, A22 + b1 * lda + b1,
                                                       lda, rem - b1, rem - b1, 1, 0, 0, 0, SU, su)))
                return rc;
            g_trace.mark(5, k, su);
        } else {
            c1.note(gemm_f64<false, true>(L21, ldm, L21, ldm, A22, lda, rem, b1, bk, -1.0, 1.0, 0, 1, su));
            c1.note(cudaEventRecord(la->ev_col, su));
            if (rem > b1)
                c1.note(gemm_f64<false, true>(L21 + b1 * ldm, ldm, L21 + b1 * ldm, ldm, A22 + b1 * lda + b1, lda, rem - b1,
                                              rem - b1, bk, -1.0, 1.0, 0, 1, su));
        }
    }
    // join all three streams back into the caller's
    c2.note(cudaEventRecord(la->ev_panel, la->s2));
    c1.note(cudaEventRecord(la->ev_su, su));
    c3.note(cudaEventRecord(la->ev_s3, s3));
    cudaError_t je = cudaStreamWaitEvent(s1, la->ev_panel, 0);
    if (je == cudaSuccess) je = cudaStreamWaitEvent(s1, la->ev_su, 0);
    if (je == cudaSuccess) je = cudaStreamWaitEvent(s1, la->ev_s3, 0);
    g_trace.mark(6, 0, s1);
    if (c2.rc || c3.rc) return c2.rc ? c2.rc : c3.rc;  // an integer-GEMM call failed: its message stands
    cudaError_t err = (c1.err != cudaSuccess) ? c1.err : ((c2.err != cudaSuccess) ? c2.err : c3.err);
    if (err == cudaSuccess) err = je;
    if (err != cudaSuccess) {
        set_error("socks_potrf_inv: CUDA call failed: %s", cudaGetErrorString(err));
        return SOCKS_ECUDA;
    }
    return SOCKS_OK;
}

static int chol_configure() {
    static std::atomic<bool> done[SOCKS_MAX_DEVICES];  // dynamic-smem opt-in is per device
    int dev = 0;
    SOCKS_CUDA(cudaGetDevice(&dev));
    const bool tracked = dev >= 0 && dev < SOCKS_MAX_DEVICES;
    if (tracked && done[dev].load(std::memory_order_acquire)) return SOCKS_OK;
    SOCKS_CUDA(cudaFuncSetAttribute(potf2_inv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LF_SMEM));
    SOCKS_CUDA(cudaFuncSetAttribute(potf2_inv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LF_SMEM));
    SOCKS_CUDA(cudaFuncSetAttribute(trsm_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TR_SMEM));
    if (tracked) done[dev].store(true, std::memory_order_release);
    return SOCKS_OK;
}

}  // namespace socks

using namespace socks;

extern "C" int64_t socks_potrf_work_bytes(int64_t n) { return socks_padded_dim(n) * LF * 8; }

extern "C" int64_t socks_trtri_work_bytes(int64_t n) {
    const int64_t np = socks_padded_dim(n);
    if (np <= LF) return 16;
    // pure recursion: the merge scratch T (n2 x n1).  Look-ahead factorisation (more than 2 block columns, the same
    // test as in socks_potrf_inv): recursion scratch, one T per large merge and the digit operands of the integer GEMMs.
    const bool lookahead = g_potrf_nb > 0 && np > 2 * g_potrf_nb;
    return lookahead ? std::max<int64_t>(trtri_scratch_doubles(np) * 8, potrf_lookahead_bytes(np))
                     : trtri_scratch_doubles(np) * 8;
}

extern "C" int64_t socks_reduce_work_bytes(int64_t cols) { return 2 * CR_MAX_SPLITS * colreduce_ldpart(cols) * 8; }

static int check_square(const void* p, int64_t n, int64_t ld, int argp) {
    SOCKS_CHECK_ARG(p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15) == 0, argp);
    SOCKS_CHECK_ARG(n >= 1, argp + 1);
    SOCKS_CHECK_ARG(ld >= socks_padded_dim(n) && ld % 2 == 0, argp + 2);
    return SOCKS_OK;
}

extern "C" int socks_potrf(double* G, int64_t n, int64_t ldg, double* dinv, int* info_dev, socks_stream_t stream) {
    int rc = check_square(G, n, ldg, 1);
    if (rc) return rc;
    SOCKS_CHECK_ARG(dinv != nullptr && (reinterpret_cast<uintptr_t>(dinv) & 15) == 0, 4);
    SOCKS_CHECK_ARG(info_dev != nullptr, 5);
    if ((rc = chol_configure())) return rc;
    CholCtx cx{(cudaStream_t)stream};
    potrf_rec(cx, G, ldg, socks_padded_dim(n), dinv, info_dev, 0);
    if (cx.err != cudaSuccess) {
        set_error("socks_potrf: CUDA launch failed: %s", cudaGetErrorString(cx.err));
        return SOCKS_ECUDA;
    }
    return SOCKS_OK;
}

// Development helper: factor ONE 128 x 128 block (A, lda) and record clock64() at the phase boundaries of the
// leaf kernel into cycles[0..14] (device buffer of 16 int64).
extern "C" int socks_dev_potf2_phases(double* A, int64_t lda, double* dinv, int* info_dev, long long* cycles,
                                      socks_stream_t stream) {
    SOCKS_CHECK_ARG(A != nullptr && dinv != nullptr && info_dev != nullptr && cycles != nullptr, 1);
    int rc = chol_configure();
    if (rc) return rc;
    potf2_inv_kernel<true><<<1, LF_THREADS, LF_SMEM, (cudaStream_t)stream>>>(A, lda, dinv, LF, info_dev, 0, cycles);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

extern "C" int socks_trtri(const double* L, int64_t n, int64_t ldl, const double* dinv, double* M, int64_t ldm,
                           double* work, socks_stream_t stream) {
    int rc = check_square(L, n, ldl, 1);
    if (rc) return rc;
    SOCKS_CHECK_ARG(dinv != nullptr, 4);
    if ((rc = check_square(M, n, ldm, 5))) return rc;
    SOCKS_CHECK_ARG(work != nullptr && (reinterpret_cast<uintptr_t>(work) & 15) == 0, 7);
    CholCtx cx{(cudaStream_t)stream};
    trtri_rec(cx, L, ldl, dinv, M, ldm, socks_padded_dim(n), work);
    if (cx.err != cudaSuccess) {
        set_error("socks_trtri: CUDA launch failed: %s", cudaGetErrorString(cx.err));
        return SOCKS_ECUDA;
    }
    return SOCKS_OK;
}

extern "C" int socks_pad_spd(const double* Gin, int64_t n, int64_t ldin, double diag_add, double* Gout, int64_t ldout,
                             socks_stream_t stream) {
    SOCKS_CHECK_ARG(Gin != nullptr, 1);
    SOCKS_CHECK_ARG(n >= 1 && socks_padded_dim(n) <= 65535, 2);
    SOCKS_CHECK_ARG(ldin >= n, 3);
    int rc = check_square(Gout, n, ldout, 5);
    if (rc) return rc > 0 ? rc : (rc == -6 ? -2 : rc);
    const int64_t np = socks_padded_dim(n);
    pad_spd_kernel<<<dim3((unsigned)ceil_div(np, 256), (unsigned)np), 256, 0, (cudaStream_t)stream>>>(
        Gin, n, ldin, diag_add, Gout, ldout, np);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

extern "C" int socks_potrf_inv(double* G, int64_t n, int64_t ldg, double* M, int64_t ldm, double* work, int* info_dev,
                               int keep_l, socks_stream_t stream) {
    int rc = check_square(G, n, ldg, 1);
    if (rc) return rc;
    if ((rc = check_square(M, n, ldm, 4))) return rc;
    SOCKS_CHECK_ARG(work != nullptr && (reinterpret_cast<uintptr_t>(work) & 15) == 0, 6);
    SOCKS_CHECK_ARG(info_dev != nullptr, 7);
    if ((rc = chol_configure())) return rc;
    const int64_t np = socks_padded_dim(n);
    if (g_potrf_nb > 0 && np > 2 * g_potrf_nb)
        return potrf_inv_lookahead(G, ldg, M, ldm, np, work, info_dev, keep_l != 0, (cudaStream_t)stream, g_potrf_nb);
    CholCtx cx{(cudaStream_t)stream};
    potrf_inv_rec(cx, G, ldg, M, ldm, np, work, info_dev, 0, keep_l != 0);
    if (cx.err != cudaSuccess) {
        set_error("socks_potrf_inv: CUDA launch failed: %s", cudaGetErrorString(cx.err));
        return SOCKS_ECUDA;
    }
    return SOCKS_OK;
}

// Block-column width of socks_potrf_inv's look-ahead factorisation (multiple of 128); 0 selects the pure recursion.
// Returns the previous value.  Tuning / measurement knob (tools/run_potrf.py); not thread-safe against running calls.
extern "C" int socks_potrf_inv_set_block(int nb) {
    const int old = g_potrf_nb;
    if (nb < 0) {  // -1 / -2: half-width tail off / on; -3 / -4: FP64 DMMA / integer tensor-core trailing updates;
                   // -5 / -6: large merges of the inverse on FP64 DMMA / on the integer tensor cores
        if (nb >= -2) g_potrf_tail = (nb == -2);
        else if (nb >= -4) g_potrf_ozaki = (nb == -4);
        else if (nb >= -6) g_potrf_ozaki_merge = (nb == -6);
        else if (nb >= -8) g_potrf_merge_digits = (nb == -7) ? 7 : 8;  // -7 / -8: 7 / 8 digits in the merges
        else if (nb >= -10) g_potrf_update_digits = (nb == -9) ? 7 : 8;  // -9 / -10: 7 / 8 digits in the trailing updates
        else g_potrf_kbalance = (nb == -12);                             // -11 / -12: K-balancing of the merge products off / on
        return old;
    }
    if (nb % LF == 0 && nb <= 2048) g_potrf_nb = nb;  // the recursion scratch is sized for block columns of <= 2048
    return old;
}

// Timeline of the NEXT socks_potrf_inv call: socks_dev_potrf_trace(nullptr, nullptr, 0) arms it; after the call (and a
// stream synchronise) socks_dev_potrf_trace(codes, ms, cap) returns the number of events and fills code = 1000 kind +
// step (kinds in PotrfTrace) and the time in ms since the start event.
extern "C" int socks_dev_potrf_trace(int* codes, double* ms, int cap) {
    if (!codes || !ms) {
        for (cudaEvent_t e : g_trace.ev) cudaEventDestroy(e);
        g_trace.ev.clear();
        g_trace.code.clear();
        g_trace.on = true;
        return 0;
    }
    g_trace.on = false;
    const int n = std::min<int>(cap, (int)g_trace.ev.size());
    for (int i = 0; i < n; ++i) {
        float t = 0.f;
        cudaEventSynchronize(g_trace.ev[i]);
        cudaEventElapsedTime(&t, g_trace.ev[0], g_trace.ev[i]);
        codes[i] = g_trace.code[i];
        ms[i] = t;
    }
    return n;
}

extern "C" int socks_diag_inv(const double* M, int64_t n, int64_t ldm, double* w, double* work,
                              socks_stream_t stream) {
    int rc = check_square(M, n, ldm, 1);
    if (rc) return rc;
    SOCKS_CHECK_ARG(w != nullptr, 4);
    SOCKS_CHECK_ARG(work != nullptr, 5);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t np = socks_padded_dim(n);
    // rows of the padding are the identity: summing over np rows adds exact zeros for columns < n
    const int splits = colreduce_launch<true>(M, np, n, ldm, nullptr, 0, 1, work, /*tri=*/1, st);
    SOCKS_CHECK_LAUNCH();
    colreduce_finish_kernel<<<dim3((unsigned)ceil_div(n, 256), 1), 256, 0, st>>>(work, colreduce_ldpart(n), splits, n,
                                                                                w, n);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

// y[j] = sum_{i < n} Z[i][j]^2 for a general (non-triangular) n x m matrix.
extern "C" int socks_colsumsq(const double* Zm, int64_t n, int64_t m, int64_t ldz, double* y, double* work,
                              socks_stream_t stream) {
    SOCKS_CHECK_ARG(Zm != nullptr, 1);
    SOCKS_CHECK_ARG(n >= 1, 2);
    SOCKS_CHECK_ARG(m >= 1, 3);
    SOCKS_CHECK_ARG(ldz >= m, 4);
    SOCKS_CHECK_ARG(y != nullptr, 5);
    SOCKS_CHECK_ARG(work != nullptr, 6);
    cudaStream_t st = (cudaStream_t)stream;
    const int splits = colreduce_launch<true>(Zm, n, m, ldz, nullptr, 0, 1, work, /*tri=*/0, st);
    SOCKS_CHECK_LAUNCH();
    colreduce_finish_kernel<<<dim3((unsigned)ceil_div(m, 256), 1), 256, 0, st>>>(work, colreduce_ldpart(m), splits, m,
                                                                                y, m);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

// X = inv(G) B = M^T (M B) for nrhs right-hand sides (nrhs a multiple of 128, rows padded to np):
// two triangular DMMA products.  B (np x nrhs) is overwritten by X; work holds np x nrhs doubles.
extern "C" int socks_potrs(const double* M, int64_t n, int64_t ldm, double* B, int64_t nrhs, int64_t ldb, double* work,
                           socks_stream_t stream) {
    int rc = check_square(M, n, ldm, 1);
    if (rc) return rc;
    SOCKS_CHECK_ARG(B != nullptr && (reinterpret_cast<uintptr_t>(B) & 15) == 0, 4);
    SOCKS_CHECK_ARG(nrhs >= 128 && nrhs % 128 == 0, 5);
    SOCKS_CHECK_ARG(ldb >= nrhs && ldb % 2 == 0, 6);
    SOCKS_CHECK_ARG(work != nullptr && (reinterpret_cast<uintptr_t>(work) & 15) == 0, 7);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t np = socks_padded_dim(n);
    // Z = M B (M lower: k <= i), then X = M^T Z (A(i,k) = M[k][i], zero for k < i)
    cudaError_t e = gemm_f64<false, false>(M, ldm, B, ldb, work, nrhs, np, nrhs, np, 1.0, 0.0, GEMM_A_K_LE_M, 0, st);
    if (e == cudaSuccess)
        e = gemm_f64<true, false>(M, ldm, work, nrhs, B, ldb, np, nrhs, np, 1.0, 0.0, GEMM_A_K_GE_M, 0, st);
    if (e != cudaSuccess) {
        set_error("socks_potrs: CUDA launch failed: %s", cudaGetErrorString(e));
        return SOCKS_ECUDA;
    }
    return SOCKS_OK;
}

// The same product on the integer tensor cores: W = A A^T with A = M^T (operand rows = columns of M, zero for k before
// the row: both K clips apply), 8 digits (the columns of inv(L) span many orders of magnitude, see g_potrf_merge_digits).
extern "C" int64_t socks_lauum_i8_work_bytes(int64_t n) {
    const int64_t np = socks_padded_dim(n);
    return ozaki_operand_bytes(np, np, 8) + 2048;
}

extern "C" int socks_lauum_i8(const double* M, int64_t n, int64_t ldm, double* W, int64_t ldw, void* work,
                              socks_stream_t stream) {
    int rc = check_square(M, n, ldm, 1);
    if (rc) return rc;
    if ((rc = check_square(W, n, ldw, 4))) return rc;
    SOCKS_CHECK_ARG(work != nullptr, 6);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t np = socks_padded_dim(n);
    uint8_t* buf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(work) + 1023) & ~(uintptr_t)1023);
    if ((rc = ozaki_slice_operand(M, np, np, ldm, 1, 2, buf, 8, nullptr, st))) return rc;
    if ((rc = ozaki_gemm_nt_sliced(buf, np, 0, buf, np, 0, np, W, ldw, np, np, 1, 1, 2, 1, 8, st))) return rc;
    symmetrize_kernel<<<dim3((unsigned)(np / 32), (unsigned)(np / 32)), dim3(32, 8), 0, st>>>(W, ldw);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

extern "C" int socks_lauum(const double* M, int64_t n, int64_t ldm, double* W, int64_t ldw, socks_stream_t stream) {
    int rc = check_square(M, n, ldm, 1);
    if (rc) return rc;
    if ((rc = check_square(W, n, ldw, 4))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t np = socks_padded_dim(n);
    // W[i][j] = sum_{k >= max(i,j)} M[k][i] M[k][j], lower tiles; then mirror.
    cudaError_t e = gemm_f64<true, false>(M, ldm, M, ldm, W, ldw, np, np, np, 1.0, 0.0,
                                          GEMM_A_K_GE_M | GEMM_B_K_GE_N, 1, st);
    if (e != cudaSuccess) {
        set_error("socks_lauum: CUDA launch failed: %s", cudaGetErrorString(e));
        return SOCKS_ECUDA;
    }
    symmetrize_kernel<<<dim3((unsigned)(np / 32), (unsigned)(np / 32)), dim3(32, 8), 0, st>>>(W, ldw);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}
