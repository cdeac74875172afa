// Stochastic-reachability backward recursion (KernelSR) and the forward-control score kernels.
//
// Fit (gym_socks/algorithms/reach/kernel_sr.py:297-309): beta lives in HBM as the un-normalised
// B = diag(w) K(X,Y) (n x n) plus its column sums; each of the T-1 time steps is one coalesced
// GEMV pass over B (reduce.cuh) followed by a fused normalise + FHT/THT step (HBM-bound).
//
// Predict (kernel_sr.py:329-342): K(X,pts) is never written.  With the fitted value functions
// known, all T-1 steps collapse into ONE contraction
//     acc[r][j] = sum_i coef[i][r] k(x_i, pts_j),  coef[.][0] = w, coef[.][r] = vf[t0+r] * w
// evaluated on the fly: each lane computes the kernel value of one (sample, point) pair in
// registers (distance + table exp, exp_f64.cuh) and feeds it straight into DMMA.8x8x4 as the B fragment,
// the 16 coefficient rows being the A fragments.  On B200 DMMA and DFMA share the FP64 pipe, so the
// kernel is FP64-pipe bound: per pair 11 scalar instructions + 2/32 of two DMMA (= 16 FMA-equivalents);
// algorithmic work N*M*(3d+2+E+2T) flops with E = 12, 8*M*(d+T) bytes of I/O.
#include <climits>
#include <cmath>

#include "exp_f64.cuh"
#include "reduce.cuh"

namespace socks {

constexpr int SR_CP = 20;  // coef row stride (16 rows + pad; 20 mod 16 = 4 -> conflict-free A fragments)

// out[T-1][j] = 1_target[T-1](pts_j)   (kernel_sr.py:57)
__global__ void sr_init_kernel(const double* __restrict__ pts, int64_t m, int d, const double* __restrict__ tubes,
                               int T, double* __restrict__ out, int64_t ldout) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double* tube = tubes + (int64_t)(T - 1) * 4 * d;
    out[(int64_t)(T - 1) * ldout + j] = in_box(pts + j * d, tube, 1, d) ? 1.0 : 0.0;
}

// out_t[j] = step(pts_j, (sum_s part[s][j]) / den[j])       (kernel_sr.py:62-63 after `normalize`)
__global__ void sr_finish_step_kernel(const double* __restrict__ part, int64_t ldpart, int splits,
                                      const double* __restrict__ den, const double* __restrict__ pts, int64_t m, int d,
                                      const double* __restrict__ tube_t, int problem, double* __restrict__ out_t) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double a = 0.0;
    for (int s = 0; s < splits; ++s) a += part[(int64_t)s * ldpart + j];
    out_t[j] = sr_step(pts + j * d, tube_t, d, problem, a / den[j]);
}

// coef[i][0] = w_i, coef[i][r] = vf[t0 + r][i] * w_i (1 <= r < R), 0 for r >= R or i >= n.
__global__ void sr_pack_coef_kernel(const double* __restrict__ w, const double* __restrict__ vf, int64_t ldvf,
                                    int64_t n, int64_t n_pad, int t0, int R, double* __restrict__ coef) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const double wi = (i < n) ? w[i] : 0.0;
    double* c = coef + i * SR_CP;
    c[0] = wi;
    for (int r = 1; r < SR_CP; ++r) c[r] = (r < R && i < n) ? vf[(int64_t)(t0 + r) * ldvf + i] * wi : 0.0;
}

constexpr int SP_TS = 64;        // samples per shared-memory stage
constexpr int SP_WARPS = 8;
constexpr int SP_Q = 4;          // 8-point n-tiles per warp
constexpr int SP_PTS = SP_WARPS * SP_Q * 8;  // 256 points per CTA
constexpr int SP_SPLIT = 2048;   // samples per CTA (blockIdx.y): a FIXED split, so the summation order -- and
                                 // hence every output bit -- does not depend on M, chunking or sharding

template <int D>
__device__ __forceinline__ void sp_load_stage(double* __restrict__ sx, double* __restrict__ sc,
                                              const double* __restrict__ X, const double* __restrict__ coef, int64_t n,
                                              int64_t s0, int tid, int nthreads) {
    // coefficient rows: SP_TS x 20 doubles in 16-byte chunks, contiguous in global memory
    const double* csrc = coef + s0 * SR_CP;
    for (int c = tid; c < SP_TS * SR_CP / 2; c += nthreads) cp_async16(sc + c * 2, csrc + c * 2);
    for (int e = tid; e < SP_TS * D; e += nthreads) {
        const int64_t i = s0 + e / D;
        sx[e] = (i < n) ? __ldg(X + s0 * D + e) : 0.0;
    }
}

// variant 0: DMMA contraction.  CTA = 8 warps x 32 points x one 2048-sample split; each lane evaluates
// k(x_s, p) for sample s = 4*ks + lane%4 and the points {8q + lane/4}, q < 4, which is exactly the B-fragment
// layout of mma.m8n8k4 (B[k = lane%4][n = lane/4]).  Partial sums go to part[split][16][ldp]; the finish
// kernel adds the splits in order, divides by the normaliser row and applies the FHT/THT step.  Splitting
// the sample axis gives N/2048 times more CTAs than point tiles alone: no wave-quantisation tail.
template <int D, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    sr_predict_dmma_kernel(const double* __restrict__ X, const double* __restrict__ coef, int64_t n, int64_t n_pad,
                           const double* __restrict__ pts, int64_t m, ExpScaled ex, double* __restrict__ part,
                           int64_t ldp) {
    __shared__ __align__(16) double sx[2][SP_TS * D];
    __shared__ __align__(16) double sc[2][SP_TS * SR_CP];
    __shared__ double etbl[EXP_TBL2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lq = lane >> 2, lr = lane & 3;
    const int64_t pbase = (int64_t)blockIdx.x * (WARPS * SP_Q * 8) + warp * (SP_Q * 8);
    exp_table2_init(etbl, tid, WARPS * 32);  // visible after the first __syncthreads of the main loop

    double p[SP_Q][D];
#pragma unroll
    for (int q = 0; q < SP_Q; ++q) {
        const int64_t j = min(pbase + q * 8 + lq, m - 1);
#pragma unroll
        for (int k = 0; k < D; ++k) p[q][k] = __ldg(pts + j * D + k);
    }
    double acc[SP_Q][2][2];
#pragma unroll
    for (int q = 0; q < SP_Q; ++q)
#pragma unroll
        for (int h = 0; h < 2; ++h) acc[q][h][0] = acc[q][h][1] = 0.0;

    const int tile0 = blockIdx.y * (SP_SPLIT / SP_TS);
    const int tile1 = min((int)(n_pad / SP_TS), tile0 + SP_SPLIT / SP_TS);
    sp_load_stage<D>(sx[0], sc[0], X, coef, n, (int64_t)tile0 * SP_TS, tid, WARPS * 32);
    cp_async_commit();
    for (int tile = tile0; tile < tile1; ++tile) {
        const int cur = (tile - tile0) & 1;
        cp_async_wait<0>();
        __syncthreads();  // stage `cur` visible to all; everyone is done reading stage cur^1
        if (tile + 1 < tile1)
            sp_load_stage<D>(sx[cur ^ 1], sc[cur ^ 1], X, coef, n, (int64_t)(tile + 1) * SP_TS, tid, WARPS * 32);
        cp_async_commit();
        const double* x_s = sx[cur];
        const double* c_s = sc[cur];
#pragma unroll 4
        for (int ks = 0; ks < SP_TS / 4; ++ks) {
            const int s = ks * 4 + lr;
            double xs[D];
#pragma unroll
            for (int k = 0; k < D; ++k) xs[k] = x_s[s * D + k];
            const double a0 = c_s[s * SR_CP + lq];
            const double a1 = c_s[s * SR_CP + lq + 8];
            {
                double arg[SP_Q], kv[SP_Q];
#pragma unroll
                for (int q = 0; q < SP_Q; ++q) {
                    double d2 = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        const double t = p[q][k] - xs[k];
                        d2 = fma(t, t, d2);
                    }
                    arg[q] = d2;
                }
                exp_scaled_x4(arg, kv, etbl, ex);
#pragma unroll
                for (int q = 0; q < SP_Q; ++q) {
                    dmma884(acc[q][0][0], acc[q][0][1], a0, kv[q]);
                    dmma884(acc[q][1][0], acc[q][1][1], a1, kv[q]);
                }
            }
        }
    }

    // epilogue: accumulator rows r = lq (+8) of this split -> part[split][r][j], 16-byte stores
    double* pbase_out = part + (int64_t)blockIdx.y * 16 * ldp;
#pragma unroll
    for (int q = 0; q < SP_Q; ++q) {
        const int64_t j = pbase + q * 8 + 2 * lr;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double* dst = pbase_out + (int64_t)(lq + 8 * h) * ldp + j;
            if (j + 1 < m) {
                *reinterpret_cast<double2*>(dst) = make_double2(acc[q][h][0], acc[q][h][1]);
            } else if (j < m) {
                dst[0] = acc[q][h][0];
            }
        }
    }
}

// out[t0+r-1][j] = step(pts_j, (sum_s part[s][r][j]) / (sum_s part[s][0][j])),  1 <= r < R
__global__ void sr_predict_finish_kernel(const double* __restrict__ part, int64_t ldp, int splits, int64_t m, int R,
                                         int t0, const double* __restrict__ pts, int d,
                                         const double* __restrict__ tubes, int problem, double* __restrict__ out,
                                         int64_t ldout) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double den = 0.0;
    for (int s = 0; s < splits; ++s) den += part[(int64_t)s * 16 * ldp + j];
    for (int r = 1; r < R; ++r) {
        double num = 0.0;
        for (int s = 0; s < splits; ++s) num += part[((int64_t)s * 16 + r) * ldp + j];
        const int t = t0 + r - 1;
        out[(int64_t)t * ldout + j] = sr_step(pts + j * d, tubes + (int64_t)t * 4 * d, d, problem, num / den);
    }
}

// variant 1: FMA-pipe contraction, one point per thread (comparison / cross-check of variant 0).
template <int D>
__global__ void __launch_bounds__(128)
    sr_predict_fma_kernel(const double* __restrict__ X, const double* __restrict__ coef, int64_t n, int64_t n_pad,
                          const double* __restrict__ pts, int64_t m, double neg_gamma,
                          const double* __restrict__ tubes, int problem, int t0, int R, double* __restrict__ out,
                          int64_t ldout) {
    __shared__ __align__(16) double sx[2][SP_TS * D];
    __shared__ __align__(16) double sc[2][SP_TS * SR_CP];
    const int tid = threadIdx.x;
    const int64_t j = (int64_t)blockIdx.x * 128 + tid;
    const int64_t jc = min(j, m - 1);
    double p[D];
#pragma unroll
    for (int k = 0; k < D; ++k) p[k] = __ldg(pts + jc * D + k);
    double acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = 0.0;

    const int ntiles = (int)(n_pad / SP_TS);
    sp_load_stage<D>(sx[0], sc[0], X, coef, n, 0, tid, 128);
    cp_async_commit();
    for (int tile = 0; tile < ntiles; ++tile) {
        const int cur = tile & 1;
        cp_async_wait<0>();
        __syncthreads();
        if (tile + 1 < ntiles)
            sp_load_stage<D>(sx[cur ^ 1], sc[cur ^ 1], X, coef, n, (int64_t)(tile + 1) * SP_TS, tid, 128);
        cp_async_commit();
        const double* x_s = sx[cur];
        const double* c_s = sc[cur];
#pragma unroll 2
        for (int s = 0; s < SP_TS; ++s) {
            double d2 = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const double t = p[k] - x_s[s * D + k];
                d2 = fma(t, t, d2);
            }
            const double kv = exp(d2 * neg_gamma);
            const double2* c2 = reinterpret_cast<const double2*>(c_s + s * SR_CP);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double2 c = c2[r];
                acc[2 * r] = fma(c.x, kv, acc[2 * r]);
                acc[2 * r + 1] = fma(c.y, kv, acc[2 * r + 1]);
            }
        }
    }
    if (j < m) {
#pragma unroll
        for (int r = 1; r < 16; ++r)
            if (r < R) {
                const int t = t0 + r - 1;
                out[(int64_t)t * ldout + j] =
                    sr_step(pts + j * D, tubes + (int64_t)t * 4 * D, D, problem, acc[r] / acc[0]);
            }
    }
}

// Runtime-d fallback (8 < d <= SOCKS_MAX_DIM): one point per thread, coordinates through L1, samples and
// coefficients straight from global memory (warp-uniform addresses).  Correctness path for high-dimensional
// states; the benchmarked dimensions (d <= 8) use the templated kernels above.
__global__ void __launch_bounds__(128)
    sr_predict_generic_kernel(const double* __restrict__ X, const double* __restrict__ coef, int64_t n, int d,
                              const double* __restrict__ pts, int64_t m, double kmul, const double* __restrict__ tubes,
                              int problem, int t0, int R, double* __restrict__ out, int64_t ldout) {
    const int64_t j = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (j >= m) return;
    const double* p = pts + j * d;
    double acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double d2 = 0.0;
        for (int k = 0; k < d; ++k) {
            const double t = p[k] - __ldg(X + i * d + k);
            d2 = fma(t, t, d2);
        }
        const double kv = exp(d2 * kmul);
        const double* c = coef + i * SR_CP;
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r] = fma(__ldg(c + r), kv, acc[r]);
    }
#pragma unroll
    for (int r = 1; r < 16; ++r)
        if (r < R) {
            const int t = t0 + r - 1;
            out[(int64_t)t * ldout + j] = sr_step(p, tubes + (int64_t)t * 4 * d, d, problem, acc[r] / acc[0]);
        }
}

// q[r][i] = z[r][i] * exp(-gamma |X_i - s|^2)     (kernel_control_fwd.py:220-222: CXT * ...)
__global__ void fwd_weights_kernel(const double* __restrict__ X, int64_t n, int d, double neg_gamma,
                                   const double* __restrict__ state, const double* __restrict__ z, int64_t ldz, int R,
                                   double* __restrict__ q, int64_t ldq) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d2 = 0.0;
    for (int k = 0; k < d; ++k) {
        const double t = __ldg(state + k) - X[i * d + k];
        d2 = fma(t, t, d2);
    }
    const double kv = exp(d2 * neg_gamma);
    for (int r = 0; r < R; ++r) q[(int64_t)r * ldq + i] = z[(int64_t)r * ldz + i] * kv;
}

// First-occurrence argmin of C over {k : D[k] <= 0}; all k if D is NULL or no k is feasible
// (gym_socks/algorithms/control/common.py:96-100, 166-174).  NaN ordering follows numpy's argmin.
__device__ __forceinline__ bool argmin_better(double c, int k, double bc, int bk) {
    if (bk == INT_MAX) return true;
    const bool cn = (c != c), bn = (bc != bc);
    if (cn != bn) return cn;  // numpy's argmin returns the first NaN
    if (cn) return k < bk;
    if (c != bc) return c < bc;
    return k < bk;
}

__global__ void argmin_masked_kernel(const double* __restrict__ C, const double* __restrict__ Dm, int P,
                                     int* __restrict__ idx_out) {
    __shared__ double sval[2][256];
    __shared__ int sidx[2][256];
    const int tid = threadIdx.x;
    double best[2] = {0.0, 0.0};
    int bidx[2] = {INT_MAX, INT_MAX};
    for (int k = tid; k < P; k += blockDim.x) {
        const double c = C[k];
        if (argmin_better(c, k, best[0], bidx[0])) { best[0] = c; bidx[0] = k; }
        if (Dm && Dm[k] <= 0.0 && argmin_better(c, k, best[1], bidx[1])) { best[1] = c; bidx[1] = k; }
    }
    for (int s = 0; s < 2; ++s) { sval[s][tid] = best[s]; sidx[s][tid] = bidx[s]; }
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if (tid < off) {
            for (int s = 0; s < 2; ++s) {
                const int ix = sidx[s][tid + off];
                if (ix != INT_MAX && argmin_better(sval[s][tid + off], ix, sval[s][tid], sidx[s][tid])) {
                    sval[s][tid] = sval[s][tid + off];
                    sidx[s][tid] = ix;
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) *idx_out = (Dm && sidx[1][0] != INT_MAX) ? sidx[1][0] : sidx[0][0];
}

}  // namespace socks

using namespace socks;

extern "C" int socks_sr_recursion(const double* B, int64_t n, int64_t m, int64_t ldb, const double* pts, int d,
                                  const double* tubes, int problem, int T, const double* vf, int64_t ldvf, double* out,
                                  int64_t ldout, double* work, socks_stream_t stream) {
    SOCKS_CHECK_ARG(B != nullptr, 1);
    SOCKS_CHECK_ARG(n >= 1, 2);
    SOCKS_CHECK_ARG(m >= 1, 3);
    SOCKS_CHECK_ARG(ldb >= m, 4);
    SOCKS_CHECK_ARG(pts != nullptr, 5);
    SOCKS_CHECK_ARG(d >= 1 && d <= SOCKS_MAX_DIM, 6);
    SOCKS_CHECK_ARG(tubes != nullptr, 7);
    SOCKS_CHECK_ARG(problem == SOCKS_PROBLEM_THT || problem == SOCKS_PROBLEM_FHT, 8);
    SOCKS_CHECK_ARG(T >= 1, 9);
    SOCKS_CHECK_ARG(vf != nullptr && ldvf >= n, 10);
    SOCKS_CHECK_ARG(out != nullptr && ldout >= m, 12);
    SOCKS_CHECK_ARG(work != nullptr, 14);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ldp = colreduce_ldpart(m);
    double* den = work;
    double* part = work + ldp;
    const unsigned gb = (unsigned)ceil_div(m, 256);
    sr_init_kernel<<<gb, 256, 0, st>>>(pts, m, d, tubes, T, out, ldout);
    SOCKS_CHECK_LAUNCH();
    if (T == 1) return SOCKS_OK;
    int splits = colreduce_launch<false>(B, n, m, ldb, nullptr, 0, 1, part, 0, st);
    SOCKS_CHECK_LAUNCH();
    colreduce_finish_kernel<<<dim3(gb, 1), 256, 0, st>>>(part, ldp, splits, m, den, ldp);
    SOCKS_CHECK_LAUNCH();
    for (int t = T - 2; t >= 0; --t) {
        splits = colreduce_launch<false>(B, n, m, ldb, vf + (int64_t)(t + 1) * ldvf, 0, 1, part, 0, st);
        SOCKS_CHECK_LAUNCH();
        sr_finish_step_kernel<<<gb, 256, 0, st>>>(part, ldp, splits, den, pts, m, d, tubes + (int64_t)t * 4 * d,
                                                  problem, out + (int64_t)t * ldout);
        SOCKS_CHECK_LAUNCH();
    }
    return SOCKS_OK;
}

template <int D>
static void launch_predict(int variant, const double* X, const double* coef, int64_t n, int64_t n_pad,
                           const double* pts, int64_t m, double kmul, const double* tubes, int problem, int t0,
                           int R, double* out, int64_t ldout, double* part, int64_t ldp, int splits,
                           cudaStream_t st) {
    if (variant == 0) {
        // 8 warps x 32 points per CTA, three CTAs per SM (<= 85 registers): best of the CTA-shape sweep in
        // profiles/r1_predict_cfg_sweep.txt (all shapes within 4 %: the kernel is FP64-pipe bound)
        const ExpScaled ex = make_exp_scaled(kmul);
        sr_predict_dmma_kernel<D, SP_WARPS, 3>
            <<<dim3((unsigned)ceil_div(m, SP_PTS), (unsigned)splits), SP_WARPS * 32, 0, st>>>(X, coef, n, n_pad, pts, m,
                                                                                             ex, part, ldp);
        sr_predict_finish_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(part, ldp, splits, m, R, t0, pts, D,
                                                                             tubes, problem, out, ldout);
    } else {
        sr_predict_fma_kernel<D><<<(unsigned)ceil_div(m, 128), 128, 0, st>>>(X, coef, n, n_pad, pts, m, kmul, tubes,
                                                                            problem, t0, R, out, ldout);
    }
}

static inline int64_t predict_coef_doubles(int64_t n) { return round_up(n, SP_TS) * SR_CP + 64; }
static inline int predict_splits(int64_t n) { return (int)ceil_div(round_up(n, SP_TS), SP_SPLIT); }

extern "C" int64_t socks_sr_predict_work_bytes(int64_t n, int64_t m) {
    return (predict_coef_doubles(n) + (int64_t)predict_splits(n) * 16 * round_up(m, 2) + 2) * 8;
}

extern "C" int socks_sr_predict(const double* X, const double* w, const double* vf, int64_t ldvf, int64_t n, int d,
                                double scale, int divide, const double* pts, int64_t m, const double* tubes,
                                int problem, int T, double* out, int64_t ldout, double* work, int variant,
                                socks_stream_t stream) {
    SOCKS_CHECK_ARG(X != nullptr, 1);
    SOCKS_CHECK_ARG(w != nullptr, 2);
    SOCKS_CHECK_ARG(vf != nullptr && ldvf >= n, 3);
    SOCKS_CHECK_ARG(n >= 1, 5);
    SOCKS_CHECK_ARG(d >= 1 && d <= SOCKS_MAX_DIM, 6);
    SOCKS_CHECK_ARG(scale == scale && (!divide || scale != 0.0), 7);
    SOCKS_CHECK_ARG(m >= 0, 10);
    SOCKS_CHECK_ARG(tubes != nullptr, 11);
    SOCKS_CHECK_ARG(problem == SOCKS_PROBLEM_THT || problem == SOCKS_PROBLEM_FHT, 12);
    SOCKS_CHECK_ARG(T >= 1, 13);
    SOCKS_CHECK_ARG(ldout >= m, 15);
    SOCKS_CHECK_ARG(work != nullptr && (reinterpret_cast<uintptr_t>(work) & 15) == 0, 16);
    SOCKS_CHECK_ARG(variant == 0 || variant == 1, 17);
    if (m == 0) return SOCKS_OK;
    SOCKS_CHECK_ARG(pts != nullptr && out != nullptr, 9);
    // the fused path multiplies by the (rounded) reciprocal: 1 ulp of the argument, far below the 1e-8 bar
    const double kmul = divide ? 1.0 / scale : scale;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_pad = round_up(n, SP_TS);
    double* coef = work;
    double* part = work + predict_coef_doubles(n);
    const int64_t ldp = round_up(m, 2);
    const int splits = predict_splits(n);
    sr_init_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(pts, m, d, tubes, T, out, ldout);
    SOCKS_CHECK_LAUNCH();
    for (int t0 = 0; t0 < T - 1; t0 += SOCKS_SR_ROWS - 1) {
        const int steps = (T - 1 - t0 < SOCKS_SR_ROWS - 1) ? (T - 1 - t0) : (SOCKS_SR_ROWS - 1);
        const int R = steps + 1;
        sr_pack_coef_kernel<<<(unsigned)ceil_div(n_pad, 256), 256, 0, st>>>(w, vf, ldvf, n, n_pad, t0, R, coef);
        SOCKS_CHECK_LAUNCH();
#define SOCKS_PRED_CASE(DD)                                                                                      \
    case DD:                                                                                                     \
        launch_predict<DD>(variant, X, coef, n, n_pad, pts, m, kmul, tubes, problem, t0, R, out, ldout, part, ldp, \
                           splits, st);                                                                          \
        break;
        switch (d) {
            SOCKS_PRED_CASE(1) SOCKS_PRED_CASE(2) SOCKS_PRED_CASE(3) SOCKS_PRED_CASE(4)
            SOCKS_PRED_CASE(5) SOCKS_PRED_CASE(6) SOCKS_PRED_CASE(7) SOCKS_PRED_CASE(8)
            default:
                sr_predict_generic_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, st>>>(X, coef, n, d, pts, m, kmul, tubes,
                                                                                     problem, t0, R, out, ldout);
        }
#undef SOCKS_PRED_CASE
        SOCKS_CHECK_LAUNCH();
    }
    return SOCKS_OK;
}

extern "C" int socks_gemv_t(const double* Bmat, int64_t n, int64_t m, int64_t ldb, const double* V, int64_t ldv,
                            int R, double* y, int64_t ldy, double* work, socks_stream_t stream) {
    SOCKS_CHECK_ARG(Bmat != nullptr, 1);
    SOCKS_CHECK_ARG(n >= 1, 2);
    SOCKS_CHECK_ARG(m >= 1, 3);
    SOCKS_CHECK_ARG(ldb >= m, 4);
    SOCKS_CHECK_ARG(V == nullptr || ldv >= n || R == 1, 6);
    SOCKS_CHECK_ARG(R == 1 || R == 2, 7);
    SOCKS_CHECK_ARG(y != nullptr && (ldy >= m || R == 1), 8);
    SOCKS_CHECK_ARG(work != nullptr, 10);
    cudaStream_t st = (cudaStream_t)stream;
    const int splits = colreduce_launch<false>(Bmat, n, m, ldb, V, ldv, R, work, 0, st);
    SOCKS_CHECK_LAUNCH();
    colreduce_finish_kernel<<<dim3((unsigned)ceil_div(m, 256), (unsigned)R), 256, 0, st>>>(
        work, colreduce_ldpart(m), splits, m, y, ldy);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

extern "C" int socks_fwd_weights(const double* X, int64_t n, int d, double scale, int divide, const double* state,
                                 const double* z, int64_t ldz, int R, double* q, int64_t ldq, socks_stream_t stream) {
    SOCKS_CHECK_ARG(X != nullptr, 1);
    SOCKS_CHECK_ARG(n >= 1, 2);
    SOCKS_CHECK_ARG(d >= 1 && d <= SOCKS_MAX_DIM, 3);
    SOCKS_CHECK_ARG(scale == scale && (!divide || scale != 0.0), 4);
    SOCKS_CHECK_ARG(state != nullptr, 6);
    SOCKS_CHECK_ARG(z != nullptr, 7);
    SOCKS_CHECK_ARG(R >= 1 && R <= 2, 9);
    SOCKS_CHECK_ARG(q != nullptr, 10);
    fwd_weights_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
        X, n, d, divide ? 1.0 / scale : scale, state, z, ldz, R, q, ldq);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

extern "C" int socks_argmin_masked(const double* C, const double* D, int P, int* idx_dev, socks_stream_t stream) {
    SOCKS_CHECK_ARG(C != nullptr, 1);
    SOCKS_CHECK_ARG(P >= 1, 3);
    SOCKS_CHECK_ARG(idx_dev != nullptr, 4);
    argmin_masked_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(C, D, P, idx_dev);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}
