// Maximal stochastic reachability (gym_socks/algorithms/reach/kernel_sr_max.py:30-79).
//
// The reference materialises beta[i,j,a] = W_ii CXY_ij CUA_ia / sum_i(...) (kernel_sr_max.py:45-47), an
// N x M x P tensor, and contracts it with vf[t+1] each step (:65-70).  The same numbers are two dense
// contractions sharing the left operand,
//     den[j][a]   = sum_i K(x_i,p_j) (w_i)           CUA[i][a]
//     num_t[j][a] = sum_i K(x_i,p_j) (vf[t+1][i] w_i) CUA[i][a],      wX = num_t / den, VX = max_a wX
// i.e. one GEMM  Cm (m x R*P) = Kmat^T (m x N) * Bm (N x R*P)  on the fp64 tensor cores with the R
// coefficient rows stacked along the columns, followed by a fused divide + row-max + FHT/THT step.
#include <cmath>

#include "gemm_f64.cuh"

namespace socks {

__global__ void srmax_pack_kernel(const double* __restrict__ CUA, int64_t n, int P, const double* __restrict__ w,
                                  const double* __restrict__ vf, int64_t ldvf, int t0, int first_row, int R,
                                  double* __restrict__ Bm, int64_t np, int64_t ldbm) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (c >= ldbm) return;
    double v = 0.0;
    if (i < n && c < (int64_t)R * P) {
        const int r = first_row + (int)(c / P), a = (int)(c % P);
        const double coef = (r == 0) ? w[i] : vf[(int64_t)(t0 + r) * ldvf + i] * w[i];
        v = coef * CUA[i * P + a];
    }
    Bm[i * ldbm + c] = v;
}

// one warp per point: VX_r = max_a num_r[a] / den[a] (np.max semantics: NaN propagates), then the step.
__global__ void srmax_step_kernel(const double* __restrict__ Cm, int64_t ldc, const double* __restrict__ Den,
                                  int64_t ldden, int64_t m, int P, int R, int t0, const double* __restrict__ pts, int d,
                                  const double* __restrict__ tubes, int problem, int T, double* __restrict__ out,
                                  int64_t ldout, int write_last) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= m) return;
    // Den == NULL: Cm holds [den | num_1 .. num_{R-1}];  else Cm holds [num_1 .. num_{R-1}] and Den the normaliser
    const double* den = Den ? Den + j * ldden : Cm + j * ldc;
    const double* row = Den ? Cm + j * ldc - P : Cm + j * ldc;
    for (int r = 1; r < R; ++r) {
        double best = -INFINITY;
        bool nan = false;
        for (int a = lane; a < P; a += 32) {
            const double v = row[(int64_t)r * P + a] / den[a];
            nan = nan || (v != v);
            best = fmax(best, v);
        }
        for (int off = 16; off > 0; off >>= 1) {
            best = fmax(best, __shfl_xor_sync(0xffffffffu, best, off));
            nan = nan || __shfl_xor_sync(0xffffffffu, (int)nan, off);
        }
        if (lane == 0) {
            const int t = t0 + r - 1;
            const double V = nan ? NAN : best;
            out[(int64_t)t * ldout + j] = sr_step(pts + j * d, tubes + (int64_t)t * 4 * d, d, problem, V);
        }
    }
    if (write_last && lane == 0)
        out[(int64_t)(T - 1) * ldout + j] = in_box(pts + j * d, tubes + (int64_t)(T - 1) * 4 * d, 1, d) ? 1.0 : 0.0;
}

// Bm[i][r*P + a] = coefs[r][i] * CUA[i][a]: the right operand of C = K(X,Y)^T diag(z) CUA for arbitrary row
// vectors z (kernel_control_bwd.py:90-97: Z = vf[t+1] @ W, R = constraint_t @ W).
__global__ void scale_pack_kernel(const double* __restrict__ CUA, int64_t n, int P, const double* __restrict__ coefs,
                                  int64_t ldcoef, int R, double* __restrict__ Bm, int64_t ldbm) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (c >= ldbm) return;
    double v = 0.0;
    if (i < n && c < (int64_t)R * P) {
        const int r = (int)(c / P), a = (int)(c % P);
        v = coefs[(int64_t)r * ldcoef + i] * CUA[i * P + a];
    }
    Bm[i * ldbm + c] = v;
}

// out[j] = add[j] + min_a { C[j][a] : D[j][a] <= 0 }  (all a if unconstrained or no a is feasible);
// C = Cm[:, 0:P], D = Cm[:, P:2P].  np.min semantics (NaN among the candidates propagates).
// kernel_control_bwd.py:99-115.  One warp per row.
__global__ void rowmin_masked_kernel(const double* __restrict__ Cm, int64_t ldc, int64_t m, int P, int constrained,
                                     const double* __restrict__ add, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= m) return;
    const double* row = Cm + j * ldc;
    double best_all = INFINITY, best_ok = INFINITY;
    int nan_all = 0, nan_ok = 0, any_ok = 0;
    for (int a = lane; a < P; a += 32) {
        const double c = row[a];
        const bool isn = (c != c);
        nan_all |= isn;
        best_all = fmin(best_all, c);
        if (constrained && row[P + a] <= 0.0) {
            any_ok = 1;
            nan_ok |= isn;
            best_ok = fmin(best_ok, c);
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        best_all = fmin(best_all, __shfl_xor_sync(0xffffffffu, best_all, off));
        best_ok = fmin(best_ok, __shfl_xor_sync(0xffffffffu, best_ok, off));
        nan_all |= __shfl_xor_sync(0xffffffffu, nan_all, off);
        nan_ok |= __shfl_xor_sync(0xffffffffu, nan_ok, off);
        any_ok |= __shfl_xor_sync(0xffffffffu, any_ok, off);
    }
    if (lane == 0) {
        double v = (constrained && any_ok) ? (nan_ok ? NAN : best_ok) : (nan_all ? NAN : best_all);
        out[j] = add[j] + v;
    }
}

}  // namespace socks

using namespace socks;

extern "C" int socks_scale_pack(const double* CUA, int64_t n, int P, const double* coefs, int64_t ldcoef, int R,
                                double* Bm, int64_t np, int64_t ldbm, socks_stream_t stream) {
    SOCKS_CHECK_ARG(CUA != nullptr, 1);
    SOCKS_CHECK_ARG(n >= 1, 2);
    SOCKS_CHECK_ARG(P >= 1, 3);
    SOCKS_CHECK_ARG(coefs != nullptr && ldcoef >= n, 4);
    SOCKS_CHECK_ARG(R >= 1, 6);
    SOCKS_CHECK_ARG(Bm != nullptr, 7);
    SOCKS_CHECK_ARG(np >= n && np <= 65535, 8);
    SOCKS_CHECK_ARG(ldbm >= (int64_t)R * P && ldbm % 2 == 0, 9);
    dim3 grid((unsigned)ceil_div(ldbm, 256), (unsigned)np);
    scale_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(CUA, n, P, coefs, ldcoef, R, Bm, ldbm);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

extern "C" int socks_rowmin_masked(const double* Cm, int64_t ldc, int64_t m, int P, int constrained,
                                   const double* add, double* out, socks_stream_t stream) {
    SOCKS_CHECK_ARG(Cm != nullptr, 1);
    SOCKS_CHECK_ARG(m >= 1, 3);
    SOCKS_CHECK_ARG(P >= 1 && ldc >= (constrained ? 2 : 1) * (int64_t)P, 4);
    SOCKS_CHECK_ARG(add != nullptr, 6);
    SOCKS_CHECK_ARG(out != nullptr, 7);
    rowmin_masked_kernel<<<(unsigned)ceil_div(m, 8), 256, 0, (cudaStream_t)stream>>>(Cm, ldc, m, P, constrained, add,
                                                                                     out);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

extern "C" int socks_srmax_pack(const double* CUA, int64_t n, int P, const double* w, const double* vf, int64_t ldvf,
                                int t0, int first_row, int R, double* Bm, int64_t np, int64_t ldbm,
                                socks_stream_t stream) {
    SOCKS_CHECK_ARG(CUA != nullptr, 1);
    SOCKS_CHECK_ARG(n >= 1, 2);
    SOCKS_CHECK_ARG(P >= 1, 3);
    SOCKS_CHECK_ARG(w != nullptr, 4);
    SOCKS_CHECK_ARG(first_row >= 0, 8);
    SOCKS_CHECK_ARG(R >= 1 && (first_row + R == 1 || vf != nullptr), 9);
    SOCKS_CHECK_ARG(Bm != nullptr, 10);
    SOCKS_CHECK_ARG(np >= n && np <= 65535 * 16, 11);
    SOCKS_CHECK_ARG(ldbm >= (int64_t)R * P && ldbm % 2 == 0, 12);
    cudaStream_t st = (cudaStream_t)stream;
    for (int64_t r0 = 0; r0 < np; r0 += 65535) {
        const int64_t rows = (np - r0 < 65535) ? (np - r0) : 65535;
        dim3 grid((unsigned)ceil_div(ldbm, 256), (unsigned)rows);
        // shift the row window: kernel indexes rows by blockIdx.y, so offset the row-indexed pointers
        srmax_pack_kernel<<<grid, 256, 0, st>>>(CUA + r0 * P, n - r0 > 0 ? n - r0 : 0, P, w + r0, vf ? vf + r0 : vf,
                                                ldvf, t0, first_row, R, Bm + r0 * ldbm, np - r0, ldbm);
    }
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

extern "C" int socks_gemm_tn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                             int64_t m, int64_t nn, int64_t K, socks_stream_t stream) {
    SOCKS_CHECK_ARG(A != nullptr && (reinterpret_cast<uintptr_t>(A) & 15) == 0, 1);
    SOCKS_CHECK_ARG(lda >= m && lda % 2 == 0, 2);
    SOCKS_CHECK_ARG(B != nullptr && (reinterpret_cast<uintptr_t>(B) & 15) == 0, 3);
    SOCKS_CHECK_ARG(ldb >= nn && ldb % 2 == 0, 4);
    SOCKS_CHECK_ARG(C != nullptr && (reinterpret_cast<uintptr_t>(C) & 15) == 0, 5);
    SOCKS_CHECK_ARG(ldc >= nn && ldc % 2 == 0, 6);
    SOCKS_CHECK_ARG(m >= 0 && m % 128 == 0, 7);
    SOCKS_CHECK_ARG(nn >= 0 && nn % 128 == 0, 8);
    SOCKS_CHECK_ARG(K >= 0 && K % 16 == 0, 9);
    cudaError_t e = gemm_f64<true, false>(A, lda, B, ldb, C, ldc, m, nn, K, 1.0, 0.0, 0, 0, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        set_error("socks_gemm_tn: CUDA launch failed: %s", cudaGetErrorString(e));
        return SOCKS_ECUDA;
    }
    return SOCKS_OK;
}

// General entry (tests, profiling, and the building block a caller can use directly).
extern "C" int socks_gemm_f64(int transa, int transb, const double* A, int64_t lda, const double* B, int64_t ldb,
                              double* C, int64_t ldc, int64_t m, int64_t nn, int64_t K, double alpha, double beta,
                              int flags, int lower_only, int tile_cfg, socks_stream_t stream) {
    SOCKS_CHECK_ARG(!(transa && transb), 1);
    SOCKS_CHECK_ARG(A != nullptr && (reinterpret_cast<uintptr_t>(A) & 15) == 0, 3);
    SOCKS_CHECK_ARG(lda % 2 == 0 && lda >= (transa ? m : K), 4);
    SOCKS_CHECK_ARG(B != nullptr && (reinterpret_cast<uintptr_t>(B) & 15) == 0, 5);
    SOCKS_CHECK_ARG(ldb % 2 == 0 && ldb >= (transb ? K : nn), 6);
    SOCKS_CHECK_ARG(C != nullptr && (reinterpret_cast<uintptr_t>(C) & 15) == 0, 7);
    SOCKS_CHECK_ARG(ldc >= nn && ldc % 2 == 0, 8);
    SOCKS_CHECK_ARG(m >= 0 && m % 128 == 0, 9);
    SOCKS_CHECK_ARG(nn >= 0 && nn % 128 == 0, 10);
    SOCKS_CHECK_ARG(K >= 0 && K % 16 == 0, 11);
    SOCKS_CHECK_ARG(alpha != 0.0, 12);
    SOCKS_CHECK_ARG(tile_cfg == 0 || tile_cfg == 1, 16);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (transa)
        e = gemm_f64<true, false>(A, lda, B, ldb, C, ldc, m, nn, K, alpha, beta, flags, lower_only, st, tile_cfg);
    else if (transb)
        e = gemm_f64<false, true>(A, lda, B, ldb, C, ldc, m, nn, K, alpha, beta, flags, lower_only, st, tile_cfg);
    else
        e = gemm_f64<false, false>(A, lda, B, ldb, C, ldc, m, nn, K, alpha, beta, flags, lower_only, st, tile_cfg);
    if (e != cudaSuccess) {
        set_error("socks_gemm_f64: CUDA launch failed: %s", cudaGetErrorString(e));
        return SOCKS_ECUDA;
    }
    return SOCKS_OK;
}

extern "C" int socks_srmax_step(const double* Cm, int64_t ldc, const double* Den, int64_t ldden, int64_t m, int P,
                                int R, int t0, const double* pts, int d, const double* tubes, int problem, int T,
                                double* out, int64_t ldout, int write_last, socks_stream_t stream) {
    SOCKS_CHECK_ARG(Cm != nullptr, 1);
    SOCKS_CHECK_ARG(Den == nullptr || ldden >= P, 4);
    SOCKS_CHECK_ARG(m >= 1, 5);
    SOCKS_CHECK_ARG(P >= 1, 6);
    SOCKS_CHECK_ARG(R >= 1 && ldc >= (int64_t)(Den ? R - 1 : R) * P, 7);
    SOCKS_CHECK_ARG(pts != nullptr, 9);
    SOCKS_CHECK_ARG(d >= 1 && d <= SOCKS_MAX_DIM, 10);
    SOCKS_CHECK_ARG(tubes != nullptr, 11);
    SOCKS_CHECK_ARG(problem == SOCKS_PROBLEM_THT || problem == SOCKS_PROBLEM_FHT, 12);
    SOCKS_CHECK_ARG(T >= 1 && t0 >= 0 && t0 + R - 1 <= T - 1, 13);
    SOCKS_CHECK_ARG(out != nullptr && ldout >= m, 14);
    srmax_step_kernel<<<(unsigned)ceil_div(m, 8), 256, 0, (cudaStream_t)stream>>>(
        Cm, ldc, Den, ldden, m, P, R, t0, pts, d, tubes, problem, T, out, ldout, write_last);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}
