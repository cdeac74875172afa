// Stochastic-reachability backward recursion (KernelSR) and the forward-control score kernels.
//
// Fit (gym_socks/algorithms/reach/kernel_sr.py:297-309): beta lives in HBM as the un-normalised
// B = diag(w) K(X,Y) (n x n) plus its column sums; each of the T-1 time steps is one coalesced
// GEMV pass over B (reduce.cuh) followed by a fused normalise + FHT/THT step (HBM-bound).
//
// Predict lives in sr_predict.cu.
#include <climits>
#include <cmath>

#include "common.cuh"
#include "reduce.cuh"

namespace socks {

// out[T-1][j] = 1_target[T-1](pts_j)   (kernel_sr.py:57)
__global__ void sr_init_kernel(const double* __restrict__ pts, int64_t m, int d, const double* __restrict__ tubes,
                               int T, double* __restrict__ out, int64_t ldout) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double* tube = tubes + (int64_t)(T - 1) * 4 * d;
    out[(int64_t)(T - 1) * ldout + j] = in_box(pts + j * d, tube, 1, d) ? 1.0 : 0.0;
}

// out_t[j] = step(pts_j, (sum_s part[s][j]) / den[j])       (kernel_sr.py:62-63 after `normalize`).
// den_part != NULL: the pass also produced the unweighted column sums (partials at den_part): den[j] is formed here,
// stored for the following steps and used at once.
__global__ void sr_finish_step_kernel(const double* __restrict__ part, int64_t ldpart, int splits,
                                      double* __restrict__ den, const double* __restrict__ den_part,
                                      const double* __restrict__ pts, int64_t m, int d,
                                      const double* __restrict__ tube_t, int problem, double* __restrict__ out_t) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double a = 0.0;
    for (int s = 0; s < splits; ++s) a += part[(int64_t)s * ldpart + j];
    double dj;
    if (den_part) {
        dj = 0.0;
        for (int s = 0; s < splits; ++s) dj += den_part[(int64_t)s * ldpart + j];
        den[j] = dj;
    } else {
        dj = den[j];
    }
    out_t[j] = sr_step(pts + j * d, tube_t, d, problem, a / dj);
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
    // T-1 passes over B in total: the normaliser (column sums of B, `normalize` of kernel_sr.py:45) is accumulated by the
    // first GEMV pass from the same loads
    for (int t = T - 2; t >= 0; --t) {
        const bool first = (t == T - 2);
        const int splits = first ? colreduce_launch<false, true>(B, n, m, ldb, vf + (int64_t)(t + 1) * ldvf, 0, 1, part, 0, st)
                                 : colreduce_launch<false>(B, n, m, ldb, vf + (int64_t)(t + 1) * ldvf, 0, 1, part, 0, st);
        SOCKS_CHECK_LAUNCH();
        sr_finish_step_kernel<<<gb, 256, 0, st>>>(part, ldp, splits, den, first ? part + (int64_t)splits * ldp : nullptr,
                                                  pts, m, d, tubes + (int64_t)t * 4 * d, problem,
                                                  out + (int64_t)t * ldout);
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
