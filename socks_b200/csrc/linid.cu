// Kernel-based linear system identification (gym_socks/algorithms/identification/kernel_linear_id.py:107-185).
// With the linear kernel the regularised least-squares estimate collapses, by the matrix inversion lemma, to the
// primal normal equations on the concatenated regressor Z = [X | U | 1] (kernel_linear_id.py:133-149):
//     C_ZZ = Z^T Z + lambda N I   (z x z, z = d + m + 1),   C_YZ = Y^T Z,   H = C_YZ inv(C_ZZ).
// The two tall-skinny products are one HBM pass over Z and Y (socks_atb_small: bytes = 8 N (z + d)); the z x z solve
// reuses the padded Cholesky path (socks_pad_spd -> socks_potrf_inv -> socks_potrs).
#include "common.cuh"

namespace socks {

constexpr int AT_MAX = 64;      // largest p, q
constexpr int AT_ROWS = 32;     // rows per shared-memory tile
constexpr int AT_CHUNK = 2048;  // rows per CTA (fixed: the partial sums, and so every output bit, depend only on n)

// part[chunk][i][j] = sum_{r in chunk} A[r][i] B[r][j]
__global__ void __launch_bounds__(256)
    atb_small_kernel(const double* __restrict__ A, int64_t lda, int p, const double* __restrict__ B, int64_t ldb, int q,
                     int64_t n, double* __restrict__ part) {
    __shared__ double sa[AT_ROWS][AT_MAX + 1];
    __shared__ double sb[AT_ROWS][AT_MAX + 1];
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * AT_CHUNK, r1 = min(n, r0 + AT_CHUNK);
    const int npairs = p * q;
    double acc[(AT_MAX * AT_MAX + 255) / 256];
#pragma unroll
    for (int e = 0; e < (AT_MAX * AT_MAX + 255) / 256; ++e) acc[e] = 0.0;
    for (int64_t t0 = r0; t0 < r1; t0 += AT_ROWS) {
        const int rows = (int)min((int64_t)AT_ROWS, r1 - t0);
        __syncthreads();
        for (int e = tid; e < AT_ROWS * p; e += 256) {
            const int r = e / p, c = e % p;
            sa[r][c] = (r < rows) ? A[(t0 + r) * lda + c] : 0.0;
        }
        for (int e = tid; e < AT_ROWS * q; e += 256) {
            const int r = e / q, c = e % q;
            sb[r][c] = (r < rows) ? B[(t0 + r) * ldb + c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < (AT_MAX * AT_MAX + 255) / 256; ++e) {
            const int idx = tid + e * 256;
            if (idx < npairs) {
                const int i = idx / q, j = idx % q;
                double a = acc[e];
                for (int r = 0; r < AT_ROWS; ++r) a = fma(sa[r][i], sb[r][j], a);
                acc[e] = a;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < (AT_MAX * AT_MAX + 255) / 256; ++e) {
        const int idx = tid + e * 256;
        if (idx < npairs) part[(int64_t)blockIdx.x * npairs + idx] = acc[e];
    }
}

__global__ void atb_finish_kernel(const double* __restrict__ part, int chunks, int p, int q, double* __restrict__ C,
                                  int64_t ldc) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p * q) return;
    double a = 0.0;
    for (int c = 0; c < chunks; ++c) a += part[(int64_t)c * p * q + idx];
    C[(int64_t)(idx / q) * ldc + idx % q] = a;
}

// out[i][j] = sum_k A[i][k] T[j][k] + sum_k B[i][k] U[j][k]
__global__ void linear_map_kernel(const double* __restrict__ A, int lda, int dx, const double* __restrict__ B, int ldb,
                                  int du, int r, const double* __restrict__ T, const double* __restrict__ U, int64_t m,
                                  double* __restrict__ out, int64_t ldout) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    for (int i = 0; i < r; ++i) {
        double a = 0.0;
        for (int k = 0; k < dx; ++k) a = fma(__ldg(A + i * lda + k), T[j * dx + k], a);
        if (U) {
            double b = 0.0;
            for (int k = 0; k < du; ++k) b = fma(__ldg(B + i * ldb + k), U[j * du + k], b);
            a += b;  // result = A T^T; result += B U^T (kernel_linear_id.py:173-181)
        }
        out[(int64_t)i * ldout + j] = a;
    }
}

}  // namespace socks

using namespace socks;

extern "C" int64_t socks_atb_small_work_bytes(int64_t n, int p, int q) {
    return ceil_div(n > 0 ? n : 1, AT_CHUNK) * (int64_t)p * q * 8 + 16;
}

extern "C" int socks_atb_small(const double* A, int64_t lda, int p, const double* B, int64_t ldb, int q, int64_t n,
                               double* C, int64_t ldc, double* work, socks_stream_t stream) {
    SOCKS_CHECK_ARG(A != nullptr, 1);
    SOCKS_CHECK_ARG(p >= 1 && p <= AT_MAX && lda >= p, 3);
    SOCKS_CHECK_ARG(B != nullptr, 4);
    SOCKS_CHECK_ARG(q >= 1 && q <= AT_MAX && ldb >= q, 6);
    SOCKS_CHECK_ARG(n >= 1, 7);
    SOCKS_CHECK_ARG(C != nullptr && ldc >= q, 8);
    SOCKS_CHECK_ARG(work != nullptr, 10);
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = (int)ceil_div(n, AT_CHUNK);
    atb_small_kernel<<<chunks, 256, 0, st>>>(A, lda, p, B, ldb, q, n, work);
    SOCKS_CHECK_LAUNCH();
    atb_finish_kernel<<<(p * q + 255) / 256, 256, 0, st>>>(work, chunks, p, q, C, ldc);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

extern "C" int socks_linear_map(const double* A, int lda, int dx, const double* B, int ldb, int du, int r,
                                const double* T, const double* U, int64_t m, double* out, int64_t ldout,
                                socks_stream_t stream) {
    SOCKS_CHECK_ARG(A != nullptr && dx >= 1 && lda >= dx, 1);
    SOCKS_CHECK_ARG(U == nullptr || (B != nullptr && du >= 1 && ldb >= du), 4);
    SOCKS_CHECK_ARG(r >= 1, 7);
    SOCKS_CHECK_ARG(m >= 0, 10);
    SOCKS_CHECK_ARG(ldout >= m, 12);
    if (m == 0) return SOCKS_OK;
    SOCKS_CHECK_ARG(T != nullptr && out != nullptr, 8);
    linear_map_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, (cudaStream_t)stream>>>(A, lda, dx, B, ldb, du, r, T, U, m, out,
                                                                                  ldout);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}
