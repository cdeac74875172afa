// Gram builders: K(X,Y), squared distances, and the padded symmetric G = K(Z,Z) + c I.
// Replaces gym_socks/kernel/metrics.py:119-138 (per-dimension np.tile temporaries + exp).
//
// Roofline: per entry 3d+2 flops for the distance + the fp64 exp (14 DFMA +
// ~4 other FP64-pipe ops in CUDA's exp) against 8 bytes written: at d=2 the FP64 pipe and the HBM
// write stream are within ~10% of each other on B200, so tiles are sized for fully coalesced
// 16-byte stores (each warp writes 512 contiguous bytes per row) and the inputs stay in registers.
#include "common.cuh"

namespace socks {

constexpr int GR_TX = 32, GR_TY = 8, GR_RI = 4;  // thread block 32x8, 4 rows x 2 cols per thread
constexpr int GR_ROWS = GR_TY * GR_RI;           // 32 rows per CTA
constexpr int GR_COLS = GR_TX * 2;               // 64 cols per CTA

// Squared distance with numpy's roundings: K += (Y_k - X_k)^2 dimension by dimension, each square and
// each sum rounded separately (metrics.py:127-128) -- no FMA contraction, so the Gram matrix differs
// from the reference only through the last-bit behaviour of exp.
template <int D>
__device__ __forceinline__ double sqdist_reg(const double (&x)[D], const double (&y)[D]) {
    double t = __dsub_rn(y[0], x[0]);
    double acc = __dmul_rn(t, t);
#pragma unroll
    for (int k = 1; k < D; ++k) {
        t = __dsub_rn(y[k], x[k]);
        acc = __dadd_rn(acc, __dmul_rn(t, t));
    }
    return acc;
}

// Kernel argument d2 * scale, or d2 / scale as a correctly rounded quotient (metrics.py:135 divides by
// -2 sigma^2): q = d2 * rcp, one Newton step on the exact residual (rcp = RN(1/scale) from the host).
struct KScale {
    double scale, rcp;
    int divide;
};
__device__ __forceinline__ double rbf_arg(double d2, const KScale& ks) {
    if (!ks.divide) return d2 * ks.scale;
    const double q = d2 * ks.rcp;
    const double r = fma(-q, ks.scale, d2);
    return fma(r, ks.rcp, q);
}
inline KScale make_kscale(double scale, int divide) { return KScale{scale, divide ? 1.0 / scale : 0.0, divide}; }

// MODE 0: rowscale * exp(arg(d2)) (RBF); 1: d2; 2: rowscale * exp(arg(sqrt(d2))) (Abel, metrics.py:194-204); 3: sqrt(d2)
template <int D, int MODE>
__global__ void __launch_bounds__(GR_TX* GR_TY)
    gram_cross_kernel(const double* __restrict__ X, int64_t n, const double* __restrict__ Y, int64_t m,
                      KScale ks, const double* __restrict__ rowscale, double* __restrict__ K, int64_t ldk,
                      int vec_ok) {
    const int64_t j0 = (int64_t)blockIdx.x * GR_COLS + threadIdx.x * 2;
    const int64_t i0 = (int64_t)blockIdx.y * GR_ROWS + threadIdx.y * GR_RI;
    double y[2][D], x[GR_RI][D];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        int64_t j = min(j0 + c, m - 1);
#pragma unroll
        for (int k = 0; k < D; ++k) y[c][k] = __ldg(Y + j * D + k);
    }
#pragma unroll
    for (int r = 0; r < GR_RI; ++r) {
        int64_t i = min(i0 + r, n - 1);
#pragma unroll
        for (int k = 0; k < D; ++k) x[r][k] = __ldg(X + i * D + k);
    }
#pragma unroll
    for (int r = 0; r < GR_RI; ++r) {
        int64_t i = i0 + r;
        if (i >= n) break;
        double v0 = sqdist_reg<D>(x[r], y[0]);
        double v1 = sqdist_reg<D>(x[r], y[1]);
        if (MODE == 2 || MODE == 3) {
            v0 = sqrt(v0);
            v1 = sqrt(v1);
        }
        if (MODE == 0 || MODE == 2) {
            double s = rowscale ? __ldg(rowscale + i) : 1.0;
            v0 = exp(rbf_arg(v0, ks));
            v1 = exp(rbf_arg(v1, ks));
            if (rowscale) { v0 *= s; v1 *= s; }
        }
        double* dst = K + i * ldk + j0;
        if (vec_ok && j0 + 1 < m) {
            *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
        } else {
            if (j0 < m) dst[0] = v0;
            if (j0 + 1 < m) dst[1] = v1;
        }
    }
}

// Runtime-d fallback (d > 8): one entry per thread, inputs through L1.
template <int MODE>
__global__ void gram_cross_generic_kernel(const double* __restrict__ X, int64_t n, const double* __restrict__ Y,
                                          int64_t m, int d, KScale ks, const double* __restrict__ rowscale,
                                          double* __restrict__ K, int64_t ldk) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t i = blockIdx.y;
    if (j >= m || i >= n) return;
    double acc = 0.0;
    for (int k = 0; k < d; ++k) {
        double t = __dsub_rn(__ldg(Y + j * d + k), __ldg(X + i * d + k));
        acc = __dadd_rn(acc, __dmul_rn(t, t));
    }
    if (MODE == 2 || MODE == 3) acc = sqrt(acc);
    if (MODE == 0 || MODE == 2) {
        acc = exp(rbf_arg(acc, ks));
        if (rowscale) acc *= __ldg(rowscale + i);
    }
    K[i * ldk + j] = acc;
}

// Lower 128-tiles of the padded G = K(Z,Z) + diag_add I; identity in the padding.
template <int D, bool ABEL>
__global__ void __launch_bounds__(GR_TX* GR_TY)
    gram_sym_kernel(const double* __restrict__ Z, int64_t n, KScale ks, double diag_add,
                    double* __restrict__ G, int64_t ldg) {
    const int64_t jt = (int64_t)blockIdx.x * GR_COLS;
    const int64_t it = (int64_t)blockIdx.y * GR_ROWS;
    if (jt / TILE > it / TILE) return;  // strictly-upper 128-tile: never read by the factorisation
    const int64_t j0 = jt + threadIdx.x * 2;
    const int64_t i0 = it + threadIdx.y * GR_RI;
    double y[2][D], x[GR_RI][D];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        int64_t j = min(j0 + c, n - 1);
#pragma unroll
        for (int k = 0; k < D; ++k) y[c][k] = __ldg(Z + j * D + k);
    }
#pragma unroll
    for (int r = 0; r < GR_RI; ++r) {
        int64_t i = min(i0 + r, n - 1);
#pragma unroll
        for (int k = 0; k < D; ++k) x[r][k] = __ldg(Z + i * D + k);
    }
#pragma unroll
    for (int r = 0; r < GR_RI; ++r) {
        int64_t i = i0 + r;
        double v[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            int64_t j = j0 + c;
            if (i < n && j < n) {
                const double d2 = sqdist_reg<D>(x[r], y[c]);
                v[c] = exp(rbf_arg(ABEL ? sqrt(d2) : d2, ks));
                if (i == j) v[c] += diag_add;
            } else {
                v[c] = (i == j) ? 1.0 : 0.0;
            }
        }
        *reinterpret_cast<double2*>(G + i * ldg + j0) = make_double2(v[0], v[1]);
    }
}

__global__ void gram_sym_generic_kernel(const double* __restrict__ Z, int64_t n, int d, KScale ks, int abel,
                                        double diag_add, double* __restrict__ G, int64_t ldg, int64_t np) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t i = blockIdx.y;
    if (j >= np || i >= np || j / TILE > i / TILE) return;
    double v;
    if (i < n && j < n) {
        double acc = 0.0;
        for (int k = 0; k < d; ++k) {
            double t = __dsub_rn(__ldg(Z + j * d + k), __ldg(Z + i * d + k));
            acc = __dadd_rn(acc, __dmul_rn(t, t));
        }
        v = exp(rbf_arg(abel ? sqrt(acc) : acc, ks)) + (i == j ? diag_add : 0.0);
    } else {
        v = (i == j) ? 1.0 : 0.0;
    }
    G[i * ldg + j] = v;
}

// ---- kernel families beyond RBF/Abel (sklearn.metrics.pairwise kernels the reference accepts as `kernel_fn`,
// gym_socks/kernel/metrics.py:239-241, pinned by gym_socks/kernel/tests/test_kernel.py:18,132-147) ------------
//   LINEAR     x.y                              sklearn linear_kernel
//   POLY       (scale x.y + coef0)^degree       sklearn polynomial_kernel (scale = gamma)
//   LAPLACIAN  exp(scale |x - y|_1)             sklearn laplacian_kernel (scale = -gamma)
// plus RBF and Abel through the same code, so that the product Gram K(X,X) * K(U,U) (metrics.py:268-276) exists
// for every family.  One entry per thread, runtime d: these families are off the benchmarked path.
struct KFamily {
    int kind;
    KScale ks;
    double coef0, degree;
};

__device__ __forceinline__ double family_value(const KFamily& f, const double* __restrict__ x,
                                               const double* __restrict__ y, int d) {
    if (f.kind == SOCKS_KERNEL_RBF || f.kind == SOCKS_KERNEL_ABEL) {
        double acc = 0.0;
        for (int k = 0; k < d; ++k) {
            const double t = __dsub_rn(__ldg(y + k), __ldg(x + k));
            acc = __dadd_rn(acc, __dmul_rn(t, t));
        }
        return exp(rbf_arg(f.kind == SOCKS_KERNEL_ABEL ? sqrt(acc) : acc, f.ks));
    }
    if (f.kind == SOCKS_KERNEL_LAPLACIAN) {
        double acc = 0.0;
        for (int k = 0; k < d; ++k) acc = __dadd_rn(acc, fabs(__dsub_rn(__ldg(x + k), __ldg(y + k))));
        return exp(rbf_arg(acc, f.ks));
    }
    double dot = 0.0;
    for (int k = 0; k < d; ++k) dot = fma(__ldg(x + k), __ldg(y + k), dot);
    if (f.kind == SOCKS_KERNEL_LINEAR) return dot;
    return pow(__dadd_rn(__dmul_rn(dot, f.ks.scale), f.coef0), f.degree);  // K *= gamma; K += coef0; K **= degree
}

__global__ void gram_family_kernel(KFamily f, const double* __restrict__ X, int64_t n, const double* __restrict__ Y,
                                   int64_t m, int d, const double* __restrict__ rowscale, double* __restrict__ K,
                                   int64_t ldk) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j >= m || i >= n) return;
    double v = family_value(f, X + i * d, Y + j * d, d);
    if (rowscale) v *= __ldg(rowscale + i);
    K[i * ldk + j] = v;
}

// Lower 128-tiles of the padded G = k(X,X) [* k(U,U)] + diag_add I; identity in the padding.
__global__ void gram_sym_family_kernel(KFamily f, const double* __restrict__ X, int64_t n, int dx,
                                       const double* __restrict__ U, int du, double diag_add, double* __restrict__ G,
                                       int64_t ldg, int64_t np) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j >= np || i >= np || j / TILE > i / TILE) return;
    double v;
    if (i < n && j < n) {
        v = family_value(f, X + i * dx, X + j * dx, dx);
        if (U) v = __dmul_rn(family_value(f, U + i * du, U + j * du, du), v);  // np.multiply(kernel_fn(U, V), K)
        if (i == j) v += diag_add;
    } else {
        v = (i == j) ? 1.0 : 0.0;
    }
    G[i * ldg + j] = v;
}

template <int MODE>
static int launch_cross(const double* X, int64_t n, const double* Y, int64_t m, int d, KScale ks,
                        const double* rowscale, double* K, int64_t ldk, cudaStream_t st) {
    if (n == 0 || m == 0) return SOCKS_OK;
    if (d <= 8 && ceil_div(n, GR_ROWS) <= 65535) {
        dim3 block(GR_TX, GR_TY), grid((unsigned)ceil_div(m, GR_COLS), (unsigned)ceil_div(n, GR_ROWS));
        int vec_ok = (ldk % 2 == 0) && ((reinterpret_cast<uintptr_t>(K) & 15) == 0);
#define SOCKS_GRAM_CASE(DD)                                                                              \
    case DD:                                                                                             \
        gram_cross_kernel<DD, MODE><<<grid, block, 0, st>>>(X, n, Y, m, ks, rowscale, K, ldk, vec_ok); \
        break;
        switch (d) {
            SOCKS_GRAM_CASE(1) SOCKS_GRAM_CASE(2) SOCKS_GRAM_CASE(3) SOCKS_GRAM_CASE(4)
            SOCKS_GRAM_CASE(5) SOCKS_GRAM_CASE(6) SOCKS_GRAM_CASE(7) SOCKS_GRAM_CASE(8)
        }
#undef SOCKS_GRAM_CASE
    } else {
        for (int64_t r0 = 0; r0 < n; r0 += 65535) {  // grid.y limit
            int64_t rows = (n - r0 < 65535) ? (n - r0) : 65535;
            dim3 block(256), grid((unsigned)ceil_div(m, 256), (unsigned)rows);
            gram_cross_generic_kernel<MODE><<<grid, block, 0, st>>>(X + r0 * d, rows, Y, m, d, ks,
                                                                   rowscale ? rowscale + r0 : nullptr,
                                                                   K + r0 * ldk, ldk);
        }
    }
    return SOCKS_OK;
}

}  // namespace socks

using namespace socks;

static int gram_kernel_impl(int kind, const double* X, int64_t n, const double* Y, int64_t m, int d, double scale,
                            int divide, const double* rowscale, double* K, int64_t ldk, socks_stream_t stream) {
    SOCKS_CHECK_ARG(kind == SOCKS_KERNEL_RBF || kind == SOCKS_KERNEL_ABEL, 1);
    SOCKS_CHECK_ARG(n >= 0, 3);
    SOCKS_CHECK_ARG(m >= 0, 5);
    SOCKS_CHECK_ARG(d >= 1 && d <= SOCKS_MAX_DIM, 6);
    SOCKS_CHECK_ARG(scale == scale && (!divide || scale != 0.0), 7);
    SOCKS_CHECK_ARG(ldk >= m, 11);
    SOCKS_CHECK_ARG((n == 0 || m == 0) || (X && Y && K), 2);
    const KScale ks = make_kscale(scale, divide);
    if (kind == SOCKS_KERNEL_RBF) launch_cross<0>(X, n, Y, m, d, ks, rowscale, K, ldk, (cudaStream_t)stream);
    else launch_cross<2>(X, n, Y, m, d, ks, rowscale, K, ldk, (cudaStream_t)stream);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

extern "C" int socks_gram_kernel(int kind, const double* X, int64_t n, const double* Y, int64_t m, int d, double scale,
                                 int divide, const double* rowscale, double* K, int64_t ldk, socks_stream_t stream) {
    return gram_kernel_impl(kind, X, n, Y, m, d, scale, divide, rowscale, K, ldk, stream);
}

extern "C" int socks_gram_rbf(const double* X, int64_t n, const double* Y, int64_t m, int d, double scale, int divide,
                              const double* rowscale, double* K, int64_t ldk, socks_stream_t stream) {
    int rc = gram_kernel_impl(SOCKS_KERNEL_RBF, X, n, Y, m, d, scale, divide, rowscale, K, ldk, stream);
    return rc < -1 ? rc + 1 : rc;  // argument positions of this entry point are shifted by one
}

extern "C" int socks_sqdist(const double* X, int64_t n, const double* Y, int64_t m, int d, double* D, int64_t ldd,
                            socks_stream_t stream) {
    SOCKS_CHECK_ARG(n >= 0, 2);
    SOCKS_CHECK_ARG(m >= 0, 4);
    SOCKS_CHECK_ARG(d >= 1 && d <= SOCKS_MAX_DIM, 5);
    SOCKS_CHECK_ARG(ldd >= m, 7);
    SOCKS_CHECK_ARG((n == 0 || m == 0) || (X && Y && D), 1);
    launch_cross<1>(X, n, Y, m, d, make_kscale(0.0, 0), nullptr, D, ldd, (cudaStream_t)stream);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

extern "C" int socks_dist(const double* X, int64_t n, const double* Y, int64_t m, int d, double* D, int64_t ldd,
                          socks_stream_t stream) {
    SOCKS_CHECK_ARG(n >= 0, 2);
    SOCKS_CHECK_ARG(m >= 0, 4);
    SOCKS_CHECK_ARG(d >= 1 && d <= SOCKS_MAX_DIM, 5);
    SOCKS_CHECK_ARG(ldd >= m, 7);
    SOCKS_CHECK_ARG((n == 0 || m == 0) || (X && Y && D), 1);
    launch_cross<3>(X, n, Y, m, d, make_kscale(0.0, 0), nullptr, D, ldd, (cudaStream_t)stream);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

static int gram_sym_impl(int kind, const double* Z, int64_t n, int d, double scale, int divide, double diag_add,
                         double* G, int64_t ldg, socks_stream_t stream) {
    SOCKS_CHECK_ARG(kind == SOCKS_KERNEL_RBF || kind == SOCKS_KERNEL_ABEL, 1);
    SOCKS_CHECK_ARG(Z != nullptr, 2);
    SOCKS_CHECK_ARG(n >= 1, 3);
    SOCKS_CHECK_ARG(d >= 1 && d <= SOCKS_MAX_DIM, 4);
    const int64_t np = socks_padded_dim(n);
    SOCKS_CHECK_ARG(scale == scale && (!divide || scale != 0.0), 5);
    SOCKS_CHECK_ARG(G != nullptr && (reinterpret_cast<uintptr_t>(G) & 15) == 0, 8);
    SOCKS_CHECK_ARG(ldg >= np && ldg % 2 == 0, 9);
    const KScale ks = make_kscale(scale, divide);
    const bool abel = (kind == SOCKS_KERNEL_ABEL);
    cudaStream_t st = (cudaStream_t)stream;
    if (d <= 8 && np / GR_ROWS <= 65535) {
        dim3 block(GR_TX, GR_TY), grid((unsigned)(np / GR_COLS), (unsigned)(np / GR_ROWS));
#define SOCKS_SYM_CASE(DD)                                                                       \
    case DD:                                                                                     \
        if (abel) gram_sym_kernel<DD, true><<<grid, block, 0, st>>>(Z, n, ks, diag_add, G, ldg); \
        else gram_sym_kernel<DD, false><<<grid, block, 0, st>>>(Z, n, ks, diag_add, G, ldg);     \
        break;
        switch (d) {
            SOCKS_SYM_CASE(1) SOCKS_SYM_CASE(2) SOCKS_SYM_CASE(3) SOCKS_SYM_CASE(4)
            SOCKS_SYM_CASE(5) SOCKS_SYM_CASE(6) SOCKS_SYM_CASE(7) SOCKS_SYM_CASE(8)
        }
#undef SOCKS_SYM_CASE
    } else {
        SOCKS_CHECK_ARG(np <= 65535, 3);
        dim3 block(256), grid((unsigned)ceil_div(np, 256), (unsigned)np);
        gram_sym_generic_kernel<<<grid, block, 0, st>>>(Z, n, d, ks, abel ? 1 : 0, diag_add, G, ldg, np);
    }
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

extern "C" int socks_gram_sym_kernel(int kind, const double* Z, int64_t n, int d, double scale, int divide,
                                     double diag_add, double* G, int64_t ldg, socks_stream_t stream) {
    return gram_sym_impl(kind, Z, n, d, scale, divide, diag_add, G, ldg, stream);
}

extern "C" int socks_gram_sym(const double* Z, int64_t n, int d, double scale, int divide, double diag_add, double* G,
                              int64_t ldg, socks_stream_t stream) {
    int rc = gram_sym_impl(SOCKS_KERNEL_RBF, Z, n, d, scale, divide, diag_add, G, ldg, stream);
    return rc < -1 ? rc + 1 : rc;
}

static int check_family(int kind, double scale, int divide, double degree, int argk) {
    SOCKS_CHECK_ARG(kind >= SOCKS_KERNEL_RBF && kind <= SOCKS_KERNEL_LAPLACIAN, 1);
    SOCKS_CHECK_ARG(scale == scale && (!divide || scale != 0.0), argk);
    SOCKS_CHECK_ARG(kind != SOCKS_KERNEL_POLY || degree == degree, argk + 3);
    return SOCKS_OK;
}

extern "C" int socks_gram_family(int kind, const double* X, int64_t n, const double* Y, int64_t m, int d, double scale,
                                 int divide, double coef0, double degree, const double* rowscale, double* K,
                                 int64_t ldk, socks_stream_t stream) {
    int rc = check_family(kind, scale, divide, degree, 7);
    if (rc) return rc;
    SOCKS_CHECK_ARG(n >= 0 && n <= 65535 * 65535LL, 3);
    SOCKS_CHECK_ARG(m >= 0, 5);
    SOCKS_CHECK_ARG(d >= 1 && d <= SOCKS_MAX_DIM, 6);
    SOCKS_CHECK_ARG(ldk >= m, 13);
    if (n == 0 || m == 0) return SOCKS_OK;
    SOCKS_CHECK_ARG(X && Y && K, 2);
    if (kind <= SOCKS_KERNEL_ABEL) {
        rc = gram_kernel_impl(kind, X, n, Y, m, d, scale, divide, rowscale, K, ldk, stream);
        return rc;
    }
    const KFamily f{kind, make_kscale(scale, divide), coef0, degree};
    cudaStream_t st = (cudaStream_t)stream;
    for (int64_t r0 = 0; r0 < n; r0 += 65535) {  // grid.y limit
        const int64_t rows = (n - r0 < 65535) ? (n - r0) : 65535;
        gram_family_kernel<<<dim3((unsigned)ceil_div(m, 256), (unsigned)rows), 256, 0, st>>>(
            f, X + r0 * d, rows, Y, m, d, rowscale ? rowscale + r0 : nullptr, K + r0 * ldk, ldk);
    }
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

extern "C" int socks_gram_sym_family(int kind, const double* X, int64_t n, int dx, const double* U, int du,
                                     double scale, int divide, double coef0, double degree, double diag_add,
                                     double* G, int64_t ldg, socks_stream_t stream) {
    int rc = check_family(kind, scale, divide, degree, 7);
    if (rc) return rc;
    SOCKS_CHECK_ARG(X != nullptr, 2);
    SOCKS_CHECK_ARG(n >= 1, 3);
    SOCKS_CHECK_ARG(dx >= 1 && dx <= SOCKS_MAX_DIM, 4);
    SOCKS_CHECK_ARG(U == nullptr || (du >= 1 && du <= SOCKS_MAX_DIM), 6);
    const int64_t np = socks_padded_dim(n);
    SOCKS_CHECK_ARG(np <= 65535, 3);
    SOCKS_CHECK_ARG(G != nullptr && (reinterpret_cast<uintptr_t>(G) & 15) == 0, 12);
    SOCKS_CHECK_ARG(ldg >= np && ldg % 2 == 0, 13);
    const KFamily f{kind, make_kscale(scale, divide), coef0, degree};
    gram_sym_family_kernel<<<dim3((unsigned)ceil_div(np, 256), (unsigned)np), 256, 0, (cudaStream_t)stream>>>(
        f, X, n, dx, U, du, diag_add, G, ldg, np);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}
