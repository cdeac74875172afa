// Error string, version and padding helper of libsocksb200.
#include <stdarg.h>

#include "common.cuh"

namespace socks {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace socks

extern "C" int socks_version(void) { return 100; }
extern "C" const char* socks_last_error(void) { return socks::g_err; }
extern "C" int64_t socks_padded_dim(int64_t n) { return n <= 0 ? 0 : socks::round_up(n, SOCKS_TILE); }

// rows x width_bytes strided copy in ONE enqueue (cudaMemcpy2DAsync): used for the device-to-pinned-host read of a
// column block of the (T, M) result, which is T row segments.  kind: 1 = host->device, 2 = device->host, 3 = d2d.
extern "C" int socks_copy2d(void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width_bytes,
                            int64_t rows, int kind, socks_stream_t stream) {
    SOCKS_CHECK_ARG(dst != nullptr, 1);
    SOCKS_CHECK_ARG(src != nullptr, 3);
    SOCKS_CHECK_ARG(width_bytes >= 0 && dpitch >= width_bytes && spitch >= width_bytes, 5);
    SOCKS_CHECK_ARG(rows >= 0, 6);
    SOCKS_CHECK_ARG(kind >= 1 && kind <= 3, 7);
    if (rows == 0 || width_bytes == 0) return SOCKS_OK;
    SOCKS_CUDA(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width_bytes, (size_t)rows,
                                 (cudaMemcpyKind)kind, (cudaStream_t)stream));
    return SOCKS_OK;
}

// out[0] = min_i |A[i][i]|, out[1] = max_i |A[i][i]| over the first n diagonal entries.  After a factorisation the
// diagonal of the factor gives cond_2(G) >= (max l_ii / min l_ii)^2: the wrappers use it to decide whether the inverse
// factor's rows span more dynamic range than the integer digit GEMMs carry.
namespace socks {
__global__ void __launch_bounds__(1024) diag_range_kernel(const double* __restrict__ A, int64_t n, int64_t ld,
                                                          double* __restrict__ out) {
    __shared__ double slo[32], shi[32];
    double lo = INFINITY, hi = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = fabs(A[i * ld + i]);
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
    for (int off = 16; off > 0; off >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if ((threadIdx.x & 31) == 0) {
        slo[threadIdx.x >> 5] = lo;
        shi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            lo = fmin(lo, slo[w]);
            hi = fmax(hi, shi[w]);
        }
        out[0] = lo;
        out[1] = hi;
    }
}
}  // namespace socks

extern "C" int socks_diag_range(const double* A, int64_t n, int64_t ld, double* out2, socks_stream_t stream) {
    SOCKS_CHECK_ARG(A != nullptr, 1);
    SOCKS_CHECK_ARG(n >= 1, 2);
    SOCKS_CHECK_ARG(ld >= n, 3);
    SOCKS_CHECK_ARG(out2 != nullptr, 4);
    socks::diag_range_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(A, n, ld, out2);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

