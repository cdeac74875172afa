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
