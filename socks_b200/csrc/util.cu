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
