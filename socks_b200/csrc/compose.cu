// Composite entry points of the C ABI: the call sequences of the reference's algorithm classes as ONE enqueue each,
// so that a non-Python binder does not have to re-implement them, plus the context / collective helpers of
// SURVEY.md 8(b):
//   socks_sr_fit          gym_socks/algorithms/reach/kernel_sr.py:286-309      (beta + in-place recursion [+ pack])
//   socks_srmax_fit       gym_socks/algorithms/reach/kernel_sr_max.py:323-366
//   socks_srmax_predict   gym_socks/algorithms/reach/kernel_sr_max.py:368-384
//   socks_fwd_scores      gym_socks/algorithms/control/kernel_control_fwd.py:217-232
//   socks_fwd_policy      ... + control/common.py:96-100,166-174 (heuristic argmin)
//   socks_median_sq_dist  gym_socks/kernel/metrics.py:130-131, :199-200 (sigma = median of the distance matrix)
//   socks_ctx_*           device-owned stream + grow-only workspace for binders without an allocator
//   socks_comm_* / socks_bcast / socks_allgather   NCCL (bound at run time with dlopen: no link-time dependency)
// The Python wrappers call these same entry points (socks_b200/algorithms/...), so they are what the tests exercise.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <mutex>

#include "common.cuh"

using namespace socks;

namespace socks {

static inline int64_t even64(int64_t n) { return n + (n & 1); }
static inline int64_t align2(int64_t doubles) { return round_up(doubles, 2); }

// ---- exact median of the non-negative values D[0 .. cnt): MSB-first radix select on the IEEE bit patterns --------
// (non-negative doubles order like their bit patterns).  Both middle order statistics are selected in the same 8
// passes; np.median = mean of the two for an even count.  sel[0..1] = prefixes, sel[2..3] = remaining ranks.
__global__ void median_hist_kernel(const double* __restrict__ D, int64_t rows, int64_t cols, int64_t ld, int pass,
                                   const unsigned long long* __restrict__ sel, unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[2][256];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    const unsigned long long p0 = sel[0], p1 = sel[1];
    const int64_t cnt = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < cnt; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / cols, c = e - r * cols;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(D[r * ld + c]);
        const unsigned long long hi = (pass == 0) ? 0ull : (bits >> (shift + 8));
        const unsigned int b = (unsigned int)((bits >> shift) & 255ull);
        if (hi == p0) atomicAdd(&sh[0][b], 1u);
        if (hi == p1) atomicAdd(&sh[1][b], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
        const unsigned int v = (&sh[0][0])[i];
        if (v) atomicAdd(hist + i, v);
    }
}

__global__ void median_pick_kernel(unsigned long long* __restrict__ sel, unsigned int* __restrict__ hist, int pass,
                                   int64_t cnt, double* __restrict__ result) {
    if (threadIdx.x < 2) {
        const int s = threadIdx.x;
        unsigned long long k = sel[2 + s], cum = 0;
        int b = 0;
        for (; b < 255; ++b) {
            const unsigned long long h = hist[s * 256 + b];
            if (cum + h > k) break;
            cum += h;
        }
        sel[2 + s] = k - cum;
        sel[s] = (sel[s] << 8) | (unsigned long long)b;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x) hist[i] = 0;
    if (pass == 7 && threadIdx.x == 0) {
        const double lo = __longlong_as_double((long long)sel[0]), hi = __longlong_as_double((long long)sel[1]);
        *result = (cnt & 1) ? hi : (lo + hi) / 2;  // numpy: mean of the two middle values
    }
}

__global__ void median_init_kernel(unsigned long long* sel, unsigned int* hist, int64_t cnt) {
    for (int i = threadIdx.x; i < 512; i += blockDim.x) hist[i] = 0;
    if (threadIdx.x == 0) {
        sel[0] = 0;
        sel[1] = 0;
        sel[2] = (unsigned long long)((cnt & 1) ? cnt / 2 : cnt / 2 - 1);  // 0-based ranks of the two middle values
        sel[3] = (unsigned long long)(cnt / 2);
    }
}

__global__ void zero_kernel(double* __restrict__ p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0.0;
}

}  // namespace socks

#define SOCKS_TRY(call)        \
    do {                       \
        int rc__ = (call);     \
        if (rc__) return rc__; \
    } while (0)

// ================================================================================================ median
extern "C" int64_t socks_median_work_bytes(int64_t n, int64_t m) { return (n * m + 2) * 8 + 4096; }

extern "C" int socks_median_sq_dist(const double* X, int64_t n, const double* Y, int64_t m, int d, int squared,
                                    double* work, double* result_dev, socks_stream_t stream) {
    SOCKS_CHECK_ARG(X != nullptr, 1);
    SOCKS_CHECK_ARG(n >= 1, 2);
    SOCKS_CHECK_ARG(Y != nullptr, 3);
    SOCKS_CHECK_ARG(m >= 1, 4);
    SOCKS_CHECK_ARG(d >= 1 && d <= SOCKS_MAX_DIM, 5);
    SOCKS_CHECK_ARG(work != nullptr && (reinterpret_cast<uintptr_t>(work) & 15) == 0, 7);
    SOCKS_CHECK_ARG(result_dev != nullptr, 8);
    cudaStream_t st = (cudaStream_t)stream;
    double* D = work;
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(work + align2(n * m));
    unsigned int* hist = reinterpret_cast<unsigned int*>(sel + 4);
    SOCKS_TRY(squared ? socks_sqdist(X, n, Y, m, d, D, m, stream) : socks_dist(X, n, Y, m, d, D, m, stream));
    const int64_t cnt = n * m;
    median_init_kernel<<<1, 256, 0, st>>>(sel, hist, cnt);
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(cnt, 256 * 8), 148 * 16);
    for (int pass = 0; pass < 8; ++pass) {
        median_hist_kernel<<<grid, 256, 0, st>>>(D, n, m, m, pass, sel, hist);
        median_pick_kernel<<<1, 256, 0, st>>>(sel, hist, pass, cnt, result_dev);
    }
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}

// ================================================================================================ KernelSR.fit
extern "C" int64_t socks_sr_fit_work_bytes(int64_t n) {
    return (n * even64(n) + (n + 4)) * 8 + socks_reduce_work_bytes(n);
}

extern "C" int socks_sr_fit(const double* X, const double* Y, const double* w, int64_t n, int d, double scale,
                            int divide, const double* tubes, int problem, int T, double* vf, int64_t ldvf,
                            double* pack, double* work, socks_stream_t stream) {
    SOCKS_CHECK_ARG(X != nullptr, 1);
    SOCKS_CHECK_ARG(Y != nullptr, 2);
    SOCKS_CHECK_ARG(w != nullptr, 3);
    SOCKS_CHECK_ARG(n >= 1, 4);
    SOCKS_CHECK_ARG(T >= 1, 10);
    SOCKS_CHECK_ARG(vf != nullptr && ldvf >= n, 11);
    SOCKS_CHECK_ARG(work != nullptr && (reinterpret_cast<uintptr_t>(work) & 15) == 0, 14);
    const int64_t ldb = even64(n);
    double* B = work;                  // un-normalised beta = diag(w) K(X, Y), resident for the T-1 passes
    double* rwork = work + n * ldb;
    SOCKS_TRY(socks_gram_rbf(X, n, Y, n, d, scale, divide, w, B, ldb, stream));
    SOCKS_TRY(socks_sr_recursion(B, n, n, ldb, Y, d, tubes, problem, T, vf, ldvf, vf, ldvf, rwork, stream));
    if (pack) SOCKS_TRY(socks_sr_pack(X, w, vf, ldvf, n, d, scale, divide, T, pack, stream));
    return SOCKS_OK;
}

// ================================================================================================ KernelMaximalSR
static inline int64_t pad128(int64_t v) { return round_up(v, 128); }

extern "C" int64_t socks_srmax_fit_work_bytes(int64_t n, int P) {
    const int64_t np = socks_padded_dim(n), ldbm = pad128(P);
    return (np * np + 3 * np * ldbm) * 8;
}

extern "C" int socks_srmax_fit(const double* X, const double* Y, const double* U, const double* A, const double* w,
                               int64_t n, int d, int du, int P, double scale, int divide, const double* tubes,
                               int problem, int T, double* CUA, double* vf, int64_t ldvf, double* work,
                               socks_stream_t stream) {
    SOCKS_CHECK_ARG(X && Y && U && A && w, 1);
    SOCKS_CHECK_ARG(n >= 1, 6);
    SOCKS_CHECK_ARG(P >= 1, 9);
    SOCKS_CHECK_ARG(T >= 1, 14);
    SOCKS_CHECK_ARG(CUA != nullptr, 15);
    SOCKS_CHECK_ARG(vf != nullptr && ldvf >= n, 16);
    SOCKS_CHECK_ARG(work != nullptr && (reinterpret_cast<uintptr_t>(work) & 15) == 0, 18);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t np = socks_padded_dim(n), ldbm = pad128(P);
    double* Kmat = work;  // K(X, Y) as the GEMM's left operand, zero padding
    double* Bm = Kmat + np * np;
    double* Cm = Bm + np * ldbm;
    double* Den = Cm + np * ldbm;
    SOCKS_TRY(socks_gram_rbf(U, n, A, P, du, scale, divide, nullptr, CUA, P, stream));  // kernel_sr_max.py:332
    zero_kernel<<<148 * 8, 256, 0, st>>>(Kmat, np * np);
    SOCKS_CHECK_LAUNCH();
    SOCKS_TRY(socks_gram_rbf(X, n, Y, n, d, scale, divide, nullptr, Kmat, np, stream));
    // vf[T-1] = 1_target(Y) (:59); the first pack below reads it
    SOCKS_TRY(socks_srmax_step(Cm, ldbm, nullptr, 0, n, P, 1, 0, Y, d, tubes, problem, T, vf, ldvf, 1, stream));
    if (T > 1) {  // the normaliser sum_i w_i C_ij CUA_ik does not depend on t: one GEMM for all steps
        SOCKS_TRY(socks_srmax_pack(CUA, n, P, w, nullptr, ldvf, 0, 0, 1, Bm, np, ldbm, stream));
        SOCKS_TRY(socks_gemm_tn(Kmat, np, Bm, ldbm, Den, ldbm, np, ldbm, np, stream));
    }
    for (int t = T - 2; t >= 0; --t) {  // numerator row vf[t+1] * w (:65-72)
        SOCKS_TRY(socks_srmax_pack(CUA, n, P, w, vf, ldvf, t, 1, 1, Bm, np, ldbm, stream));
        SOCKS_TRY(socks_gemm_tn(Kmat, np, Bm, ldbm, Cm, ldbm, np, ldbm, np, stream));
        SOCKS_TRY(socks_srmax_step(Cm, ldbm, Den, ldbm, n, P, 2, t, Y, d, tubes, problem, T, vf, ldvf, 0, stream));
    }
    return SOCKS_OK;
}

extern "C" int64_t socks_srmax_predict_work_bytes(int64_t n, int P, int T, int64_t chunk) {
    const int64_t np = socks_padded_dim(n), ldbm = pad128((int64_t)T * P), mc = pad128(chunk);
    return (np * ldbm + np * mc + mc * ldbm) * 8;
}

extern "C" int socks_srmax_predict(const double* X, const double* w, const double* vf, int64_t ldvf,
                                   const double* CUA, int64_t n, int d, int P, double scale, int divide,
                                   const double* pts, int64_t m, const double* tubes, int problem, int T, double* out,
                                   int64_t ldout, double* work, int64_t chunk, socks_stream_t stream) {
    SOCKS_CHECK_ARG(X && w && vf && CUA, 1);
    SOCKS_CHECK_ARG(n >= 1, 6);
    SOCKS_CHECK_ARG(P >= 1, 8);
    SOCKS_CHECK_ARG(m >= 0, 12);
    SOCKS_CHECK_ARG(T >= 1, 15);
    SOCKS_CHECK_ARG(ldout >= m, 17);
    SOCKS_CHECK_ARG(work != nullptr && (reinterpret_cast<uintptr_t>(work) & 15) == 0, 18);
    SOCKS_CHECK_ARG(chunk >= 1, 19);
    if (m == 0) return SOCKS_OK;
    SOCKS_CHECK_ARG(pts != nullptr && out != nullptr, 11);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t np = socks_padded_dim(n), ldbm = pad128((int64_t)T * P), mc = pad128(std::min(chunk, m));
    double* Bm = work;
    double* Kmat = Bm + np * ldbm;
    double* Cm = Kmat + np * mc;
    // all rows at once: r = 0 normaliser, r = 1..T-1 time steps 0..T-2
    SOCKS_TRY(socks_srmax_pack(CUA, n, P, w, vf, ldvf, 0, 0, T, Bm, np, ldbm, stream));
    zero_kernel<<<148 * 8, 256, 0, st>>>(Kmat, np * mc);
    SOCKS_CHECK_LAUNCH();
    for (int64_t s = 0; s < m; s += mc) {
        const int64_t cnt = std::min(mc, m - s);
        if (cnt < mc) {  // ragged tail: stale columns of the previous chunk must not reach the GEMM as NaN/garbage
            zero_kernel<<<148 * 8, 256, 0, st>>>(Kmat, np * mc);
            SOCKS_CHECK_LAUNCH();
        }
        const double* p = pts + s * d;
        SOCKS_TRY(socks_gram_rbf(X, n, p, cnt, d, scale, divide, nullptr, Kmat, mc, stream));
        SOCKS_TRY(socks_gemm_tn(Kmat, mc, Bm, ldbm, Cm, ldbm, mc, ldbm, np, stream));
        SOCKS_TRY(socks_srmax_step(Cm, ldbm, nullptr, 0, cnt, P, T, 0, p, d, tubes, problem, T, out + s, ldout, 1, stream));
    }
    return SOCKS_OK;
}

// ================================================================================================ KernelControlFwd
extern "C" int64_t socks_fwd_scores_work_bytes(int64_t n, int P) {
    return 4 * (n + 2) * 8 + 2 * socks_reduce_work_bytes(std::max<int64_t>(n, P));
}

// scores[r][a] = sum_i (W c_r)_i k(x_i, s) CUA[i][a]   (r = 0 cost, r = 1 constraint), kernel_control_fwd.py:217-232
extern "C" int socks_fwd_scores(const double* W, int64_t ldw, const double* X, const double* CUA, int64_t n, int d,
                                int P, double scale, int divide, const double* state, const double* c, int64_t ldc,
                                int R, double* scores, double* work, socks_stream_t stream) {
    SOCKS_CHECK_ARG(W != nullptr && ldw >= n, 1);
    SOCKS_CHECK_ARG(X && CUA, 3);
    SOCKS_CHECK_ARG(n >= 1, 5);
    SOCKS_CHECK_ARG(P >= 1, 7);
    SOCKS_CHECK_ARG(state != nullptr, 10);
    SOCKS_CHECK_ARG(c != nullptr && (ldc >= n || R == 1), 11);
    SOCKS_CHECK_ARG(R == 1 || R == 2, 13);
    SOCKS_CHECK_ARG(scores != nullptr, 14);
    SOCKS_CHECK_ARG(work != nullptr && (reinterpret_cast<uintptr_t>(work) & 15) == 0, 15);
    const int64_t ldz = even64(n);
    double* z = work;
    double* q = z + 2 * ldz;
    double* rwork = q + 2 * ldz;
    SOCKS_TRY(socks_gemv_t(W, n, n, ldw, c, ldc, R, z, ldz, rwork, stream));  // z = W c (W symmetric)
    SOCKS_TRY(socks_fwd_weights(X, n, d, scale, divide, state, z, ldz, R, q, ldz, stream));
    SOCKS_TRY(socks_gemv_t(CUA, n, P, P, q, ldz, R, scores, P, rwork, stream));
    return SOCKS_OK;
}

extern "C" int socks_fwd_policy(const double* W, int64_t ldw, const double* X, const double* CUA, int64_t n, int d,
                                int P, double scale, int divide, const double* state, const double* c, int64_t ldc,
                                int R, double* scores, double* work, int* idx_dev, socks_stream_t stream) {
    SOCKS_CHECK_ARG(idx_dev != nullptr, 16);
    SOCKS_TRY(socks_fwd_scores(W, ldw, X, CUA, n, d, P, scale, divide, state, c, ldc, R, scores, work, stream));
    return socks_argmin_masked(scores, R == 2 ? scores + P : nullptr, P, idx_dev, stream);
}

// ================================================================================================ context
struct socks_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    void* arena = nullptr;
    int64_t arena_bytes = 0;
    void* comm = nullptr;  // ncclComm_t
    int rank = 0, nranks = 1;
};

extern "C" int socks_ctx_create(int device, socks_ctx** out) {
    SOCKS_CHECK_ARG(device >= 0, 1);
    SOCKS_CHECK_ARG(out != nullptr, 2);
    SOCKS_CUDA(cudaSetDevice(device));
    socks_ctx* c = new socks_ctx();
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        set_error("socks_ctx_create: %s", cudaGetErrorString(e));
        return SOCKS_ECUDA;
    }
    *out = c;
    return SOCKS_OK;
}

extern "C" socks_stream_t socks_ctx_stream(socks_ctx* c) { return c ? (socks_stream_t)c->stream : nullptr; }

// Grow-only device workspace owned by the context; the pointer stays valid until a LARGER request (which synchronises the
// context stream, frees and reallocates) or socks_ctx_destroy.
extern "C" int socks_ctx_workspace(socks_ctx* c, int64_t bytes, void** ptr) {
    SOCKS_CHECK_ARG(c != nullptr, 1);
    SOCKS_CHECK_ARG(bytes >= 0, 2);
    SOCKS_CHECK_ARG(ptr != nullptr, 3);
    if (bytes > c->arena_bytes) {
        SOCKS_CUDA(cudaSetDevice(c->device));
        SOCKS_CUDA(cudaStreamSynchronize(c->stream));
        if (c->arena) SOCKS_CUDA(cudaFree(c->arena));
        c->arena = nullptr;
        c->arena_bytes = 0;
        SOCKS_CUDA(cudaMalloc(&c->arena, (size_t)bytes));
        c->arena_bytes = bytes;
    }
    *ptr = c->arena;
    return SOCKS_OK;
}

// ---- NCCL, bound at run time -------------------------------------------------------------------------------
namespace {
struct NcclId {
    char internal[128];
};
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // torch ships libnccl.so.2 and has normally loaded it already: dlopen by SONAME returns that same library
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        api.GetUniqueId = (int (*)(NcclId*))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(api.handle, "ncclCommInitRank");
        api.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.handle, "ncclBroadcast");
        api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(api.handle, "ncclAllGather");
        api.CommDestroy = (int (*)(void*))dlsym(api.handle, "ncclCommDestroy");
        api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
    });
    const bool ok = api.handle && api.GetUniqueId && api.CommInitRank && api.Broadcast && api.AllGather && api.CommDestroy;
    return ok ? &api : nullptr;
}

int nccl_fail(const char* what, NcclApi* api, int rc) {
    set_error("%s: NCCL error %d (%s)", what, rc, (api && api->GetErrorString) ? api->GetErrorString(rc) : "?");
    return SOCKS_ECUDA + 1;
}
constexpr int kNcclChar = 0;  // ncclInt8 / ncclChar
}  // namespace

#define SOCKS_NCCL_API(api)                                                                      \
    NcclApi* api = nccl_api();                                                                   \
    if (!api) {                                                                                  \
        set_error("%s: libnccl.so.2 could not be loaded (dlopen): %s", __func__, dlerror());     \
        return SOCKS_ECUDA + 1;                                                                  \
    }

extern "C" int socks_comm_unique_id(void* id128) {
    SOCKS_CHECK_ARG(id128 != nullptr, 1);
    SOCKS_NCCL_API(api);
    const int rc = api->GetUniqueId(reinterpret_cast<NcclId*>(id128));
    return rc ? nccl_fail("socks_comm_unique_id", api, rc) : SOCKS_OK;
}

extern "C" int socks_comm_init(socks_ctx* c, int rank, int nranks, const void* id128) {
    SOCKS_CHECK_ARG(c != nullptr && c->comm == nullptr, 1);
    SOCKS_CHECK_ARG(nranks >= 1 && rank >= 0 && rank < nranks, 2);
    SOCKS_CHECK_ARG(id128 != nullptr, 4);
    SOCKS_NCCL_API(api);
    SOCKS_CUDA(cudaSetDevice(c->device));
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    const int rc = api->CommInitRank(&c->comm, nranks, id, rank);
    if (rc) return nccl_fail("socks_comm_init", api, rc);
    c->rank = rank;
    c->nranks = nranks;
    return SOCKS_OK;
}

// In-place broadcast of `bytes` bytes from `root` on the context stream (fit state X, w, V_fit -- or W -- SURVEY 8(e)).
extern "C" int socks_bcast(socks_ctx* c, void* ptr, int64_t bytes, int root) {
    SOCKS_CHECK_ARG(c != nullptr && c->comm != nullptr, 1);
    SOCKS_CHECK_ARG(ptr != nullptr, 2);
    SOCKS_CHECK_ARG(bytes >= 0, 3);
    SOCKS_CHECK_ARG(root >= 0 && root < c->nranks, 4);
    SOCKS_NCCL_API(api);
    const int rc = api->Broadcast(ptr, ptr, (size_t)bytes, kNcclChar, root, c->comm, c->stream);
    return rc ? nccl_fail("socks_bcast", api, rc) : SOCKS_OK;
}

// recv[r * bytes_per_rank ...] = rank r's send buffer (result columns of the test-point shards).
extern "C" int socks_allgather(socks_ctx* c, const void* send, void* recv, int64_t bytes_per_rank) {
    SOCKS_CHECK_ARG(c != nullptr && c->comm != nullptr, 1);
    SOCKS_CHECK_ARG(send != nullptr, 2);
    SOCKS_CHECK_ARG(recv != nullptr, 3);
    SOCKS_CHECK_ARG(bytes_per_rank >= 0, 4);
    SOCKS_NCCL_API(api);
    const int rc = api->AllGather(send, recv, (size_t)bytes_per_rank, kNcclChar, c->comm, c->stream);
    return rc ? nccl_fail("socks_allgather", api, rc) : SOCKS_OK;
}

extern "C" int socks_ctx_synchronize(socks_ctx* c) {
    SOCKS_CHECK_ARG(c != nullptr, 1);
    SOCKS_CUDA(cudaStreamSynchronize(c->stream));
    return SOCKS_OK;
}

extern "C" int socks_ctx_destroy(socks_ctx* c) {
    if (!c) return SOCKS_OK;
    cudaSetDevice(c->device);
    if (c->comm) {
        NcclApi* api = nccl_api();
        if (api) api->CommDestroy(c->comm);
    }
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->arena) cudaFree(c->arena);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return SOCKS_OK;
}
