// Integer tensor-core (tcgen05.mma kind::i8, TMEM accumulators) building blocks.
// Phase-1 check kernel: C (M x N, int32) = A (M x K, int8, K contiguous) * B (N x K, int8, K contiguous)^T, one CTA per
// 128 x 64 tile, one MMA (K = 32) per staged K-block.  Validates the descriptor / layout conventions of umma.cuh against
// an exact integer reference (tools/i8gemm_check.py) before anything is built on them.
#include "umma.cuh"

namespace socks {

constexpr int I8_BM = 128, I8_BN = 64, I8_KB = 32;

// Core-matrix layout of an (R rows x 32 bytes) K-major operand tile: [r / 8][k-chunk c (16 B)][r % 8][16 B].
__device__ __forceinline__ int core_offset(int r, int c) { return (r >> 3) * 256 + c * 128 + (r & 7) * 16; }

__global__ void __launch_bounds__(128)
    i8gemm_check_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ C, int M, int N,
                        int K, int variant) {
    __shared__ __align__(128) uint8_t sA[I8_BM * I8_KB];
    __shared__ __align__(128) uint8_t sB[I8_BN * I8_KB];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.y * I8_BM, n0 = blockIdx.x * I8_BN;
    if (warp == 0) umma::tmem_alloc(&tmem_base, 64);
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::mbar_fence_init();
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_base;
    const uint32_t idesc = umma::idesc_i8(I8_BM, I8_BN, 1, 1);
    const uint32_t lbo = variant ? 256 : 128, sbo = variant ? 128 : 256;
    uint32_t phase = 0;
    for (int k0 = 0; k0 < K; k0 += I8_KB) {
        for (int ch = tid; ch < I8_BM * 2; ch += 128) {
            const int r = ch >> 1, c = ch & 1;
            *reinterpret_cast<int4*>(sA + core_offset(r, c)) =
                *reinterpret_cast<const int4*>(A + (int64_t)(m0 + r) * K + k0 + c * 16);
        }
        for (int ch = tid; ch < I8_BN * 2; ch += 128) {
            const int r = ch >> 1, c = ch & 1;
            *reinterpret_cast<int4*>(sB + core_offset(r, c)) =
                *reinterpret_cast<const int4*>(B + (int64_t)(n0 + r) * K + k0 + c * 16);
        }
        umma::fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
        __syncthreads();
        if (tid == 0) {
            umma::tc_fence_after();
            const uint64_t da = umma::smem_desc_kmajor_noswizzle(umma::smem_u32(sA), lbo, sbo);
            const uint64_t db = umma::smem_desc_kmajor_noswizzle(umma::smem_u32(sB), lbo, sbo);
            umma::mma_i8(tmem, da, db, idesc, k0 > 0 ? 1u : 0u);
            umma::mma_commit(&bar);
        }
        umma::mbar_wait(&bar, phase);  // the MMA has consumed this K-block: shared memory may be overwritten
        phase ^= 1;
    }
    umma::tc_fence_after();
    for (int c0 = 0; c0 < I8_BN; c0 += 16) {
        uint32_t v[16];
        umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) C[(int64_t)(m0 + tid) * N + n0 + c0 + i] = (int32_t)v[i];
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

}  // namespace socks

using namespace socks;

extern "C" int socks_dev_i8gemm_check(const void* A, const void* B, void* C, int M, int N, int K, int variant,
                                      socks_stream_t stream) {
    SOCKS_CHECK_ARG(A && B && C, 1);
    SOCKS_CHECK_ARG(M > 0 && M % I8_BM == 0, 4);
    SOCKS_CHECK_ARG(N > 0 && N % I8_BN == 0, 5);
    SOCKS_CHECK_ARG(K > 0 && K % I8_KB == 0, 6);
    i8gemm_check_kernel<<<dim3(N / I8_BN, M / I8_BM), 128, 0, (cudaStream_t)stream>>>(
        (const int8_t*)A, (const int8_t*)B, (int32_t*)C, M, N, K, variant);
    SOCKS_CHECK_LAUNCH();
    return SOCKS_OK;
}
