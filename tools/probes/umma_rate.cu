// Rate probe: clocks per tcgen05.mma kind::i8 (M = 128, K = 32) for N in {64, 128, 256}, with the A operand read from
// shared memory (SS form) or from tensor memory (TS form).  One CTA per SM, one issuing thread; the accumulator
// pattern is the digit scheme's (several accumulators, back-to-back accumulating MMAs).  Timing only: operands are noise.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I socks_b200/csrc -o tools/probes/umma_rate tools/probes/umma_rate.cu
#include <cstdio>
#include <vector>
#include "umma.cuh"
using namespace socks;

__device__ __forceinline__ void mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    const uint32_t zero = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc), "r"(zero), "r"(zero), "r"(zero), "r"(zero)
        : "memory");
}

// mode 0: SS; mode 1: TS (A in TMEM columns 448..); mode 2: SS but all MMAs share ONE A tile address (tests operand caching)
// mode 3: TS with the D window SHIFTED by 16 columns from one MMA to the next (overlapping accumulator windows, 8 per K-block)
// mode 4: TS, 8 MMAs per K-block all on the SAME D window (reference for mode 3)
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
    if (threadIdx.x < 32) umma::tmem_alloc(&slot, 512);
    if (threadIdx.x == 32) {
        umma::mbar_init(&bar, 1);
        umma::mbar_fence_init();
    }
    umma::fence_proxy_async();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = slot;
    constexpr int NACC = (N == 64) ? 7 : (N == 128 ? 3 : 1);
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma::idesc_i8(128, N, 1, 1);
        const uint32_t sA = umma::smem_u32(smem), sB = sA + 32 * 1024;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 28; ++j) {
                const int s = j % 7, t = (j * 3) % 7;
                const uint64_t da = umma::smem_desc_kmajor_noswizzle(sA + (MODE == 2 ? 0 : s) * 4096, 128, 256);
                const uint64_t db = umma::smem_desc_kmajor_noswizzle(sB + t * (N * 32), 128, 256);
                const uint32_t d = tmem + (uint32_t)((j % NACC) * N);
                if (MODE == 3) mma_i8_ts(tmem + 16 * (j % 8), tmem + 448 + (j % 8) * 8, db, idesc, 1u);
                else if (MODE == 4) mma_i8_ts(tmem, tmem + 448 + (j % 8) * 8, db, idesc, 1u);
                else if (MODE == 1) mma_i8_ts(d, tmem + 448 + s * 8, db, idesc, (it | j) ? 1u : 0u);
                else umma::mma_i8(d, da, db, idesc, (it | j) ? 1u : 0u);
            }
        }
        umma::mma_commit(&bar);
        umma::mbar_wait(&bar, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    umma::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 512);
}

template <int N, int MODE>
void run(const char* name, long long* dout) {
    const int iters = 2000;
    cudaFuncSetAttribute(rate_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    std::vector<long long> h(148);
    for (int rep = 0; rep < 2; ++rep) {
        rate_kernel<N, MODE><<<148, 128, 96 * 1024>>>(iters, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("%s: %s\n", name, cudaGetErrorString(e));
            return;
        }
    }
    cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    const double per = (double)mx / (iters * 28.0);
    printf("{\"case\": \"%s\", \"N\": %d, \"clk_per_mma\": %.2f, \"ideal_math_clk\": %.1f, \"mac_per_clk_per_sm\": %.0f}\n", name, N, per,
           128.0 * N * 32 / 8192.0, 128.0 * N * 32 / per);
}

int main() {
    long long* dout;
    cudaMalloc(&dout, 148 * 8);
    run<64, 0>("SS (A and B from shared memory)", dout);
    run<128, 0>("SS (A and B from shared memory)", dout);
    run<256, 0>("SS (A and B from shared memory)", dout);
    run<64, 2>("SS, one A tile for every MMA", dout);
    run<64, 1>("TS (A from tensor memory)", dout);
    run<128, 1>("TS (A from tensor memory)", dout);
    run<128, 3>("TS, D window shifted by 16 columns per MMA (overlapping windows)", dout);
    run<128, 4>("TS, same D window for every MMA", dout);
    return 0;
}
