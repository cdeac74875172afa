// Pipeline probe for a two-pass N = 128 digit GEMM: loader thread (bulk copies of digit slices from an L2-resident operand
// buffer) + MMA thread (tcgen05.mma kind::i8, M = 128, N = 128), 148 CTAs, no epilogue.  Question: does L2 -> SM bandwidth
// sustain pass A (slices 0..3 of A and B: 32 KB per K-block, 10 MMAs) and pass B (all 7 slices: 56 KB, 18 MMAs)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I socks_b200/csrc -o tools/probes/umma_pipe tools/probes/umma_pipe.cu
#include <cstdio>
#include <vector>
#include "umma.cuh"
using namespace socks;

constexpr int SLICE = 4096;           // 128 rows x 32 B
constexpr int TILE_KB = 7 * SLICE;    // one K-block of one 128-row tile: 7 digit slices

__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            umma::smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(umma::smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     umma::smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// MC = 1: cluster of two CTAs along x sharing the A tile: each loads half of the A slices and multicasts them to both
template <int NS, int STAGES, int MC = 0>         // NS slices of each operand per K-block (4: pass A, 7: pass B)
__global__ void __launch_bounds__(128, 1) pipe_kernel(const uint8_t* __restrict__ buf, int kblocks, int tiles, int col_tiles,
                                                       long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int STAGE = 2 * NS * SLICE;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
    uint64_t* empty = full + STAGES;
    uint64_t* done = empty + STAGES;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) umma::tmem_alloc(&slot, 512);
    if (threadIdx.x == 32) {
        for (int i = 0; i < STAGES; ++i) {
            umma::mbar_init(full + i, 1);
            umma::mbar_init(empty + i, MC == 1 ? 2 : 1);
        }
        umma::mbar_init(done, 1);
        umma::mbar_fence_init();
    }
    umma::tc_fence_before();
    __syncthreads();
    if (MC == 1) cluster_sync_all();  // the peer's barriers exist before anything is multicast to them
    umma::tc_fence_after();
    const uint32_t tmem = slot;
    const long long t0 = clock64();
    const int rank = MC == 1 ? (int)(blockIdx.x & 1) : 0;
    const int a_tile = MC == 1 ? (int)(blockIdx.x >> 1) % 8 : blockIdx.x / col_tiles, b_tile = 8 + blockIdx.x % col_tiles;
    if (warp == 2 && lane == 0) {
        uint32_t it = 0;
        for (int t = 0; t < tiles; ++t) {
            const uint8_t* a_src = buf + (size_t)((a_tile + 2 * t) % 160) * kblocks * TILE_KB;
            const uint8_t* b_src = buf + (size_t)((b_tile + 3 * t) % 160) * kblocks * TILE_KB;
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                const uint32_t st = it % STAGES, round = it / STAGES;
                if (round > 0) umma::mbar_wait(empty + st, (round - 1) & 1u);
                uint8_t* s = smem + st * STAGE;
                umma::mbar_arrive_expect_tx(full + st, STAGE);
                if (MC == 1) {
                    constexpr int HALF = NS * SLICE / 2;  // multiple of 16 bytes
                    bulk_g2s_mc(s + rank * HALF, a_src + (size_t)kb * TILE_KB + rank * HALF, HALF, full + st, (uint16_t)3);
                } else if (MC == 2) {  // one bulk copy per digit slice (4 KB) instead of one per operand
#pragma unroll
                    for (int q = 0; q < NS; ++q) {
                        umma::bulk_g2s(s + q * SLICE, a_src + (size_t)kb * TILE_KB + q * SLICE, SLICE, full + st);
                        umma::bulk_g2s(s + (NS + q) * SLICE, b_src + (size_t)kb * TILE_KB + q * SLICE, SLICE, full + st);
                    }
                } else {
                    umma::bulk_g2s(s, a_src + (size_t)kb * TILE_KB, NS * SLICE, full + st);
                }
                if (MC != 2) umma::bulk_g2s(s + NS * SLICE, b_src + (size_t)kb * TILE_KB, NS * SLICE, full + st);
            }
        }
    } else if (warp == 3 && lane == 0) {
        const uint32_t idesc = umma::idesc_i8(128, 128, 1, 1);
        uint32_t it = 0;
        for (int t = 0; t < tiles; ++t) {
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                const uint32_t st = it % STAGES, round = it / STAGES;
                umma::mbar_wait(full + st, round & 1u);
                umma::tc_fence_after();
                const uint32_t sA = umma::smem_u32(smem + st * STAGE), sB = sA + NS * SLICE;
                if (NS == 4) {  // groups 0..3: pairs with a + b <= 3
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; a + b < 4; ++b)
                            umma::mma_i8(tmem + (a + b) * 128, umma::smem_desc_kmajor_noswizzle(sA + a * SLICE, 128, 256),
                                         umma::smem_desc_kmajor_noswizzle(sB + b * SLICE, 128, 256), idesc, (kb | a) ? 1u : 0u);
                } else {  // groups 4..6
#pragma unroll
                    for (int a = 0; a < 7; ++a)
#pragma unroll
                        for (int b = 0; b < 7; ++b)
                            if (a + b >= 4 && a + b <= 6)
                                umma::mma_i8(tmem + (a + b - 4) * 128, umma::smem_desc_kmajor_noswizzle(sA + a * SLICE, 128, 256),
                                             umma::smem_desc_kmajor_noswizzle(sB + b * SLICE, 128, 256), idesc,
                                             (kb > 0 || a > 0) ? 1u : 0u);
                }
                if (MC == 1) mma_commit_mc(empty + st, (uint16_t)3);
                else umma::mma_commit(empty + st);
            }
        }
        umma::mma_commit(done);
        umma::mbar_wait(done, 0);
        out[blockIdx.x] = clock64() - t0;
    }
    umma::tc_fence_before();
    __syncthreads();
    if (MC == 1) cluster_sync_all();  // nobody leaves while the peer may still write into its shared memory
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

template <int NS, int STAGES, int MC = 0>
void run(const uint8_t* buf, long long* dout, int col_tiles) {
    const int kblocks = 32, tiles = 20;
    const int smem = STAGES * 2 * NS * SLICE + (2 * STAGES + 1) * 8 + 64;
    cudaFuncSetAttribute(pipe_kernel<NS, STAGES, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<long long> h(148);
    for (int rep = 0; rep < 2; ++rep) {
        if (MC == 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(148);
            cfg.blockDim = dim3(128);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, pipe_kernel<NS, STAGES, MC>, buf, kblocks, tiles, col_tiles, dout);
        } else {
            pipe_kernel<NS, STAGES, MC><<<148, 128, smem>>>(buf, kblocks, tiles, col_tiles, dout);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("error: %s\n", cudaGetErrorString(e));
            return;
        }
    }
    cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    const int mmas = NS == 4 ? 10 : 18;
    const double per_kb = (double)mx / (tiles * kblocks);
    printf("{\"mode(0 plain, 1 multicast pairs, 2 one copy per slice)\": %d, \"slices\": %d, \"stages\": %d, \"col_tiles_per_row_tile\": %d, \"stage_kb\": %d, \"clk_per_kblock\": %.0f, \"ideal_clk\": %d, "
           "\"bytes_per_clk_per_sm\": %.1f, \"mma_efficiency\": %.3f}\n",
           MC, NS, STAGES, col_tiles, 2 * NS * 4, per_kb, mmas * 64, 2.0 * NS * SLICE / per_kb, mmas * 64 / per_kb);
}

int main() {
    uint8_t* buf;
    const size_t bytes = (size_t)160 * 32 * TILE_KB;  // 160 tiles x K = 1024: 147 MB (does not fit L2 entirely; reuse across CTAs does)
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0x11, bytes);
    long long* dout;
    cudaMalloc(&dout, 148 * 8);
    run<4, 6>(buf, dout, 74);
    run<7, 3>(buf, dout, 74);
    run<4, 6>(buf, dout, 37);
    run<7, 3>(buf, dout, 37);
    run<4, 4>(buf, dout, 74);
    run<4, 5, 1>(buf, dout, 74);
    run<7, 3, 1>(buf, dout, 74);
    run<4, 5, 2>(buf, dout, 74);
    run<7, 3, 2>(buf, dout, 74);
    return 0;
}
