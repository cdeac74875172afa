// Pipeline probe, 2-CTA form: tcgen05.mma.cta_group::2 kind::i8, M = 256 per CTA pair (each CTA: its 128 rows of A and
// HALF of the 128-row B tile), N = 128.  Question: with 25 % fewer bytes staged and 25 % fewer operand reads per SM, do
// pass A (digits 0..3, 10 MMAs per K-block) and pass B (7 digits, 18 MMAs) reach the MMA rate (640 / 1152 clocks)?
// Timing only (operands are noise).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I socks_b200/csrc
//   -o tools/probes/umma_pair tools/probes/umma_pair.cu
#include <cstdio>
#include <vector>
#include "umma.cuh"
using namespace socks;

constexpr int SLICE = 4096, HALF = 2048, TILE_KB = 7 * SLICE;

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_cta(const void* smem_ptr, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(umma::smem_u32(smem_ptr)), "r"(rank));
    return out;
}
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mma_i8_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        :
        : "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(z)
        : "memory");
}
__device__ __forceinline__ void commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     umma::smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

template <int NS, int STAGES>
__global__ void __launch_bounds__(128, 1) pair_kernel(const uint8_t* __restrict__ buf, int kblocks, int tiles, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int STAGE = NS * (SLICE + HALF);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
    uint64_t* empty = full + STAGES;
    uint64_t* peer_full = empty + STAGES;  // used on the leader: the peer's stage has landed
    uint64_t* done = peer_full + STAGES;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    if (threadIdx.x == 32) {
        for (int i = 0; i < STAGES; ++i) {
            umma::mbar_init(full + i, 1);
            umma::mbar_init(empty + i, 1);
            umma::mbar_init(peer_full + i, 1);
        }
        umma::mbar_init(done, 1);
        umma::mbar_fence_init();
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma::smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    umma::tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    umma::tc_fence_after();
    const uint32_t tmem = slot;
    const long long t0 = clock64();
    const int pair = blockIdx.x >> 1;
    const int a_tile = (2 * pair + (int)rank) % 16, b_tile = 16 + pair % 74;
    if (warp == 2 && lane == 0) {  // loader (both CTAs): own A rows, own half of B
        uint32_t it = 0;
        for (int t = 0; t < tiles; ++t) {
            const uint8_t* a_src = buf + (size_t)((a_tile + 2 * t) % 160) * kblocks * TILE_KB;
            const uint8_t* b_src = buf + (size_t)((b_tile + 3 * t) % 160) * kblocks * TILE_KB + rank * HALF;
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                const uint32_t st = it % STAGES, round = it / STAGES;
                if (round > 0) umma::mbar_wait(empty + st, (round - 1) & 1u);
                uint8_t* s = smem + st * STAGE;
                umma::mbar_arrive_expect_tx(full + st, STAGE);
                umma::bulk_g2s(s, a_src + (size_t)kb * TILE_KB, NS * SLICE, full + st);
#pragma unroll
                for (int q = 0; q < NS; ++q)
                    umma::bulk_g2s(s + NS * SLICE + q * HALF, b_src + (size_t)kb * TILE_KB + q * SLICE, HALF, full + st);
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 1) {  // peer: tell the leader when a stage has landed here
        uint32_t it = 0;
        for (int t = 0; t < tiles; ++t)
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                const uint32_t st = it % STAGES, round = it / STAGES;
                umma::mbar_wait(full + st, round & 1u);
                remote_arrive(map_to_cta(peer_full + st, 0));
            }
    } else if (warp == 3 && lane == 0 && rank == 0) {  // leader: MMA issue for the pair
        const uint32_t idesc = umma::idesc_i8(256, 128, 1, 1);
        uint32_t it = 0;
        for (int t = 0; t < tiles; ++t) {
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                const uint32_t st = it % STAGES, round = it / STAGES;
                umma::mbar_wait(full + st, round & 1u);
                umma::mbar_wait(peer_full + st, round & 1u);
                umma::tc_fence_after();
                const uint32_t sA = umma::smem_u32(smem + st * STAGE), sB = sA + NS * SLICE;
                if (NS == 4) {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; a + b < 4; ++b)
                            mma_i8_pair(tmem + (a + b) * 128, umma::smem_desc_kmajor_noswizzle(sA + a * SLICE, 128, 256),
                                        umma::smem_desc_kmajor_noswizzle(sB + b * HALF, 128, 256), idesc, (kb | a) ? 1u : 0u);
                } else {
#pragma unroll
                    for (int a = 0; a < 7; ++a)
#pragma unroll
                        for (int b = 0; b < 7; ++b)
                            if (a + b >= 4 && a + b <= 6)
                                mma_i8_pair(tmem + (a + b - 4) * 128, umma::smem_desc_kmajor_noswizzle(sA + a * SLICE, 128, 256),
                                            umma::smem_desc_kmajor_noswizzle(sB + b * HALF, 128, 256), idesc,
                                            (kb > 0 || a > 0) ? 1u : 0u);
                }
                commit_pair(empty + st);
            }
        }
        commit_pair(done);
    }
    if (threadIdx.x == 0) {
        umma::mbar_wait(done, 0);
        out[blockIdx.x] = clock64() - t0;
    }
    umma::tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int NS, int STAGES>
void run(const uint8_t* buf, long long* dout) {
    const int kblocks = 32, tiles = 20;
    const int smem = STAGES * NS * (SLICE + HALF) + (3 * STAGES + 1) * 8 + 64;
    cudaFuncSetAttribute(pair_kernel<NS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<long long> h(148);
    for (int rep = 0; rep < 2; ++rep) {
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
        cudaError_t e = cudaLaunchKernelEx(&cfg, pair_kernel<NS, STAGES>, buf, kblocks, tiles, dout);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
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
    printf("{\"form\": \"cta_group::2 (M = 256 per pair)\", \"slices\": %d, \"stages\": %d, \"stage_kb_per_sm\": %d, \"clk_per_kblock\": %.0f, "
           "\"ideal_clk\": %d, \"bytes_per_clk_per_sm\": %.1f, \"mma_efficiency\": %.3f}\n",
           NS, STAGES, NS * 6, per_kb, mmas * 64, (double)NS * (SLICE + HALF) / per_kb, mmas * 64 / per_kb);
}

int main() {
    uint8_t* buf;
    const size_t bytes = (size_t)160 * 32 * TILE_KB;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0x11, bytes);
    long long* dout;
    cudaMalloc(&dout, 148 * 8);
    run<4, 6>(buf, dout);
    run<7, 4>(buf, dout);
    return 0;
}
