// Microbenchmark: how fast can every SM stream a small, shared, L2-resident matrix through a TMA ring?
// (the weight stream of the persistent LSTM kernel).  Build on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I generative-audio_b200/csrc -I include \
//        tools/micro/tma_bench.cu generative-audio_b200/csrc/tc_host.cu generative-audio_b200/csrc/core.cu -o /tmp/tma_bench -ldl
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "tc_common.cuh"
using namespace nppc::tc;

constexpr int ROWS_W = 1536, K = 384;

// MODE 0: unicast, local barrier.  MODE 1: cta_group::2 load, barrier in the leader (count 2).  MODE 2: cluster-2 multicast
template <int MODE, int BOX_ROWS, int NST>
__global__ void __launch_bounds__(128, 1) ring_kernel(const __grid_constant__ CUtensorMap tm, int nloads, int ncopies,
                                                      long long* cycles) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int STAGE = BOX_ROWS * 64 * 2;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NST * STAGE);
    uint64_t* empty = full + NST;
    const uint32_t crank = MODE ? cluster_ctarank() : 0;
    const int copy = (blockIdx.x / (MODE ? 2 : 1)) % ncopies;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) {
            mbar_init(&full[i], MODE == 1 ? 2 : 1);
            mbar_init(&empty[i], MODE == 2 ? 2 : 1);
        }
        fence_barrier_init();
    }
    __syncthreads();
    if (MODE) cluster_sync_all();
    const int boxes_per_mat = (ROWS_W / (MODE == 1 ? 2 * BOX_ROWS : BOX_ROWS)) * (K / 64);
    long long t0 = clock64();
    if (threadIdx.x == 0) {   // producer
        int stage = 0; uint32_t phase = 0;
        const uint32_t f0 = MODE == 1 ? mapa_u32(smem_u32(full), 0) : 0;
        for (int i = 0; i < nloads; ++i) {
            const int b = (i + 7 * (blockIdx.x >> 1)) % boxes_per_mat;
            const int k = b % (K / 64), j = b / (K / 64);
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char* dst = smem + stage * STAGE;
            if (MODE == 0) {
                mbar_arrive_expect_tx(&full[stage], STAGE);
                tma_load_2d(dst, &tm, &full[stage], k * 64, copy * ROWS_W + j * BOX_ROWS);
            } else if (MODE == 1) {
                mbar_arrive_expect_tx_cluster(f0 + stage * 8, STAGE);
                tma_load_2d_pair(dst, &tm, f0 + stage * 8, k * 64, copy * ROWS_W + j * 2 * BOX_ROWS + crank * BOX_ROWS);
            } else {
                mbar_arrive_expect_tx(&full[stage], STAGE);
                if ((i & 1) == (int)crank) tma_load_2d_mcast(dst, &tm, &full[stage], k * 64, copy * ROWS_W + j * BOX_ROWS, 3);
            }
            if (++stage == NST) { stage = 0; phase ^= 1; }
        }
    } else if (threadIdx.x == 32) {   // consumer: frees the slot as soon as the data has landed
        int stage = 0; uint32_t phase = 0;
        if (MODE != 1 || crank == 0) {
            for (int i = 0; i < nloads; ++i) {
                mbar_wait(&full[stage], phase);
                if (MODE == 0) mbar_arrive(&empty[stage]);
                else if (MODE == 1) {
                    mbar_arrive_cluster(mapa_u32(smem_u32(&empty[stage]), 0));
                    mbar_arrive_cluster(mapa_u32(smem_u32(&empty[stage]), 1));
                } else {
                    mbar_arrive_cluster(mapa_u32(smem_u32(&empty[stage]), 0));
                    mbar_arrive_cluster(mapa_u32(smem_u32(&empty[stage]), 1));
                }
                if (++stage == NST) { stage = 0; phase ^= 1; }
            }
        }
    }
    __syncthreads();
    if (MODE) cluster_sync_all();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

template <int MODE, int BOX_ROWS, int NST>
void run(const char* name, const __half* w, int ncopies, int grid, int nloads) {
    CUtensorMap tm;
    if (make_tmap_bf16_2d(&tm, w, (uint64_t)ROWS_W * ncopies, K, K * 2, BOX_ROWS, 64)) { printf("tmap failed\n"); exit(1); }
    auto kern = ring_kernel<MODE, BOX_ROWS, NST>;
    const int smem = NST * BOX_ROWS * 128 + 1024 + 256;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* cyc;
    cudaMalloc(&cyc, sizeof(long long) * grid);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = MODE ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        cudaError_t err = cudaLaunchKernelEx(&cfg, kern, tm, nloads, ncopies, cyc);
        cudaEventRecord(e1);
        cudaError_t e2 = cudaDeviceSynchronize();
        if (err != cudaSuccess || e2 != cudaSuccess) { printf("%s: launch failed %s %s\n", name, cudaGetErrorString(err), cudaGetErrorString(e2)); exit(1); }
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    const double bytes_per_cta = (double)nloads * BOX_ROWS * 128 * (MODE == 2 ? 1 : 1);
    printf("%-44s grid %3d copies %2d: %8.3f ms  %6.1f B/clk/SM landed  %7.1f GB/s chip landed  (%lld cyc, %.0f cyc/load)\n", name, grid,
           ncopies, ms, bytes_per_cta / mx, bytes_per_cta * grid / ms / 1e6, mx, (double)mx / nloads);
    cudaFree(cyc);
}

int main() {
    const int NC = 32;
    __half* w;
    cudaMalloc(&w, sizeof(__half) * (size_t)ROWS_W * K * NC);
    cudaMemset(w, 0, sizeof(__half) * (size_t)ROWS_W * K * NC);
    const int N = 20000;
    for (int grid : {2, 130, 148}) {
        for (int nc : {1, NC}) {
            run<0, 128, 6>("unicast box128 ring6", w, nc, grid, N);
            run<0, 64, 12>("unicast box64 ring12", w, nc, grid, N);
            run<1, 64, 12>("pair cta_group::2 box64 ring12", w, nc, grid, N);
            run<1, 64, 4>("pair cta_group::2 box64 ring4", w, nc, grid, N);
            run<2, 128, 6>("cluster2 multicast box128 ring6", w, nc, grid, N);
        }
    }
    return 0;
}
