// Shared helpers for the nppc_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nppc_b200.h"

namespace nppc {

void set_error(const char* fmt, ...);

#define NPPC_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            nppc::set_error(__VA_ARGS__);         \
            return NPPC_ERR_INVALID_ARGUMENT;     \
        }                                         \
    } while (0)

#define NPPC_CUDA_OK(expr)                                                                  \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            nppc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return NPPC_ERR_CUDA;                                                           \
        }                                                                                   \
    } while (0)

#define NPPC_LAUNCH_OK() NPPC_CUDA_OK(cudaGetLastError())

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of a double; result valid in thread 0. `red` needs >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    return v;
}

int sm_count();

}  // namespace nppc

#include <atomic>
namespace nppc {
extern std::atomic<long long> g_launches;
}
#define NPPC_COUNT_LAUNCH(n) nppc::g_launches.fetch_add((n), std::memory_order_relaxed)
