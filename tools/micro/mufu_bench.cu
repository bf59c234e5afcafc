// Throughput of the special-function ops the LSTM gate math can be built from (per SM, per clock).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/micro/mufu_bench.cu -o /tmp/mufu_bench && /tmp/mufu_bench
#include <cstdio>
#include <cuda_fp16.h>
template <int OP>
__global__ void k(float* out, int iters, long long* cyc) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 3) { unsigned u = __float_as_uint(v[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
            if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
            if (OP == 5) { unsigned u = __float_as_uint(v[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name) {
    float* out; long long* cyc;
    const int threads = 1024, blocks = 148, iters = 2000;
    cudaMalloc(&out, sizeof(float) * threads * blocks);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    k<OP><<<blocks, threads>>>(out, iters, cyc);
    k<OP><<<blocks, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double ops = (double)threads * iters * 8;
    printf("%-24s %.2f thread-ops/clk/SM\n", name, ops / h[0]);
}
int main() {
    run<0>("tanh.approx.f32");
    run<1>("ex2.approx.ftz.f32");
    run<2>("rcp.approx.ftz.f32");
    run<3>("tanh.approx.f16x2 (instr)");
    run<5>("ex2.approx.ftz.f16x2 (instr)");
    run<4>("fma.rn.f32");
    return 0;
}
