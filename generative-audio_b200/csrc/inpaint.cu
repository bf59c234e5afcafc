// a16 (inpainting variant): the HBM-bound glue around the UNets.
//   logmag stats / apply : utils.preprocess_data + preprocess_log_magnitude (utils.py:281-306): |S| -> log(|S| + 1e-6),
//                          ONE global mean / unbiased std over the whole clean batch tensor (fp64 sums), then both the
//                          clean and the masked spectrogram are normalised with those scalars.
//   mask_blend           : RestorationWrapper.forward (networks/unet.py:298-313): x_in[:,0]*m + net(x_in)*(1-m), and the
//                          PC wrapper's  alternatives * (1 - m)  (inpainting/nppc/pc_wrapper.py:75-81) when x_in == NULL.
// Coalesced, vectorised where the row length allows; grids sized in multiples of the SM count.
#include "common.cuh"

namespace {
constexpr int TPB = 256;

// spec [B,2,P] (re, im planes) -> sums[0] += sum log(|S|+eps), sums[1] += sum log^2
__global__ void __launch_bounds__(TPB) logmag_stats_kernel(const float* __restrict__ spec, int B, long long P, float eps,
                                                          double* __restrict__ sums) {
    __shared__ double red[32];
    const long long n = (long long)B * P;
    float s = 0.f, ss = 0.f;
    double ds = 0.0, dss = 0.0;
    int cnt = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / P, p = i - b * P;
        const float re = spec[(size_t)(2 * b) * P + p], im = spec[(size_t)(2 * b + 1) * P + p];
        const float l = logf(sqrtf(re * re + im * im) + eps);
        s += l;
        ss += l * l;
        if (++cnt == 16) { ds += s; dss += ss; s = ss = 0.f; cnt = 0; }
    }
    ds += s; dss += ss;
    ds = nppc::block_sum(ds, red);
    dss = nppc::block_sum(dss, red);
    if (threadIdx.x == 0) {
        atomicAdd(&sums[0], ds);
        atomicAdd(&sums[1], dss);
    }
}

// out[b,p] = (log(|spec[b]| + eps) - mean) / std with mean / std from sums over n_stat elements (unbiased std)
__global__ void __launch_bounds__(TPB) logmag_apply_kernel(const float* __restrict__ spec, int B, long long P, float eps,
                                                          const double* __restrict__ sums, double n_stat,
                                                          float* __restrict__ out) {
    const double mean_d = sums[0] / n_stat;
    const double var_d = (sums[1] - n_stat * mean_d * mean_d) / (n_stat - 1.0);
    const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(fmax(var_d, 0.0)));
    const long long n = (long long)B * P;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / P, p = i - b * P;
        const float re = spec[(size_t)(2 * b) * P + p], im = spec[(size_t)(2 * b + 1) * P + p];
        out[i] = (logf(sqrtf(re * re + im * im) + eps) - mean) * rstd;
    }
}

// out[b,c,p] = (x_in ? x_in[b,0,p] * m : 0) + x[b,c,p] * (1 - m),  m = mask[b,p]
__global__ void __launch_bounds__(TPB) mask_blend_kernel(const float* __restrict__ x_in, int Cin, const float* __restrict__ x,
                                                        const float* __restrict__ mask, int C, long long P,
                                                        float* __restrict__ out) {
    const int b = blockIdx.y;
    const float* mb = mask + (size_t)b * P;
    const float* xi = x_in ? x_in + (size_t)b * Cin * P : nullptr;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const float m = mb[p];
        const float keep = xi ? xi[p] * m : 0.f;
        for (int c = 0; c < C; ++c) {
            const size_t o = ((size_t)b * C + c) * P + p;
            out[o] = keep + x[o] * (1.f - m);
        }
    }
}
}  // namespace

extern "C" int nppc_logmag_stats(const float* spec, int B, long long P, double* sums, void* stream) {
    NPPC_CHECK_ARG(spec && sums && B > 0 && P > 0, "nppc_logmag_stats: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    NPPC_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * 2, s));
    long long blocks = ((long long)B * P + TPB * 8 - 1) / (TPB * 8);
    const long long cap = 8LL * nppc::sm_count();
    if (blocks > cap) blocks = cap;
    logmag_stats_kernel<<<(unsigned)(blocks < 1 ? 1 : blocks), TPB, 0, s>>>(spec, B, P, 1e-6f, sums);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_logmag_apply(const float* spec, int B, long long P, const double* sums, long long n_stat, float* out,
                                 void* stream) {
    NPPC_CHECK_ARG(spec && sums && out && B > 0 && P > 0 && n_stat > 1, "nppc_logmag_apply: bad arguments");
    long long blocks = ((long long)B * P + TPB * 4 - 1) / (TPB * 4);
    const long long cap = 8LL * nppc::sm_count();
    if (blocks > cap) blocks = cap;
    logmag_apply_kernel<<<(unsigned)(blocks < 1 ? 1 : blocks), TPB, 0, (cudaStream_t)stream>>>(spec, B, P, 1e-6f, sums,
                                                                                               (double)n_stat, out);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_mask_blend(const float* x_in, int Cin, const float* x, const float* mask, int B, int C, long long P,
                               float* out, void* stream) {
    NPPC_CHECK_ARG(x && mask && out && B > 0 && C > 0 && P > 0 && B <= 65535 && (!x_in || Cin > 0), "nppc_mask_blend: bad arguments");
    long long blocks = (P + TPB - 1) / TPB;
    long long want = (8LL * nppc::sm_count() + B - 1) / B;
    if (blocks > want) blocks = want;
    mask_blend_kernel<<<dim3((unsigned)(blocks < 1 ? 1 : blocks), B), TPB, 0, (cudaStream_t)stream>>>(x_in, Cin, x, mask, C, P, out);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}
