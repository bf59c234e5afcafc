// a3 TSSE channel attention and the element-wise / normalisation half of the a4 TCN blocks (the 1x1 convolutions
// stay on the library GEMM for now).  All HBM-bound; layout is channel-first [B, C, T'] fp32
// as the 1x1 convolutions produce it.
//
// TCNBlock (causal_conv.py:96-108):  y1 = conv1x1(x);  h = GroupNorm1(PReLU1(y1));  z = PReLU2(depthwise(h));
//                                    x' = x + sconv(GroupNorm2(z))
// GroupNorm(1, C) needs per-SAMPLE statistics over all C*T' elements, so the block is cut at the two reductions:
//   prelu_stats : one read of y1            -> stats1[b] = (sum, sum of squares) of PReLU1(y1 + b1)     (fp64 atomics)
//   tcn_mid     : one read of y1, one write -> z = PReLU2(dw_b + sum_j k_j * norm1(PReLU1(y1))[t+(j-1)d]), stats2[b] of z
//   (library)   : o = conv1x1(z; W2 * diag(gamma2))            (GroupNorm2's affine folded into the weights)
//   tcn_out     : x' = x + o * rstd2[b] + (W2 beta2 + b2)[c] - mean2[b] * rstd2[b] * (W2 gamma2)[c]
#include "common.cuh"

namespace {
constexpr int TPB = 256;

__device__ __forceinline__ float prelu(float v, float a) { return v >= 0.f ? v : a * v; }

// grid (blocks per sample, B): each warp owns whole channel rows (coalesced along T'); y + bias[c] is the conv1x1 output
// (the bias of the 1x1 convolution is folded in here and in tcn_mid, so the library GEMM runs without a bias pass).
__global__ void __launch_bounds__(TPB) prelu_stats_kernel(const float* __restrict__ y, int C, int T, const float* __restrict__ bias,
                                                         const float* __restrict__ a_ptr, double* __restrict__ stats) {
    __shared__ double red[32];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float a = *a_ptr;
    double ds = 0.0, dss = 0.0;
    for (int c = blockIdx.x * nw + warp; c < C; c += gridDim.x * nw) {
        const float bc = bias ? bias[c] : 0.f;
        const float* row = y + ((size_t)b * C + c) * T;
        float s = 0.f, ss = 0.f;
        for (int t = lane; t < T; t += 32) {
            const float p = prelu(row[t] + bc, a);
            s += p;
            ss += p * p;
        }
        ds += (double)s;
        dss += (double)ss;
    }
    ds = nppc::block_sum(ds, red);
    dss = nppc::block_sum(dss, red);
    if (threadIdx.x == 0) {
        atomicAdd(&stats[2 * b], ds);
        atomicAdd(&stats[2 * b + 1], dss);
    }
}

// grid (blocks per sample, B), 8 warps per block: each warp owns whole channel rows (coalesced along T', the 3 taps hit
// L1), moments are accumulated per thread in fp32 over one row, per block in fp64, ONE atomic pair per block.
__global__ void __launch_bounds__(TPB) tcn_mid_kernel(const float* __restrict__ y1, int C, int T, const float* __restrict__ bias1,
                                                     const float* __restrict__ a1_ptr,
                                                     const double* __restrict__ stats1, const float* __restrict__ g1,
                                                     const float* __restrict__ b1, const float* __restrict__ dw_w,
                                                     const float* __restrict__ dw_b, int dil, const float* __restrict__ a2_ptr,
                                                     float* __restrict__ z, double* __restrict__ stats2) {
    __shared__ double red[32];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const double n = (double)C * T;
    const double mu_d = stats1[2 * b] / n;
    const double var_d = stats1[2 * b + 1] / n - mu_d * mu_d;
    const float mu = (float)mu_d, rstd = (float)(1.0 / sqrt(fmax(var_d, 0.0) + 1e-8));
    const float a1 = *a1_ptr, a2 = *a2_ptr;
    double ds = 0.0, dss = 0.0;
    for (int c = blockIdx.x * nw + warp; c < C; c += gridDim.x * nw) {
        const float sc = rstd * g1[c], sh = b1[c] - mu * rstd * g1[c];   // norm1(v) = v * sc + sh
        const float k0 = dw_w[c * 3], k1 = dw_w[c * 3 + 1], k2 = dw_w[c * 3 + 2], kb = dw_b[c];
        const float bc = bias1 ? bias1[c] : 0.f;
        const float* row = y1 + ((size_t)b * C + c) * T;
        float* zrow = z + ((size_t)b * C + c) * T;
        float s = 0.f, ss = 0.f;
        for (int t = lane; t < T; t += 32) {
            float acc = kb;
            int tm = t - dil, tp = t + dil;
            if (tm >= 0) acc += k0 * (prelu(row[tm] + bc, a1) * sc + sh);   // zero padding applies to the NORMALISED signal
            acc += k1 * (prelu(row[t] + bc, a1) * sc + sh);
            if (tp < T) acc += k2 * (prelu(row[tp] + bc, a1) * sc + sh);
            float out = prelu(acc, a2);
            zrow[t] = out;
            s += out;
            ss += out * out;
        }
        ds += (double)s;
        dss += (double)ss;
    }
    ds = nppc::block_sum(ds, red);
    dss = nppc::block_sum(dss, red);
    if (threadIdx.x == 0) {
        atomicAdd(&stats2[2 * b], ds);
        atomicAdd(&stats2[2 * b + 1], dss);
    }
}

__global__ void __launch_bounds__(TPB) tcn_out_kernel(const float* __restrict__ o, const float* __restrict__ x, int C, int T,
                                                     int Ch, const double* __restrict__ stats2, const float* __restrict__ u,
                                                     const float* __restrict__ vb, float* __restrict__ xnew) {
    const int b = blockIdx.y;
    const double n = (double)Ch * T;
    const double mu_d = stats2[2 * b] / n;
    const double var_d = stats2[2 * b + 1] / n - mu_d * mu_d;
    const float rstd = (float)(1.0 / sqrt(fmax(var_d, 0.0) + 1e-8));
    const float mr = (float)mu_d * rstd;
    const size_t off = (size_t)b * C * T;
    const int total = C * T;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int c = i / T;
        xnew[off + i] = x[off + i] + o[off + i] * rstd + (vb[c] - mr * u[c]);
    }
}

// ---- TSSE (attention_model.py:78-98) -----------------------------------------------------------------------------
// squeeze: s[b,c] = a0 + sum_k a_k * relu(mean_t(depthwise_valid_conv_k(x)[c]))  with
// mean_t(conv_k(x))[c] = bias + sum_j w[c][j] * (S - prefix(j) - suffix(k-1-j)) / (T-k+1): one warp per (b, c) row.
constexpr int KMAX = 16;
__global__ void __launch_bounds__(TPB) tsse_squeeze_kernel(const float* __restrict__ x, int C, int T, int k0, int k1, int k2,
                                                          const float* __restrict__ w0, const float* __restrict__ bb0,
                                                          const float* __restrict__ w1, const float* __restrict__ bb1,
                                                          const float* __restrict__ w2, const float* __restrict__ bb2,
                                                          const float* __restrict__ fcw, const float* __restrict__ fcb,
                                                          float* __restrict__ s_out, int rows) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int c = warp % C;
    const float* row = x + (size_t)warp * T;
    float tot = 0.f;
    for (int t = lane; t < T; t += 32) tot += row[t];
    tot = nppc::warp_sum(tot);
    // prefix(j) = sum of the first j samples, suffix(m) = sum of the last m samples (j, m < KMAX)
    float head = lane < KMAX && lane < T ? row[lane] : 0.f;
    float tail = lane < KMAX && lane < T ? row[T - 1 - lane] : 0.f;
    float pre[KMAX], suf[KMAX];
    float ph = 0.f, pt = 0.f;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
        pre[j] = ph; suf[j] = pt;
        ph += __shfl_sync(0xffffffffu, head, j);
        pt += __shfl_sync(0xffffffffu, tail, j);
    }
    const int ks[3] = {k0, k1, k2};
    const float* ws[3] = {w0, w1, w2};
    const float* bs[3] = {bb0, bb1, bb2};
    float s = fcb[0];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int k = ks[q];
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < k) acc += ws[q][c * k + j] * (tot - pre[j] - suf[k - 1 - j]);
        float f = bs[q][c] + acc / (float)(T - k + 1);
        s += fcw[q] * fmaxf(f, 0.f);
    }
    if (lane == 0) s_out[warp] = s;
}

// excitation: g = sigmoid(W2 relu(W1 s + b1) + b2) per sample (one CTA per sample), then y = x * g[c].
// one CTA per sample (the two mat-vecs are dependent); 32 warps so the 128 + C latency-bound warp dot products finish in
// a quarter of the time 8 warps took (51 us -> ~15 us per call, 9 calls per step)
__global__ void __launch_bounds__(1024) tsse_excite_kernel(const float* __restrict__ s_in, int C, int Cr,
                                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                                         float* __restrict__ g_out) {
    extern __shared__ float sh[];  // s [C], hidden [Cr]
    float* ss = sh;
    float* hh = sh + C;
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < C; i += blockDim.x) ss[i] = s_in[(size_t)b * C + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < Cr; r += nw) {
        float a = 0.f;
        for (int i = lane; i < C; i += 32) a += w1[(size_t)r * C + i] * ss[i];
        a = nppc::warp_sum(a);
        if (lane == 0) hh[r] = fmaxf(a + b1[r], 0.f);
    }
    __syncthreads();
    for (int c = warp; c < C; c += nw) {
        float a = 0.f;
        for (int i = lane; i < Cr; i += 32) a += w2[(size_t)c * Cr + i] * hh[i];
        a = nppc::warp_sum(a);
        if (lane == 0) g_out[(size_t)b * C + c] = 1.f / (1.f + expf(-(a + b2[c])));
    }
}

__global__ void __launch_bounds__(TPB) tsse_scale_kernel(const float* __restrict__ x, const float* __restrict__ g, int T,
                                                        long long total, float* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        y[i] = x[i] * g[i / T];
}

int blocks_for(long long n, int mult) {
    long long g = (n + TPB - 1) / TPB, cap = (long long)nppc::sm_count() * mult;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
}  // namespace

extern "C" int nppc_prelu_stats(const float* y, int B, int C, int T, const float* bias, const float* prelu_a, double* stats,
                                void* stream) {
    NPPC_CHECK_ARG(y && prelu_a && stats && B > 0 && C > 0 && T > 0 && B <= 65535, "nppc_prelu_stats: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    NPPC_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B, s));
    int per = nppc::cdiv((long long)nppc::sm_count() * 8, B);
    int gx = nppc::cdiv(C, TPB / 32);
    if (gx > per) gx = per < 1 ? 1 : per;
    prelu_stats_kernel<<<dim3(gx, B), TPB, 0, s>>>(y, C, T, bias, prelu_a, stats);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_tcn_mid(const float* y1, int B, int C, int T, const float* bias1, const float* prelu1_a, const double* stats1,
                            const float* gamma1, const float* beta1, const float* dw_w, const float* dw_b, int dilation,
                            const float* prelu2_a, float* z, double* stats2, void* stream) {
    NPPC_CHECK_ARG(y1 && prelu1_a && stats1 && gamma1 && beta1 && dw_w && dw_b && prelu2_a && z && stats2, "nppc_tcn_mid: null pointer");
    NPPC_CHECK_ARG(B > 0 && C > 0 && T > 0 && dilation > 0 && C <= 65535 && B <= 65535, "nppc_tcn_mid: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    NPPC_CUDA_OK(cudaMemsetAsync(stats2, 0, sizeof(double) * 2 * B, s));
    {
        int per = nppc::cdiv((long long)nppc::sm_count() * 8, B);
        int gx = nppc::cdiv(C, TPB / 32);
        if (gx > per) gx = per < 1 ? 1 : per;
        tcn_mid_kernel<<<dim3(gx, B), TPB, 0, s>>>(y1, C, T, bias1, prelu1_a, stats1, gamma1, beta1, dw_w, dw_b, dilation,
                                                   prelu2_a, z, stats2);
    }
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_tcn_out(const float* o, const float* x, int B, int C, int T, int C_hidden, const double* stats2,
                            const float* u, const float* vb, float* xnew, void* stream) {
    NPPC_CHECK_ARG(o && x && stats2 && u && vb && xnew && B > 0 && C > 0 && T > 0 && C_hidden > 0, "nppc_tcn_out: bad arguments");
    int per = nppc::cdiv((long long)nppc::sm_count() * 8, B);
    int gx = blocks_for((long long)C * T, 8);
    if (gx > per) gx = per < 1 ? 1 : per;
    tcn_out_kernel<<<dim3(gx, B), TPB, 0, (cudaStream_t)stream>>>(o, x, C, T, C_hidden, stats2, u, vb, xnew);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_tsse(const float* x, int B, int C, int T, const int* kersize, const float* const* conv_w,
                         const float* const* conv_b, const float* fcat_w, const float* fcat_b, const float* fc1_w,
                         const float* fc1_b, const float* fc2_w, const float* fc2_b, int C_reduced, float* scratch,
                         float* y, void* stream) {
    NPPC_CHECK_ARG(x && kersize && conv_w && conv_b && fcat_w && fcat_b && fc1_w && fc1_b && fc2_w && fc2_b && scratch && y,
                   "nppc_tsse: null pointer");
    NPPC_CHECK_ARG(B > 0 && C > 0 && C_reduced > 0, "nppc_tsse: bad sizes");
    for (int q = 0; q < 3; ++q)
        NPPC_CHECK_ARG(kersize[q] >= 1 && kersize[q] <= KMAX && kersize[q] <= T, "nppc_tsse: kernel size %d unsupported (1..%d, <= T)",
                       kersize[q], KMAX);
    cudaStream_t s = (cudaStream_t)stream;
    float* sq = scratch;                 // [B, C] squeeze
    float* g = scratch + (size_t)B * C;  // [B, C] gates
    int rows = B * C;
    tsse_squeeze_kernel<<<nppc::cdiv((long long)rows * 32, TPB), TPB, 0, s>>>(x, C, T, kersize[0], kersize[1], kersize[2],
                                                                            conv_w[0], conv_b[0], conv_w[1], conv_b[1], conv_w[2],
                                                                            conv_b[2], fcat_w, fcat_b, sq, rows);
    tsse_excite_kernel<<<B, 1024, sizeof(float) * (C + C_reduced), s>>>(sq, C, C_reduced, fc1_w, fc1_b, fc2_w, fc2_b, g);
    long long total = (long long)rows * T;
    tsse_scale_kernel<<<blocks_for(total, 8), TPB, 0, s>>>(x, g, T, total, y);
    NPPC_COUNT_LAUNCH(3);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}
