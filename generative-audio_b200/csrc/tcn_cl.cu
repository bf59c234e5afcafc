// Row N2: the full-band TCN stack with its 1x1 convolutions on the tcgen05 GEMM (gemm_tc.cu) instead of the library.
// A 1x1 Conv1d over [B, C, T'] is a TN GEMM once the activations are CHANNEL-LAST:  Y[b*T'+t, co] = sum_ci X[b*T'+t, ci] W[co, ci]
// (both operands K-major), so inside the stack everything lives as [M = B*T' rows][channels]:
//   x32 [M][Kp]  fp32  residual stream (exact accumulation across the 8 blocks), rows padded like xh (16-byte vectors)
//   xh  [M][Kp]  fp16  the same values as GEMM A operand, C zero-padded to Kp (multiple of 64)
//   y1  [M][512] fp16  conv1x1 output WITHOUT bias (bias is applied where y1 is read)
//   z   [M][512] fp16  PReLU2(depthwise(GroupNorm1(PReLU1(y1 + b1))))                       (causal_conv.py:100-104)
//   o   [M][Np]  fp16  z * (W2 diag(gamma2))^T, Np = C rounded up to 128                      (GroupNorm2 folded, :105-106)
//   x32 <- x32 + o * rstd2[b] + vb[c] - mean2[b] * rstd2[b] * u[c]                            (:107-108)
// fp16 RANGE: the real / imag streams are normalised by a cancelling mean (base_model.py:219-222) and can reach 1e4 and more,
// so xh holds x / scale[b] with scale[b] = max|x_0| of the sample (<= 1 on entry, huge head-room for the residual growth);
// the consumers of y1 (and of the final Linear) multiply the per-sample scale back in fp32: y1 = scale[b] * (W1 xh) + b1.
// fp16 MANTISSA: the frozen backbone's output feeds the PC head's cancelling-mean normalisers (mean(enhanced_real) etc.), which
// amplify any backbone error by the cancellation depth (measured 200x on speech, tests/golden/speech12.npz).  Error model
// (tools/tcn_fp16_error_model.py): of all fp16 rounding points only three matter — the residual stream as GEMM A operand, the
// conv1x1 / fc weights, and the fc output.  With `split` the A operand is written as hi + lo fp16 halves ([M][2 Kp]: hi | lo)
// and the weights as [Whi | Whi | Wlo] ([N][3 Kp]): the GEMM computes hi*Whi + lo*Whi + hi*Wlo (~22 mantissa bits, fp32
// accumulate) by wrapping its A K-blocks (gemm_tc.cu a_wrap), and the fc GEMM writes fp32.
// All kernels here are HBM-bound streaming passes over fp16 tensors (half the bytes of the channel-first fp32 versions
// in tcn.cu); per-sample GroupNorm moments are fp64 block reductions + fp64 atomics as before.
#include <cuda_fp16.h>
#include "common.cuh"

namespace {
constexpr int TPB = 256;
constexpr int HID = 512;   // TCNBlock hidden width (causal_conv.py:67 default; fb_model_hidden_size is ignored upstream)

__device__ __forceinline__ float prelu(float v, float a) { return v >= 0.f ? v : a * v; }

// GroupNorm(1, 512) moments of a sample from its fp64 (sum, sum of squares): the cancellation-prone part (E[x^2] - mean^2)
// stays in fp64, the division is a multiplication by the host-computed 1 / (512 T) and the reciprocal square root is the
// fp32 MUFU one (2 ulp) -- a dozen instructions per thread instead of the ~150 of fp64 division + sqrt.
__device__ __forceinline__ void moments(const double* __restrict__ st, double inv_n, float& mu, float& rstd) {
    const double mu_d = st[0] * inv_n;
    const double var_d = fma(st[1], inv_n, -mu_d * mu_d);
    mu = (float)mu_d;
    rstd = rsqrtf(fmaxf((float)var_d, 0.f) + 1e-8f);
}

// channel-first fp32 [B][C][T] -> x32 [B*T][Kp] fp32 and xh [B*T][Kp] fp16 (columns [C, Kp) of xh are written as zeros).
// 32x32 shared-memory transpose tiles; grid (ceil(T/32), ceil(C/32), B), block (32, 8).
// hi / lo split of a scaled value: hi = fp16(v), lo = fp16(v - hi) (exact subtraction in fp32)
__device__ __forceinline__ void split_h(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

__global__ void pack_cl_kernel(const float* __restrict__ x, int C, int T, int Kp, const float* __restrict__ inv_scale,
                               float* __restrict__ x32, __half* __restrict__ xh, int split) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && t < T) ? x[((size_t)b * C + c) * T + t] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        if (t < T && c < Kp) {   // columns [C, Kp) of xh are the GEMM's K padding: written as zeros here (no separate fill)
            const float v = tile[threadIdx.x][i];
            const size_t row = (size_t)b * T + t;
            if (c < C) x32[row * Kp + c] = v;
            if (split) {
                __half hi, lo;
                split_h(v * inv_scale[b], hi, lo);
                xh[row * (2 * Kp) + c] = hi;
                xh[row * (2 * Kp) + Kp + c] = lo;
            } else {
                xh[row * Kp + c] = __float2half_rn(v * inv_scale[b]);
            }
        }
    }
}

// per-sample fp16 range scale: scale[b] = max(max |x[b]|, 1e-30), inv_scale[b] = 1 / scale[b].  grid (chunks, B): every CTA folds
// its slice into the sample's maximum with one atomicMax on the float's bit pattern (non-negative floats order like unsigned
// integers; the slot is zeroed first), a B-thread kernel then clamps and inverts.
__global__ void __launch_bounds__(TPB) absmax_cl_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ bits) {
    __shared__ float red[TPB / 32];
    const float* xb = x + (size_t)blockIdx.y * n;
    const long long per = (n + gridDim.x - 1) / gridDim.x, i0 = blockIdx.x * per, i1 = min(n, i0 + per);
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
    long long i = i0 + threadIdx.x;
    for (; i + 3 * TPB < i1; i += 4 * TPB) {   // four independent loads in flight
        m0 = fmaxf(m0, fabsf(xb[i]));
        m1 = fmaxf(m1, fabsf(xb[i + TPB]));
        m2 = fmaxf(m2, fabsf(xb[i + 2 * TPB]));
        m3 = fmaxf(m3, fabsf(xb[i + 3 * TPB]));
    }
    for (; i < i1; i += TPB) m0 = fmaxf(m0, fabsf(xb[i]));
    float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < TPB / 32; ++w) m = fmaxf(m, red[w]);
        atomicMax(bits + blockIdx.y, __float_as_uint(m));
    }
}
__global__ void scale_finish_kernel(int B, float* __restrict__ scale, float* __restrict__ inv_scale) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        // floor 1: the residual stream grows by O(1) per block whatever the input level (GroupNorm'd branch), so a quiet or
        // silent sample (max|x| -> 0) must not get a tiny scale — x32 / scale would overflow fp16 after the first block
        // (measured on the digitally silent utterance of tests/golden/speech12.npz: pred_crm 0.76 off before this floor).
        float sc = fmaxf(scale[b], 1.0f);   // (fmaxf drops NaNs: a NaN input still poisons the output through the GEMMs)
        scale[b] = sc;
        inv_scale[b] = 1.0f / sc;
    }
}

// o [B*T][Np] fp16 (+ bias[c], optional ReLU) -> channel-first fp32 [B][C][T]
template <typename TO>
__global__ void unpack_cl_kernel(const TO* __restrict__ o, int C, int T, int Np, const float* __restrict__ scale,
                                 const float* __restrict__ bias, int relu, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        float v = 0.f;
        if (t < T && c < C) {
            v = (float)o[((size_t)b * T + t) * Np + c] * (scale ? scale[b] : 1.f) + (bias ? bias[c] : 0.f);
            if (relu) v = fmaxf(v, 0.f);
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        if (c < C && t < T) out[((size_t)b * C + c) * T + t] = tile[threadIdx.x][i];
    }
}

// stats[b] = (sum, sum of squares) of PReLU(y1 + bias) over the sample's [T][512]; thread = channel pair, grid (chunks, B)
__global__ void __launch_bounds__(TPB) prelu_stats_cl_kernel(const __half2* __restrict__ y1, int T, const float* __restrict__ scale,
                                                            const float* __restrict__ bias, const float* __restrict__ a_ptr,
                                                            double* __restrict__ stats) {
    __shared__ double red[32];
    const int b = blockIdx.y, c2 = threadIdx.x;
    const float sb = scale[b];
    const float a = *a_ptr, b0 = bias[2 * c2], b1 = bias[2 * c2 + 1];
    const __half2* base = y1 + (size_t)b * T * (HID / 2) + c2;
    float s = 0.f, ss = 0.f;
    // each CTA owns a contiguous run of rows; 4 independent row loads in flight per thread
    const int rows = (T + gridDim.x - 1) / gridDim.x, tb = blockIdx.x * rows, te = min(T, tb + rows);
    for (int t = tb; t < te; t += 4) {
        __half2 raw[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (t + q < te) raw[q] = base[(size_t)(t + q) * (HID / 2)];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (t + q < te) {
                const float2 v = __half22float2(raw[q]);
                const float p0 = prelu(fmaf(v.x, sb, b0), a), p1 = prelu(fmaf(v.y, sb, b1), a);
                s += p0 + p1;
                ss += p0 * p0 + p1 * p1;
            }
    }
    double ds = nppc::block_sum((double)s, red);
    double dss = nppc::block_sum((double)ss, red);
    if (threadIdx.x == 0) {
        atomicAdd(&stats[2 * b], ds);
        atomicAdd(&stats[2 * b + 1], dss);
    }
}

// z = PReLU2(dw_b + k0 n(t-d) + k1 n(t) + k2 n(t+d)), n(t) = GroupNorm1(PReLU1(y1[t] + b1)) with ZERO padding of n outside
// [0,T); stats2[b] = moments of z.  thread = channel pair (per-channel constants in registers), grid (chunks of rows, B).
__global__ void __launch_bounds__(TPB) tcn_mid_cl_kernel(const __half2* __restrict__ y1, int T, const float* __restrict__ scale,
                                                        const float* __restrict__ bias1,
                                                        const float* __restrict__ a1_ptr, const double* __restrict__ stats1,
                                                        const float* __restrict__ g1, const float* __restrict__ be1,
                                                        const float* __restrict__ dw_w, const float* __restrict__ dw_b, int dil,
                                                        const float* __restrict__ a2_ptr, __half2* __restrict__ z,
                                                        double* __restrict__ stats2, double inv_n) {
    __shared__ double red[32];
    const int b = blockIdx.y, c2 = threadIdx.x;
    float mu, rstd;
    moments(stats1 + 2 * b, inv_n, mu, rstd);
    const float a1 = *a1_ptr, a2 = *a2_ptr, sb = scale[b];
    float sc[2], sh[2], k0[2], k1[2], k2[2], kb[2], bb[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int c = 2 * c2 + e;
        sc[e] = rstd * g1[c];
        sh[e] = be1[c] - mu * rstd * g1[c];
        k0[e] = dw_w[c * 3]; k1[e] = dw_w[c * 3 + 1]; k2[e] = dw_w[c * 3 + 2]; kb[e] = dw_b[c];
        bb[e] = bias1[c];
    }
    const __half2* base = y1 + (size_t)b * T * (HID / 2) + c2;
    __half2* zb = z + (size_t)b * T * (HID / 2) + c2;
    auto nrm = [&](int t, int e, float2 v) { return prelu(fmaf(e ? v.y : v.x, sb, bb[e]), a1) * sc[e] + sh[e]; };
    float s = 0.f, ss = 0.f;
    const int rows = (T + gridDim.x - 1) / gridDim.x, tb = blockIdx.x * rows, te = min(T, tb + rows);
#pragma unroll 2
    for (int t = tb; t < te; ++t) {
        const int tm = t - dil, tp = t + dil;
        const float2 vc = __half22float2(base[(size_t)t * (HID / 2)]);
        float2 vm = make_float2(0.f, 0.f), vp = make_float2(0.f, 0.f);
        if (tm >= 0) vm = __half22float2(base[(size_t)tm * (HID / 2)]);
        if (tp < T) vp = __half22float2(base[(size_t)tp * (HID / 2)]);
        float outv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float acc = kb[e] + k1[e] * nrm(t, e, vc);
            if (tm >= 0) acc += k0[e] * nrm(tm, e, vm);
            if (tp < T) acc += k2[e] * nrm(tp, e, vp);
            outv[e] = prelu(acc, a2);
            s += outv[e];
            ss += outv[e] * outv[e];
        }
        zb[(size_t)t * (HID / 2)] = __floats2half2_rn(outv[0], outv[1]);
    }
    double ds = nppc::block_sum((double)s, red);
    double dss = nppc::block_sum((double)ss, red);
    if (threadIdx.x == 0) {
        atomicAdd(&stats2[2 * b], ds);
        atomicAdd(&stats2[2 * b + 1], dss);
    }
}

// x32 <- x32 + o * rstd + (vb - mean * rstd * u); xh <- fp16(x32 / scale) or fp16(relu(x32) / scale) (last block: the stack's
// trailing ReLU).  thread = (row lane, 4 consecutive channels): x32 rows are padded to Kp floats, so every access is a 16-byte
// (fp32) or 8-byte (fp16) vector; the per-channel constants of the quad stay in registers while the thread walks down its rows.
// grid (row chunks, B).  Channels >= C of the last quad: xh stays zero (GEMM K padding), x32 padding is never read as data.
constexpr int OUT_TPB = 288;
__global__ void __launch_bounds__(OUT_TPB) tcn_out_cl_kernel(const __half* __restrict__ o, float* __restrict__ x32, int T, int C, int Np,
                                                            int Kp, const double* __restrict__ stats2, const float* __restrict__ u,
                                                            const float* __restrict__ vb, const float* __restrict__ inv_scale,
                                                            __half* __restrict__ xh, int relu_h, double inv_n, int split) {
    const int b = blockIdx.y;
    const int CQ = (C + 3) >> 2;                 // channel quads per row
    const int RL = OUT_TPB / CQ;                 // rows worked on side by side (host checks RL >= 1)
    const int cq = threadIdx.x % CQ, rl = threadIdx.x / CQ;
    if (rl >= RL) return;
    const float is = inv_scale[b];
    float mu, rstd;
    moments(stats2 + 2 * b, inv_n, mu, rstd);
    const float mr = mu * rstd;
    const float lo = relu_h ? 0.f : -3.0e38f;    // trailing ReLU of the stack (last block) as a lower clamp
    const int rows = (T + gridDim.x - 1) / gridDim.x, tb = blockIdx.x * rows, te = min(T, tb + rows);
    const int c0 = cq * 4;
    float kc[4];
    bool ok[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        ok[e] = c0 + e < C;
        kc[e] = ok[e] ? vb[c0 + e] - mr * u[c0 + e] : 0.f;
    }
    const size_t r0 = (size_t)b * T + tb + rl;
    float4* xp = reinterpret_cast<float4*>(x32 + r0 * Kp + c0);
    const uint2* op = reinterpret_cast<const uint2*>(o + r0 * Np + c0);
    const int XS = split ? 2 * Kp : Kp;             // xh row stride (hi | lo halves when split)
    uint2* hp = reinterpret_cast<uint2*>(xh + r0 * XS + c0);
    const int sx = RL * Kp / 4, so = RL * Np / 4;   // strides between this thread's consecutive rows, in vectors
    const int sh = RL * XS / 4;
    auto finish = [&](float4 xv, uint2 ov, float4* xd, uint2* hd) {
        const float2 o01 = __half22float2(*reinterpret_cast<const __half2*>(&ov.x));
        const float2 o23 = __half22float2(*reinterpret_cast<const __half2*>(&ov.y));
        float v[4] = {fmaf(o01.x, rstd, xv.x) + kc[0], fmaf(o01.y, rstd, xv.y) + kc[1], fmaf(o23.x, rstd, xv.z) + kc[2],
                      fmaf(o23.y, rstd, xv.w) + kc[3]};
        float h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] = ok[e] ? fminf(fmaxf(fmaxf(v[e], lo) * is, -65504.f), 65504.f) : 0.f;
        *xd = make_float4(v[0], v[1], v[2], v[3]);
        const __half2 h01 = __floats2half2_rn(h[0], h[1]), h23 = __floats2half2_rn(h[2], h[3]);
        *hd = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        if (split) {   // lo half: the fp16 rounding residual of the scaled value (zero in the K padding)
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(h[0] - f01.x, h[1] - f01.y), l23 = __floats2half2_rn(h[2] - f23.x, h[3] - f23.y);
            hd[Kp / 4] = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
        }
    };
    int nr = (te - tb - rl + RL - 1) / RL;   // rows of this chunk that fall on this row lane
    for (; nr >= 2; nr -= 2) {               // two independent rows in flight
        const float4 xa = xp[0], xb = xp[sx];
        const uint2 oa = __ldg(op), ob = __ldg(op + so);
        finish(xa, oa, xp, hp);
        finish(xb, ob, xp + sx, hp + sh);
        xp += 2 * sx; op += 2 * so; hp += 2 * sh;
    }
    if (nr > 0) finish(xp[0], __ldg(op), xp, hp);
}
}  // namespace

extern "C" int nppc_tcn_cl_scale(const float* x, int B, long long n_per_sample, float* scale, float* inv_scale, void* stream) {
    NPPC_CHECK_ARG(x && scale && inv_scale && B > 0 && n_per_sample > 0, "nppc_tcn_cl_scale: bad arguments");
    NPPC_CHECK_ARG(B <= 65535, "nppc_tcn_cl_scale: B too large");
    cudaStream_t s = (cudaStream_t)stream;
    NPPC_CUDA_OK(cudaMemsetAsync(scale, 0, sizeof(float) * B, s));
    int chunks = nppc::cdiv((long long)nppc::sm_count() * 4, B);
    const int cap = (int)((n_per_sample + 4 * TPB - 1) / (4 * TPB));   // at least one full 4-load round per CTA
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    absmax_cl_kernel<<<dim3(chunks, B), TPB, 0, s>>>(x, n_per_sample, reinterpret_cast<unsigned int*>(scale));
    scale_finish_kernel<<<nppc::cdiv(B, 128), 128, 0, s>>>(B, scale, inv_scale);
    NPPC_COUNT_LAUNCH(2);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_tcn_cl_pack(const float* x, int B, int C, int T, int Kp, const float* inv_scale, float* x32, void* xh,
                                int split, void* stream) {
    NPPC_CHECK_ARG(x && inv_scale && x32 && xh && B > 0 && C > 0 && T > 0 && Kp >= C && B <= 65535, "nppc_tcn_cl_pack: bad arguments");
    pack_cl_kernel<<<dim3(nppc::cdiv(T, 32), nppc::cdiv(Kp, 32), B), dim3(32, 8), 0, (cudaStream_t)stream>>>(x, C, T, Kp, inv_scale, x32, (__half*)xh, split);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_tcn_cl_unpack(const void* o, int o_f32, int B, int C, int T, int Np, const float* scale, const float* bias,
                                  int relu, float* out, void* stream) {
    NPPC_CHECK_ARG(o && out && B > 0 && C > 0 && T > 0 && Np >= C && B <= 65535, "nppc_tcn_cl_unpack: bad arguments");
    const dim3 grid(nppc::cdiv(T, 32), nppc::cdiv(C, 32), B), block(32, 8);
    if (o_f32) unpack_cl_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)o, C, T, Np, scale, bias, relu, out);
    else unpack_cl_kernel<__half><<<grid, block, 0, (cudaStream_t)stream>>>((const __half*)o, C, T, Np, scale, bias, relu, out);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

// CTAs per sample: contiguous row chunks, ~4 waves of 8 CTAs/SM over the whole batch, at least 8 rows per CTA
static int rows_grid(int T, int B, int min_rows = 8) {
    int per = nppc::cdiv((long long)nppc::sm_count() * 32, B);
    int cap = nppc::cdiv(T, min_rows);
    if (per > cap) per = cap;
    return per < 1 ? 1 : per;
}

extern "C" int nppc_prelu_stats_cl(const void* y1, int B, int T, int H, const float* scale, const float* bias, const float* prelu_a,
                                   double* stats, void* stream) {
    NPPC_CHECK_ARG(y1 && scale && bias && prelu_a && stats && B > 0 && T > 0 && B <= 65535, "nppc_prelu_stats_cl: bad arguments");
    NPPC_CHECK_ARG(H == HID, "nppc_prelu_stats_cl: hidden width must be %d (got %d)", HID, H);
    cudaStream_t s = (cudaStream_t)stream;
    NPPC_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B, s));
    prelu_stats_cl_kernel<<<dim3(rows_grid(T, B), B), TPB, 0, s>>>((const __half2*)y1, T, scale, bias, prelu_a, stats);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_tcn_mid_cl(const void* y1, int B, int T, int H, const float* scale, const float* bias1, const float* prelu1_a,
                               const double* stats1,
                               const float* gamma1, const float* beta1, const float* dw_w, const float* dw_b, int dilation,
                               const float* prelu2_a, void* z, double* stats2, void* stream) {
    NPPC_CHECK_ARG(y1 && scale && bias1 && prelu1_a && stats1 && gamma1 && beta1 && dw_w && dw_b && prelu2_a && z && stats2,
                   "nppc_tcn_mid_cl: null pointer");
    NPPC_CHECK_ARG(B > 0 && T > 0 && dilation > 0 && B <= 65535 && H == HID, "nppc_tcn_mid_cl: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    NPPC_CUDA_OK(cudaMemsetAsync(stats2, 0, sizeof(double) * 2 * B, s));
    // measured and rejected: 16 rows per CTA; computing n(t) once per row into a shared-memory window (dil 9 gets slower,
    // net -6 %); 4 channels per thread; per-warp REDs instead of the block reduction.  The CTA's fixed costs dominate.
    tcn_mid_cl_kernel<<<dim3(rows_grid(T, B), B), TPB, 0, s>>>((const __half2*)y1, T, scale, bias1, prelu1_a, stats1, gamma1, beta1, dw_w,
                                                              dw_b, dilation, prelu2_a, (__half2*)z, stats2, 1.0 / ((double)HID * T));
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_tcn_out_cl(const void* o, float* x32, int B, int T, int C, int Np, int Kp, int H, const double* stats2,
                               const float* u, const float* vb, const float* inv_scale, void* xh, int relu_h, int split,
                               void* stream) {
    NPPC_CHECK_ARG(o && x32 && stats2 && u && vb && inv_scale && xh && B > 0 && T > 0 && C > 0 && Np >= C && Kp >= C && B <= 65535 && H == HID,
                   "nppc_tcn_out_cl: bad arguments");
    NPPC_CHECK_ARG(Kp % 4 == 0 && Np % 4 == 0 && (C + 3) / 4 <= OUT_TPB, "nppc_tcn_out_cl: Kp, Np must be multiples of 4 and C <= %d", 4 * OUT_TPB);
    tcn_out_cl_kernel<<<dim3(rows_grid(T, B), B), OUT_TPB, 0, (cudaStream_t)stream>>>((const __half*)o, x32, T, C, Np, Kp, stats2, u, vb, inv_scale, (__half*)xh, relu_h,
                                                                                          1.0 / ((double)HID * T), split);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}
