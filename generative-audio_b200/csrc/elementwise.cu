// HBM-bound elementwise / indexing kernels: cIRM (de)compression + mask apply (a9, a10), laplace norms (a2),
// sub-band unfold (a5), drop_band (a6), fused sub-band feature packing for the LSTM, output assembly (a8/a11).
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ float decompress1(float m) {
    // mask.py:57-60: limit*(m>=limit) - limit*(m<=-limit) + m*(|m|<limit); -K*log((K-m)/(K+m)), K=10, limit=9.9
    // = 2K atanh(c / K).  |c| <= 3: odd series 2K t (1 + t^2/3 + ... + t^10/11), t = c/K (truncation < 5e-8 relative; the
    // reference's own fp32 log of a ratio near 1 is only ~3e-5 relative there, and these small masks are what the PC head's
    // cancelling-mean normalisers amplify — speech fixtures).  |c| > 3: K ln2 (lg2(K + c) - lg2(K - c)) on two MUFU.LG2
    // (absolute error ~3e-6 on values >= 6.2); the libm logf + division this replaced made the kernel instruction-bound.
    float c = (m >= 9.9f) ? 9.9f : ((m <= -9.9f) ? -9.9f : m);
    // branch-free: both forms are a handful of instructions; the series wins for |c| <= 3 (t <= 0.3: truncation t^12/13 < 5e-8)
    const float t = c * 0.1f, u = t * t;
    float p = fmaf(u, 1.0f / 11.0f, 1.0f / 9.0f);
    p = fmaf(u, p, 1.0f / 7.0f);
    p = fmaf(u, p, 1.0f / 5.0f);
    p = fmaf(u, p, 1.0f / 3.0f);
    p = fmaf(u, p, 1.0f);
    const float ys = 20.0f * t * p;
    const float yl = 6.931471805599453f * (__log2f(10.0f + c) - __log2f(10.0f - c));
    return fabsf(c) <= 3.0f ? ys : yl;
}

__device__ __forceinline__ float compress1(float m) {
    // mask.py:44-50: m <- -100 where m <= -100 ; K*(1-exp(-C m))/(1+exp(-C m)), K=10, C=0.1
    m = (m <= -100.0f) ? -100.0f : m;
    float e = expf(-0.1f * m);
    return 10.0f * (1.0f - e) / (1.0f + e);
}

__global__ void __launch_bounds__(TPB) crm_apply_kernel(const float* __restrict__ crm, const float* __restrict__ re,
                                                       const float* __restrict__ im, int FT, int conj,
                                                       float* __restrict__ omag, float* __restrict__ ore,
                                                       float* __restrict__ oim) {
    const int b = blockIdx.y;
    const float* m0p = crm + (size_t)b * 2 * FT;
    const float* m1p = m0p + FT;
    const size_t off = (size_t)b * FT;
#pragma unroll 4
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < FT; i += gridDim.x * blockDim.x) {
        float m0 = decompress1(m0p[i]), m1 = decompress1(m1p[i]);
        float r = re[off + i], q = im[off + i];
        float er, ei;
        if (conj) {  // utils.py:241-249 through the swapped-argument call of :75-79  => conj(M) * N
            er = m0 * r + m1 * q;
            ei = m0 * q - m1 * r;
        } else {     // utils.py:54,75-79 => M * N
            er = m0 * r - m1 * q;
            ei = m1 * r + m0 * q;
        }
        ore[off + i] = er;
        oim[off + i] = ei;
        if (omag) omag[off + i] = sqrtf(er * er + ei * ei);
    }
}

// N1 (NPPCAudioValidator._crm_directions_to_spectograms + the alpha sweep of visualize_pc_spectrograms,
// nppc_audio/validator.py:55-102,246-290): for every direction d of w_mat, pc = decompress(w_d) * noisy (M*N, utils.py:252-256)
// and for every alpha a, var = enhanced + alpha_a * pc.  One read of w_mat / noisy / enhanced, all n*A variations written.
// grid (chunks, n, B)
__global__ void __launch_bounds__(TPB) pc_variations_kernel(const float* __restrict__ w_mat, const float* __restrict__ nre,
                                                           const float* __restrict__ nim, const float* __restrict__ ere,
                                                           const float* __restrict__ eim, int n, int FT,
                                                           const float* __restrict__ alphas, int A,
                                                           float* __restrict__ pc_re, float* __restrict__ pc_im,
                                                           float* __restrict__ var_re, float* __restrict__ var_im) {
    const int b = blockIdx.z, d = blockIdx.y;
    const float* m0p = w_mat + ((size_t)(b * n + d) * 2) * FT;
    const float* m1p = m0p + FT;
    const size_t off = (size_t)b * FT, poff = (size_t)(b * n + d) * FT;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < FT; i += gridDim.x * blockDim.x) {
        const float m0 = decompress1(m0p[i]), m1 = decompress1(m1p[i]);
        const float r = nre[off + i], q = nim[off + i];
        const float pr = m0 * r - m1 * q, pi = m1 * r + m0 * q;
        if (pc_re) { pc_re[poff + i] = pr; pc_im[poff + i] = pi; }
        const float er = ere[off + i], ei = eim[off + i];
        for (int a = 0; a < A; ++a) {
            const float al = alphas[a];
            const size_t vo = ((size_t)(b * n + d) * A + a) * FT + i;
            var_re[vo] = er + al * pr;
            var_im[vo] = ei + al * pi;
        }
    }
}

// save_audio_files' peak normalisation (validator.py:118-119,133-134,281-282): x[r] /= max|x[r]| + 1e-8, one CTA per row
__global__ void __launch_bounds__(TPB) peak_normalize_kernel(float* __restrict__ x, int L) {
    __shared__ float red[32];
    float* row = x + (size_t)blockIdx.x * L;
    float m = 0.f;
    for (int i = threadIdx.x; i < L; i += blockDim.x) m = fmaxf(m, fabsf(row[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    const float denom = m + 1e-8f;
    for (int i = threadIdx.x; i < L; i += blockDim.x) row[i] = row[i] / denom;
}

__global__ void __launch_bounds__(TPB) decompress_kernel(const float* __restrict__ m, long long n, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = decompress1(m[i]);
}

__global__ void __launch_bounds__(TPB) build_cirm_kernel(const float* __restrict__ nr, const float* __restrict__ ni,
                                                        const float* __restrict__ cr, const float* __restrict__ ci,
                                                        int FT, float* __restrict__ gt) {
    const int b = blockIdx.y;
    const size_t off = (size_t)b * FT;
    float* g0 = gt + (size_t)b * 2 * FT;
    float* g1 = g0 + FT;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < FT; i += gridDim.x * blockDim.x) {
        float a = nr[off + i], bb = ni[off + i], c = cr[off + i], d = ci[off + i];
        float den = a * a + bb * bb + 1.1920928955078125e-07f;  // EPSILON = finfo(float32).eps (constant.py:8)
        g0[i] = compress1((a * c + bb * d) / den);
        g1[i] = compress1((a * d - bb * c) / den);
    }
}

// ---- norms --------------------------------------------------------------------------------------------
// per-sample sum in fp64 (the reference's divisor is a cancelling sum for signed re/im planes)
__global__ void __launch_bounds__(TPB) sample_sum_kernel(const float* __restrict__ x, long long n, double* __restrict__ sums) {
    __shared__ double red[32];
    const int b = blockIdx.y;
    const float* xb = x + (size_t)b * n;
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {   // 4 independent loads in flight per thread
        float v0 = xb[i], v1 = xb[i + stride], v2 = xb[i + 2 * stride], v3 = xb[i + 3 * stride];
        acc += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
    }
    for (; i < n; i += stride) acc += (double)xb[i];
    acc = nppc::block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(&sums[b], acc);
}

// per-sample (sum x, sum |x|) in fp64 -> cancellation depth of the offline normaliser's divisor, see nppc_cancel_depth
__global__ void __launch_bounds__(TPB) sample_sum_abs_kernel(const float* __restrict__ x, long long n, double* __restrict__ sums) {
    __shared__ double red[32];
    const int b = blockIdx.y;
    const float* xb = x + (size_t)b * n;
    double acc = 0.0, aab = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = (double)xb[i];
        acc += v;
        aab += fabs(v);
    }
    acc = nppc::block_sum(acc, red);
    aab = nppc::block_sum(aab, red);
    if (threadIdx.x == 0) {
        atomicAdd(&sums[2 * b], acc);
        atomicAdd(&sums[2 * b + 1], aab);
    }
}
__global__ void cancel_depth_finish_kernel(const double* __restrict__ sums, int B, double count, float* __restrict__ depth) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double mean = sums[2 * b] / count, mabs = sums[2 * b + 1] / count;
    const float d = (float)(mabs / (fabs(mean) + 1e-5) / sqrt(count));
    depth[b] = fmaxf(depth[b], d);
}

__global__ void __launch_bounds__(TPB) offline_norm_kernel(const float* __restrict__ x, long long n,
                                                          const double* __restrict__ sums, double count,
                                                          float* __restrict__ y) {
    const int b = blockIdx.y;
    const float mu = (float)(sums[b] / count);
    const float inv = 1.0f / (mu + 1e-5f);   // x * (1/den): within 1 ulp of the reference's x / den
    const float* xb = x + (size_t)b * n;
    float* yb = y + (size_t)b * n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {   // 4 independent loads in flight per thread
        float v0 = xb[i], v1 = xb[i + stride], v2 = xb[i + 2 * stride], v3 = xb[i + 3 * stride];
        yb[i] = v0 * inv; yb[i + stride] = v1 * inv; yb[i + 2 * stride] = v2 * inv; yb[i + 3 * stride] = v3 * inv;
    }
    for (; i < n; i += stride) yb[i] = xb[i] * inv;
}

// x [B,F,T] -> y [B,F,T+la] = pad + norm (mean counts the zero look-ahead frames, fullsubnet_plus.py:158-165)
__global__ void __launch_bounds__(TPB) pad_norm_kernel(const float* __restrict__ x, int F, int T, int la,
                                                      const double* __restrict__ sums, float* __restrict__ y) {
    const int b = blockIdx.y;
    const int Tp = T + la;
    const float mu = (float)(sums[b] / ((double)F * Tp));
    const float inv = 1.0f / (mu + 1e-5f);
    const float* xb = x + (size_t)b * F * T;
    float* yb = y + (size_t)b * F * Tp;
    const int n = F * Tp;
#pragma unroll 4
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int f = i / Tp, t = i - f * Tp;
        yb[i] = (t < T) ? xb[f * T + t] * inv : 0.0f;
    }
}

// cumulative_laplace_norm: one CTA per (b*c) plane [F,T]: column sums -> block scan over T -> scale.
__global__ void __launch_bounds__(TPB) cumulative_norm_kernel(const float* __restrict__ x, int F, int T,
                                                             float* __restrict__ y) {
    extern __shared__ float sh[];  // [T] cumulative means, + 32 warp totals
    float* cmean = sh;
    float* wtot = sh + T;
    __shared__ float carry;
    const float* xb = x + (size_t)blockIdx.x * F * T;
    float* yb = y + (size_t)blockIdx.x * F * T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) carry = 0.f;
    __syncthreads();
    for (int tb = 0; tb < T; tb += blockDim.x) {
        int t = tb + threadIdx.x;
        float s = 0.f;
        if (t < T)
            for (int f = 0; f < F; ++f) s += xb[(size_t)f * T + t];  // coalesced across threads
        // inclusive warp scan (shuffle), then block-level carry
        float v = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            float u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        if (lane == 31) wtot[warp] = v;
        __syncthreads();
        float base = carry;
        for (int w = 0; w < warp; ++w) base += wtot[w];
        v += base;
        if (t < T) cmean[t] = v / ((float)F * (float)(t + 1)) + 1.1920928955078125e-07f;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = v;
        __syncthreads();
    }
    (void)nw;
    const int n = F * T;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int t = i % T;
        yb[i] = xb[i] / cmean[t];
    }
}

// ---- unfold / drop_band (bit-exact copies) -------------------------------------------------------------
__device__ __forceinline__ int reflect_idx(int p, int F) {
    if (p < 0) p = -p;
    if (p > F - 1) p = 2 * (F - 1) - p;
    return p;
}

// out [B,F,C,K,T]; one CTA per (b, f, c): thread = frame t, loop over the K neighbour rows (source and destination rows
// are both T-contiguous -> coalesced, no index division in the loop).
__global__ void __launch_bounds__(TPB) unfold_kernel(const float* __restrict__ x, int C, int F, int T, int nn,
                                                    float* __restrict__ out) {
    const int K = 2 * nn + 1;
    long long row = blockIdx.x;  // (b*F + f)*C + c
    int c = (int)(row % C);
    long long bf = row / C;
    int f = (int)(bf % F);
    int b = (int)(bf / F);
    const float* src = x + ((size_t)b * C + c) * F * T;
    float* dst = out + (size_t)row * K * T;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
#pragma unroll 8
        for (int k = 0; k < K; ++k) dst[(size_t)k * T + t] = src[(size_t)reflect_idx(f + k - nn, F) * T + t];
    }
}

// out [B,C,F/G,T]: out[ob, c, j, t] = x[g + G*(ob - start_g), c, g + G*j, t] where group g owns a run of batches
__global__ void __launch_bounds__(TPB) drop_band_kernel(const float* __restrict__ x, int B, int C, int F, int T,
                                                       int G, float* __restrict__ out) {
    const int Fg = F / G;
    long long row = blockIdx.x;  // (ob*C + c)*Fg + j
    int j = (int)(row % Fg);
    long long oc = row / Fg;
    int c = (int)(oc % C);
    int ob = (int)(oc / C);
    // group g has ceil((B-g)/G) samples, laid out group after group
    int g = 0, start = 0;
    for (; g < G; ++g) {
        int cnt = (B - g + G - 1) / G;
        if (ob < start + cnt) break;
        start += cnt;
    }
    int sb = g + G * (ob - start);
    const float* src = x + (((size_t)sb * C + c) * F + (g + G * j)) * T;
    float* dst = out + (size_t)row * T;
    for (int t = threadIdx.x; t < T; t += blockDim.x) dst[t] = src[t];
}

// ---- fused sub-band packing ---------------------------------------------------------------------------
__device__ __forceinline__ int count_f(int q, int F, int nn) {
    int lo = max(0, q - nn), hi = min(F - 1, q + nn);
    return hi >= lo ? hi - lo + 1 : 0;
}

// Per-sample sum of the (never materialised) [F,S,T'] sub-band input: sum_{f,k} unfold(nbr)[f,k,:] + fb + fbr + fbi.
// Row p of nbr_src appears cnt(p) = #{(f,k): reflect(f+k-N) == p} times.
__global__ void __launch_bounds__(TPB) subband_sum_kernel(const float* __restrict__ nbr, const float* __restrict__ fb,
                                                         const float* __restrict__ fbr, const float* __restrict__ fbi,
                                                         int F, int Tp, int nn, double* __restrict__ sums) {
    __shared__ double red[32];
    const int b = blockIdx.y;
    const size_t off = (size_t)b * F * Tp;
    const int n = F * Tp;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int p = i / Tp;
        // multiplicity of source row p: #f in [0,F) with |f-q| <= N, summed over the positions q that reflect onto p
        int cnt = count_f(p, F, nn);
        if (p >= 1) cnt += count_f(-p, F, nn);
        if (p <= F - 2) cnt += count_f(2 * (F - 1) - p, F, nn);
        acc += (double)cnt * (double)nbr[off + i] + (double)fb[off + i] + (double)fbr[off + i] + (double)fbi[off + i];
    }
    acc = nppc::block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(&sums[b], acc);
}

// xs[t][row][k], row = ob*Fg + j  (drop_band order), k < S real features, zero padded to KP = 64.
// One CTA per (8 rows, 32 frames): thread (r, tt) gathers its S features (reads coalesced along T'), the tile is
// transposed through shared memory and written as 16-byte pieces of KP-contiguous rows (8 consecutive rows of one
// frame = 8*KP contiguous elements).  All index arithmetic in the write phase is compile-time shifts.
constexpr int PK_ROWS = 8;
constexpr int PK_T = 32;
constexpr int PK_KP = 64;
// CUM: norm_type = cumulative_laplace_norm (base_model.py:227-257 applied to the [B, F, S, T'] sub-band tensor, i.e. per LSTM
// row: y[k, t] = x[k, t] / (cumsum_t(sum_k x[k, .]) / (S (t + 1)) + EPSILON)).  The CTA then owns its 8 rows for ALL frames
// (grid.y = 1) and walks the 32-frame tiles in order, carrying each row's running sum (fp64) in its warp.
template <bool F16OUT, bool CUM>
__global__ void __launch_bounds__(TPB) subband_pack_kernel(const float* __restrict__ nbr, const float* __restrict__ fb,
                                                          const float* __restrict__ fbr, const float* __restrict__ fbi,
                                                          int B, int F, int Tp, int nn, int G, int RS,
                                                          const double* __restrict__ sums, void* __restrict__ xs_out) {
    extern __shared__ float tile[];  // [PK_ROWS][PK_KP][PK_T + 1] (+1 word skew per row): both phases (nearly) conflict-free
    constexpr int TS = PK_T + 1;
    constexpr int RSK = PK_KP * TS + 1;
    const int S = 2 * nn + 1 + 3;
    const int Fg = F / G;
    const long long R = (long long)B * Fg;
    const int r = threadIdx.x >> 5, tt = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * PK_ROWS + r;
    float* mine = tile + (size_t)r * RSK + tt;   // element k at mine[k * TS]
    int sb = 0, f = 0;
    if (row < R) {
        int ob = (int)(row / Fg), j = (int)(row % Fg);
        sb = ob; f = j;
        if (G > 1) {  // drop_band: group g owns a run of output batches; sample = g + G*(ob - start), freq = g + G*j
            int g = 0, start = 0;
            for (; g < G; ++g) {
                int cnt = (B - g + G - 1) / G;
                if (ob < start + cnt) break;
                start += cnt;
            }
            sb = g + G * (ob - start);
            f = g + G * j;
        }
    }
    double carry = 0.0;   // CUM: running sum of the row up to the previous tile (same value in every lane of the warp)
    const int ntile = CUM ? (Tp + PK_T - 1) / PK_T : 1;
    for (int tb = 0; tb < ntile; ++tb) {
    const int tile_y = CUM ? tb : (int)blockIdx.y;
    const int t = tile_y * PK_T + tt;
    if (CUM && tb > 0) __syncthreads();   // the previous tile has been written out
    if (row < R && t < Tp) {
        const float* base = nbr + (size_t)sb * F * Tp + t;
        const size_t off = (size_t)sb * F * Tp + t;
        if (CUM) {
            double ssum = 0.0;
#pragma unroll 8
            for (int k = 0; k < 2 * nn + 1; ++k) { const float v = base[reflect_idx(f + k - nn, F) * Tp]; mine[k * TS] = v; ssum += (double)v; }
            const float v1 = fb[off + (size_t)f * Tp], v2 = fbr[off + (size_t)f * Tp], v3 = fbi[off + (size_t)f * Tp];
            mine[(2 * nn + 1) * TS] = v1; mine[(2 * nn + 2) * TS] = v2; mine[(2 * nn + 3) * TS] = v3;
            ssum += (double)v1 + (double)v2 + (double)v3;
            double cum = ssum;   // inclusive scan over the warp's 32 frames
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, cum, o); if (tt >= o) cum += u; }
            cum += carry;
            const float inv = 1.0f / ((float)(cum / ((double)S * (t + 1))) + 1.1920929e-07f);
            for (int k = 0; k < S; ++k) mine[k * TS] *= inv;
            for (int k = S; k < PK_KP; ++k) mine[k * TS] = 0.f;
            carry = __shfl_sync(0xffffffffu, cum, 31);
        } else {
            const float den = (float)(sums[sb] / ((double)F * S * Tp)) + 1e-5f;
            const float inv = 1.0f / den;   // x * (1/den): within 1 ulp of the reference's x / den, well inside the 1e-4 budget
#pragma unroll 8
            for (int k = 0; k < 2 * nn + 1; ++k) mine[k * TS] = base[reflect_idx(f + k - nn, F) * Tp] * inv;
            mine[(2 * nn + 1) * TS] = fb[off + (size_t)f * Tp] * inv;
            mine[(2 * nn + 2) * TS] = fbr[off + (size_t)f * Tp] * inv;
            mine[(2 * nn + 3) * TS] = fbi[off + (size_t)f * Tp] * inv;
            for (int k = S; k < PK_KP; ++k) mine[k * TS] = 0.f;
        }
    } else {
        for (int k = 0; k < PK_KP; ++k) mine[k * TS] = 0.f;
        if (CUM) {   // keep the warp's shuffles convergent (only rows >= R or the last, partial tile get here)
            double cum = 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, cum, o); if (tt >= o) cum += u; }
            carry = __shfl_sync(0xffffffffu, cum + carry, 31);
        }
    }
    __syncthreads();
    // write phase: piece = 8 consecutive k of one (frame, row); PK_T*PK_ROWS*8 pieces per CTA
    constexpr int PIECES = PK_T * PK_ROWS * (PK_KP / 8);
    for (int idx = threadIdx.x; idx < PIECES; idx += TPB) {
        const int kq = idx & 7, rr = (idx >> 3) & (PK_ROWS - 1), ft = idx >> 6;
        const long long orow = (long long)blockIdx.x * PK_ROWS + rr;
        const int ot = tile_y * PK_T + ft;
        if (orow >= RS || ot >= Tp) continue;
        const float* src = tile + (size_t)rr * RSK + (size_t)(kq * 8) * TS + ft;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = src[e * TS];
        const size_t o = ((size_t)ot * RS + orow) * PK_KP + kq * 8;
        if (F16OUT) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                __half2 h = __floats2half2_rn(fminf(fmaxf(v[2 * e], -65504.f), 65504.f), fminf(fmaxf(v[2 * e + 1], -65504.f), 65504.f));
                pk[e] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(xs_out) + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        } else {
            float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(xs_out) + o);
            d[0] = make_float4(v[0], v[1], v[2], v[3]);
            d[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    }   // tile loop
}

// y [B*Fp, O, Tp] -> out [B, O, Fp, Tp-la] (drop first `la` frames)
__global__ void __launch_bounds__(TPB) assemble_kernel(const float* __restrict__ y, int Fp, int O, int Tp, int la,
                                                      float* __restrict__ out) {
    long long row = blockIdx.x;  // (b*O + o)*Fp + f
    int f = (int)(row % Fp);
    long long bo = row / Fp;
    int o = (int)(bo % O);
    int b = (int)(bo / O);
    const float* src = y + (((size_t)b * Fp + f) * O + o) * Tp + la;
    float* dst = out + (size_t)row * (Tp - la);
    for (int t = threadIdx.x; t < Tp - la; t += blockDim.x) dst[t] = src[t];
}

int grid_for(long long n, int cap_mult = 8) {
    long long g = (n + TPB - 1) / TPB;
    long long cap = (long long)nppc::sm_count() * cap_mult;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int nppc_crm_decompress_apply(const float* crm, const float* real, const float* imag, int B, int FT,
                                         int conj, float* out_mag, float* out_real, float* out_imag, void* stream) {
    NPPC_CHECK_ARG(crm && real && imag && out_real && out_imag, "nppc_crm_decompress_apply: null pointer");
    NPPC_CHECK_ARG(B > 0 && FT > 0, "nppc_crm_decompress_apply: bad sizes B=%d FT=%d", B, FT);
    int gx = grid_for(FT, 8);
    int per = nppc::cdiv((long long)nppc::sm_count() * 8, B);
    if (gx > per && per >= 1) gx = per;
    crm_apply_kernel<<<dim3(gx, B), TPB, 0, (cudaStream_t)stream>>>(crm, real, imag, FT, conj, out_mag, out_real, out_imag);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_pc_variations(const float* w_mat, const float* noisy_real, const float* noisy_imag,
                                  const float* enh_real, const float* enh_imag, int B, int n, int FT, const float* alphas,
                                  int A, float* pc_real, float* pc_imag, float* var_real, float* var_imag, void* stream) {
    NPPC_CHECK_ARG(w_mat && noisy_real && noisy_imag && enh_real && enh_imag && alphas && var_real && var_imag,
                   "nppc_pc_variations: null pointer");
    NPPC_CHECK_ARG((pc_real == nullptr) == (pc_imag == nullptr), "nppc_pc_variations: pc_real / pc_imag must both be set or NULL");
    NPPC_CHECK_ARG(B > 0 && n > 0 && FT > 0 && A > 0 && n <= 65535 && B <= 65535, "nppc_pc_variations: bad sizes");
    int gx = grid_for(FT, 4);
    int per = nppc::cdiv((long long)nppc::sm_count() * 8, (long long)B * n);
    if (gx > per && per >= 1) gx = per;
    pc_variations_kernel<<<dim3(gx, n, B), TPB, 0, (cudaStream_t)stream>>>(w_mat, noisy_real, noisy_imag, enh_real, enh_imag, n,
                                                                         FT, alphas, A, pc_real, pc_imag, var_real, var_imag);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_peak_normalize(float* x, int rows, int L, void* stream) {
    NPPC_CHECK_ARG(x && rows > 0 && L > 0, "nppc_peak_normalize: bad arguments");
    peak_normalize_kernel<<<rows, TPB, 0, (cudaStream_t)stream>>>(x, L);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_decompress_cirm(const float* m, long long n, float* out, void* stream) {
    NPPC_CHECK_ARG(m && out && n > 0, "nppc_decompress_cirm: bad arguments");
    decompress_kernel<<<grid_for(n), TPB, 0, (cudaStream_t)stream>>>(m, n, out);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_build_cirm(const float* nr, const float* ni, const float* cr, const float* ci, int B, int FT,
                               float* gt, void* stream) {
    NPPC_CHECK_ARG(nr && ni && cr && ci && gt && B > 0 && FT > 0, "nppc_build_cirm: bad arguments");
    int gx = grid_for(FT, 8);
    build_cirm_kernel<<<dim3(gx, B), TPB, 0, (cudaStream_t)stream>>>(nr, ni, cr, ci, FT, gt);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

static int per_sample_grid(long long n, int B) {
    int gx = grid_for(n, 8);
    int per = nppc::cdiv((long long)nppc::sm_count() * 8, B);
    if (per < 1) per = 1;
    return gx > per ? per : gx;
}

extern "C" int nppc_offline_laplace_norm(const float* x, int B, long long n, double* sums, float* y, void* stream) {
    NPPC_CHECK_ARG(x && y && sums && B > 0 && n > 0, "nppc_offline_laplace_norm: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    NPPC_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * B, s));
    int gx = per_sample_grid(n, B);
    sample_sum_kernel<<<dim3(gx, B), TPB, 0, s>>>(x, n, sums);
    offline_norm_kernel<<<dim3(gx, B), TPB, 0, s>>>(x, n, sums, (double)n, y);
    NPPC_COUNT_LAUNCH(2);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

// depth[b] = max(depth[b], (mean|x_b| / (|mean x_b| + 1e-5)) / sqrt(count)): how much deeper the divisor of
// offline_laplace_norm (base_model.py:219-222, a signed mean for the real / imag streams) cancels than a random-sign sum of
// `count` terms would (depth ~ 1).  Errors of relative size eps in x come out of the normaliser amplified by ~ eps * depth.
extern "C" int nppc_cancel_depth(const float* x, int B, long long n, double count, double* sums, float* depth, void* stream) {
    NPPC_CHECK_ARG(x && sums && depth && B > 0 && n > 0 && count > 0, "nppc_cancel_depth: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    NPPC_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * B, s));
    sample_sum_abs_kernel<<<dim3(per_sample_grid(n, B), B), TPB, 0, s>>>(x, n, sums);
    cancel_depth_finish_kernel<<<nppc::cdiv(B, 128), 128, 0, s>>>(sums, B, count, depth);
    NPPC_COUNT_LAUNCH(2);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_pad_offline_laplace_norm(const float* x, int B, int F, int T, int look_ahead, double* sums,
                                             float* y, void* stream) {
    NPPC_CHECK_ARG(x && y && sums && B > 0 && F > 0 && T > 0 && look_ahead >= 0, "nppc_pad_offline_laplace_norm: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    NPPC_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * B, s));
    int gx = per_sample_grid((long long)F * T, B);
    sample_sum_kernel<<<dim3(gx, B), TPB, 0, s>>>(x, (long long)F * T, sums);
    pad_norm_kernel<<<dim3(gx, B), TPB, 0, s>>>(x, F, T, look_ahead, sums, y);
    NPPC_COUNT_LAUNCH(2);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_cumulative_laplace_norm(const float* x, int BC, int F, int T, float* y, void* stream) {
    NPPC_CHECK_ARG(x && y && BC > 0 && F > 0 && T > 0, "nppc_cumulative_laplace_norm: bad arguments");
    size_t smem = sizeof(float) * (T + 32);
    NPPC_CHECK_ARG(smem <= 200 * 1024, "nppc_cumulative_laplace_norm: T=%d too large for the single-CTA scan", T);
    if (smem > 48 * 1024)
        NPPC_CUDA_OK(cudaFuncSetAttribute(cumulative_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cumulative_norm_kernel<<<BC, TPB, smem, (cudaStream_t)stream>>>(x, F, T, y);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_unfold(const float* x, int B, int C, int F, int T, int num_neighbor, float* out, void* stream) {
    NPPC_CHECK_ARG(x && out && B > 0 && C > 0 && F > 0 && T > 0 && num_neighbor >= 0, "nppc_unfold: bad arguments");
    NPPC_CHECK_ARG(num_neighbor < F, "nppc_unfold: reflect padding needs num_neighbor < F (%d >= %d)", num_neighbor, F);
    unfold_kernel<<<(unsigned)((long long)B * F * C), TPB, 0, (cudaStream_t)stream>>>(x, C, F, T, num_neighbor, out);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_drop_band(const float* x, int B, int C, int F, int T, int groups, float* out, void* stream) {
    NPPC_CHECK_ARG(x && out && C > 0 && F > 0 && T > 0, "nppc_drop_band: bad arguments");
    // feature.py:263 asserts before the num_groups<=1 early-out
    NPPC_CHECK_ARG(B > groups, "Batch size = %d, num_groups = %d. The batch size should larger than the num_groups.", B, groups);
    int G = groups <= 1 ? 1 : groups;
    int Fg = F / G;
    drop_band_kernel<<<(unsigned)((long long)B * C * Fg), 128, 0, (cudaStream_t)stream>>>(x, B, C, F, T, G, out);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_subband_pack(const float* nbr_src, const float* fb, const float* fbr, const float* fbi, int B,
                                 int F, int Tp, int num_neighbor, int groups, int KP, int R_stride, int cumulative, double* sums,
                                 float* xs_f32, void* xs_f16, void* stream) {
    NPPC_CHECK_ARG(nbr_src && fb && fbr && fbi && sums, "nppc_subband_pack: null pointer");
    NPPC_CHECK_ARG(xs_f32 || xs_f16, "nppc_subband_pack: no output requested");
    int S = 2 * num_neighbor + 4;
    NPPC_CHECK_ARG(B > 0 && F > 0 && Tp > 0 && num_neighbor >= 1 && num_neighbor < F && KP >= S,
                   "nppc_subband_pack: bad sizes (B=%d F=%d Tp=%d N=%d KP=%d)", B, F, Tp, num_neighbor, KP);
    int G = groups <= 1 ? 1 : groups;
    if (B > 1) NPPC_CHECK_ARG(B > groups, "Batch size = %d, num_groups = %d. The batch size should larger than the num_groups.", B, groups);
    else G = 1;  // fullsubnet_plus.py:213: drop_band only runs when batch_size > 1
    cudaStream_t s = (cudaStream_t)stream;
    NPPC_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * B, s));
    int gx = per_sample_grid((long long)F * Tp, B);
    subband_sum_kernel<<<dim3(gx, B), TPB, 0, s>>>(nbr_src, fb, fbr, fbi, F, Tp, num_neighbor, sums);
    long long R = (long long)B * (F / G);
    NPPC_CHECK_ARG(R_stride >= R, "nppc_subband_pack: R_stride (%d) < rows (%lld)", R_stride, R);
    NPPC_CHECK_ARG(KP == PK_KP, "nppc_subband_pack: KP must be %d (got %d)", PK_KP, KP);
    dim3 grid((unsigned)nppc::cdiv(R_stride, PK_ROWS), (unsigned)nppc::cdiv(Tp, PK_T));
    int nlaunch = 1;
    const size_t smem = sizeof(float) * PK_ROWS * (PK_KP * (PK_T + 1) + 1);
    NPPC_CUDA_OK(cudaFuncSetAttribute(subband_pack_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NPPC_CUDA_OK(cudaFuncSetAttribute(subband_pack_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NPPC_CUDA_OK(cudaFuncSetAttribute(subband_pack_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NPPC_CUDA_OK(cudaFuncSetAttribute(subband_pack_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (cumulative) grid.y = 1;   // the CTA walks all frames of its 8 rows (running mean along T')
    if (xs_f32) {
        if (cumulative) subband_pack_kernel<false, true><<<grid, TPB, smem, s>>>(nbr_src, fb, fbr, fbi, B, F, Tp, num_neighbor, G, R_stride, sums, xs_f32);
        else subband_pack_kernel<false, false><<<grid, TPB, smem, s>>>(nbr_src, fb, fbr, fbi, B, F, Tp, num_neighbor, G, R_stride, sums, xs_f32);
        ++nlaunch;
    }
    if (xs_f16) {
        if (cumulative) subband_pack_kernel<true, true><<<grid, TPB, smem, s>>>(nbr_src, fb, fbr, fbi, B, F, Tp, num_neighbor, G, R_stride, sums, xs_f16);
        else subband_pack_kernel<true, false><<<grid, TPB, smem, s>>>(nbr_src, fb, fbr, fbi, B, F, Tp, num_neighbor, G, R_stride, sums, xs_f16);
        ++nlaunch;
    }
    NPPC_COUNT_LAUNCH(nlaunch);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_assemble_mask(const float* y, int B, int Fp, int O, int Tp, int look_ahead, float* out, void* stream) {
    NPPC_CHECK_ARG(y && out && B > 0 && Fp > 0 && O > 0 && Tp > look_ahead && look_ahead >= 0, "nppc_assemble_mask: bad arguments");
    assemble_kernel<<<(unsigned)((long long)B * O * Fp), 128, 0, (cudaStream_t)stream>>>(y, Fp, O, Tp, look_ahead, out);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}
