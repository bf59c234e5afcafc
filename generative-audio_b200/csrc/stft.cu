// a1 STFT front end and a15 iSTFT back end (n_fft = win = 512, hop = 256).
//
// HBM-bound kernels: one CTA transforms 16 consecutive frames of one utterance so that the [B,F,T]
// (T contiguous) planes are read/written in 64-byte runs.  The 512-point real FFT of a frame is a 256-point complex
// FFT (even/odd packing) done by a HALF warp as 16 x 16: two register-resident 16-point FFTs per lane (radix 4 x 4)
// with one conflict-free (17-padded) shared-memory transpose in between — two frames per warp at a time, 3 warp
// syncs per transform instead of the 8 shared-memory radix-2 stages this replaced.  Twiddles and the periodic hann
// window are staged in shared memory once per CTA.
#include <mutex>
#include <math.h>
#include "common.cuh"

namespace {

constexpr int NFFT = 512;
constexpr int HOP = 256;
constexpr int NBIN = 257;
constexpr int M = 256;        // complex FFT size
constexpr int FRAMES = 16;    // frames per CTA (64-byte output runs; 4 CTAs / SM)
constexpr int WARPS = 8;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// 4-point DFT in place (forward: W4 = -i, inverse: +i); outputs k = 0..3 land in x0..x3
template <bool INVERSE>
__device__ __forceinline__ void fft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 s02 = cadd(x0, x2), d02 = csub(x0, x2), s13 = cadd(x1, x3), d13 = csub(x1, x3);
    const float2 jd = INVERSE ? make_float2(-d13.y, d13.x) : make_float2(d13.y, -d13.x);   // (-/+ i) * d13
    x0 = cadd(s02, s13);
    x2 = csub(s02, s13);
    x1 = cadd(d02, jd);
    x3 = csub(d02, jd);
}

// 16-point DFT in registers (radix 4 x 4).  In: a[n] natural order.  Out: A[k] is found at a[4*(k & 3) + (k >> 2)].
template <bool INVERSE>
__device__ __forceinline__ void fft16(float2 (&a)[16]) {
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) fft4<INVERSE>(a[n2], a[4 + n2], a[8 + n2], a[12 + n2]);   // -> b[k1][n2] at a[4*k1 + n2]
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R = 0.70710678118654752f;
    // W16^m = (cos, -/+ sin)(2 pi m / 16) for m = n2 * k1
    const float sg = INVERSE ? 1.f : -1.f;
    const float2 w1 = make_float2(C1, sg * S1), w2 = make_float2(R, sg * R), w3 = make_float2(S1, sg * C1);
    const float2 w4 = make_float2(0.f, sg), w6 = make_float2(-R, sg * R), w9 = make_float2(-C1, -sg * S1);
    a[4 + 1] = cmul(a[4 + 1], w1); a[4 + 2] = cmul(a[4 + 2], w2); a[4 + 3] = cmul(a[4 + 3], w3);
    a[8 + 1] = cmul(a[8 + 1], w2); a[8 + 2] = cmul(a[8 + 2], w4); a[8 + 3] = cmul(a[8 + 3], w6);
    a[12 + 1] = cmul(a[12 + 1], w3); a[12 + 2] = cmul(a[12 + 2], w6); a[12 + 3] = cmul(a[12 + 3], w9);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) fft4<INVERSE>(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);   // -> A[k1 + 4*k2] at a[4*k1 + k2]
}

constexpr int FBUF = 16 * 17;   // float2 per frame buffer: 256 natural-order points, or the 17-padded 16 x 16 transpose

// In-place 256-point complex FFT of `buf` (natural order in, natural order out), forward (e^{-i..}) or inverse (e^{+i..}),
// unnormalised, executed by ONE HALF WARP (hl = lane & 15); both halves of the warp must call it together (it contains
// __syncwarp), each on its own buffer.  tw[k] = exp(-2*pi*i*k/512), k < 256.
//   Z[k1 + 16 k2] = sum_m2 W16^(m2 k2) [ W256^(m2 k1) sum_m1 z[16 m1 + m2] W16^(m1 k1) ]
template <bool INVERSE>
__device__ __forceinline__ void half_warp_fft256(float2* buf, const float2* tw, int hl) {
    float2 a[16];
#pragma unroll
    for (int m1 = 0; m1 < 16; ++m1) a[m1] = buf[16 * m1 + hl];       // z[16 m1 + m2], m2 = hl
    __syncwarp();
    fft16<INVERSE>(a);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        float2 v = a[4 * (k1 & 3) + (k1 >> 2)];
        const int m = hl * k1;                                        // W256^m, m <= 225
        float2 w = tw[(2 * m) & 255];
        if (m >= 128) { w.x = -w.x; w.y = -w.y; }
        if (INVERSE) w.y = -w.y;
        buf[k1 * 17 + hl] = cmul(v, w);                               // Y[k1][m2]
    }
    __syncwarp();
#pragma unroll
    for (int m2 = 0; m2 < 16; ++m2) a[m2] = buf[hl * 17 + m2];       // lane = k1
    __syncwarp();
    fft16<INVERSE>(a);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) buf[hl + 16 * k2] = a[4 * (k2 & 3) + (k2 >> 2)];
    __syncwarp();
}

struct __align__(16) SmemTables {
    float2 tw[M];     // exp(-2 pi i k / 512)
    float win[NFFT];  // periodic hann
};

__device__ SmemTables g_tables;   // per device; filled once per device by ensure_tables (host-computed, blocking upload)
__device__ __forceinline__ void fill_tables(SmemTables* tb) {
    const float4* src = reinterpret_cast<const float4*>(&g_tables);
    float4* dst = reinterpret_cast<float4*>(tb);
    for (int i = threadIdx.x; i < (int)(sizeof(SmemTables) / 16); i += blockDim.x) dst[i] = src[i];
}
// Twiddles and window are computed on the host in double precision and uploaded with a BLOCKING cudaMemcpyToSymbol the first
// time each device is used (per-device flag under a mutex): the copy has completed before this returns, so kernels launched
// afterwards on ANY stream of that device, from any thread, see the tables.
int ensure_tables(cudaStream_t) {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev]) return 0;
    static SmemTables host;
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < M; ++k) host.tw[k] = make_float2((float)cos(-pi * k / 256.0), (float)sin(-pi * k / 256.0));
    for (int n = 0; n < NFFT; ++n) host.win[n] = (float)(0.5 - 0.5 * cos(pi * n / 256.0));
    if (cudaMemcpyToSymbol(g_tables, &host, sizeof(SmemTables)) != cudaSuccess) return -1;
    done[dev] = true;
    return 0;
}

// grid: (ceil(T/32), B), block 256. dynamic smem: tables + per-warp FFT buffers + output tile.
constexpr int SFB = FBUF + 1;   // per-frame buffer stride in the STFT kernel: odd, so the output stage (lanes = frames) is conflict-free

__global__ void __launch_bounds__(WARPS * 32) stft_kernel(const float* __restrict__ wave, int L, int T,
                                                         float* __restrict__ mag, float* __restrict__ re,
                                                         float* __restrict__ im) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemTables* tb = reinterpret_cast<SmemTables*>(smem_raw);
    float2* fftbuf = reinterpret_cast<float2*>(smem_raw + sizeof(SmemTables));       // [FRAMES][SFB]: Z, then X[0..256] in place
    float* xs = reinterpret_cast<float*>(fftbuf + FRAMES * SFB);                     // [(FRAMES + 1) * HOP] signal span
    const int b = blockIdx.y, t0 = blockIdx.x * FRAMES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hw = lane >> 4, hl = lane & 15;
    fill_tables(tb);
    const float* x = wave + (size_t)b * L;
    // the CTA's span of the reflect-padded signal, (FRAMES + 1) hops, loaded ONCE (coalesced); every sample feeds two frames
    {
        const int p0 = t0 * HOP - NFFT / 2;
        for (int i = threadIdx.x; i < (FRAMES + 1) * HOP; i += blockDim.x) {
            int p = p0 + i;
            if (p < 0) p = -p;
            if (p > L - 1) p = 2 * (L - 1) - p;
            xs[i] = (p >= 0 && p < L) ? x[p] : 0.f;   // frames beyond T never read their samples
        }
    }
    __syncthreads();
    static_assert(FRAMES == 2 * WARPS, "one pass: every half warp owns exactly one frame");
    {
        const int fi = warp * 2 + hw, t = t0 + fi;
        float2* buf = fftbuf + fi * SFB;
        const bool ok = t < T;   // per half warp; both halves run the transform (it syncs the warp)
        // 512 windowed samples packed even/odd into 256 complex points, natural order
        const float2* xf = reinterpret_cast<const float2*>(xs + fi * HOP);
        const float2* wf = reinterpret_cast<const float2*>(tb->win);
#pragma unroll 4
        for (int q = 0; q < 16; ++q) {
            const int m = hl + 16 * q;
            const float2 xv = xf[m], wv = wf[m];
            buf[m] = ok ? make_float2(xv.x * wv.x, xv.y * wv.y) : make_float2(0.f, 0.f);
        }
        __syncwarp();
        half_warp_fft256<false>(buf, tb->tw, hl);
        // real-FFT recovery, in place and two bins at a time.  With E = (Z[k] + conj Z[M-k])/2, O = (Z[k] - conj Z[M-k])/(2i),
        // T = W512^k O:   X[k] = E + T,   X[M-k] = conj(E - T)   (E[M-k] = conj E[k], O[M-k] = conj O[k], W^(M-k) = -conj W^k)
        for (int k = hl; k <= M / 2; k += 16) {
            const float2 zk = buf[k];
            if (k == 0) {
                buf[0] = make_float2(zk.x + zk.y, 0.f);
                buf[M] = make_float2(zk.x - zk.y, 0.f);
            } else if (k == M / 2) {
                buf[k] = make_float2(zk.x, -zk.y);   // W^(M/2) = -i
            } else {
                const float2 zm = buf[M - k];
                const float2 E = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                const float2 D = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));  // (Zk - conj Zm)/2
                const float2 Tw = cmul(tb->tw[k], make_float2(D.y, -D.x));                  // W^k * (D / i)
                buf[k] = make_float2(E.x + Tw.x, E.y + Tw.y);
                buf[M - k] = make_float2(E.x - Tw.x, Tw.y - E.y);
            }
        }
    }
    __syncthreads();
    // output: consecutive threads = consecutive frames (64-byte runs along T in each of the three planes)
    const size_t plane = (size_t)b * NBIN * T;
    for (int idx = threadIdx.x; idx < NBIN * FRAMES; idx += blockDim.x) {
        const int k = idx / FRAMES, fi = idx % FRAMES;
        const int t = t0 + fi;
        if (t < T) {
            const float2 v = fftbuf[fi * SFB + k];
            const size_t o = plane + (size_t)k * T + t;
            re[o] = v.x;
            im[o] = v.y;
            mag[o] = sqrtf(v.x * v.x + v.y * v.y);
        }
    }
}

// iSTFT: CTA handles 32 hop-blocks of output (hb = (n + 256) / 256 in [hb0, hb0+32)) and needs frames hb0-1 .. hb0+31.
__global__ void __launch_bounds__(WARPS * 32) istft_kernel(const float* __restrict__ re, const float* __restrict__ im,
                                                          int T, int length, float* __restrict__ wave) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemTables* tb = reinterpret_cast<SmemTables*>(smem_raw);
    float2* fftbuf = reinterpret_cast<float2*>(smem_raw + sizeof(SmemTables));  // [2*WARPS][FBUF]
    float* frames = reinterpret_cast<float*>(fftbuf + 2 * WARPS * FBUF);         // [FRAMES+1][NFFT] windowed
    float2* xin = reinterpret_cast<float2*>(frames + (FRAMES + 1) * NFFT);       // [NBIN][FRAMES+2] spectrum tile
    constexpr int XS = FRAMES + 2;
    const int b = blockIdx.y, hb0 = 1 + blockIdx.x * FRAMES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hw = lane >> 4, hl = lane & 15;
    fill_tables(tb);
    const size_t plane = (size_t)b * NBIN * T;
    const int tf0 = hb0 - 1;  // first frame needed
    for (int idx = threadIdx.x; idx < NBIN * (FRAMES + 1); idx += blockDim.x) {
        int k = idx / (FRAMES + 1), fi = idx - k * (FRAMES + 1);
        int t = tf0 + fi;
        float2 v = make_float2(0.f, 0.f);
        if (t < T) {
            size_t o = plane + (size_t)k * T + t;
            v = make_float2(re[o], im[o]);
        }
        xin[k * XS + fi] = v;
    }
    __syncthreads();
    float2* buf = fftbuf + (warp * 2 + hw) * FBUF;
    for (int fi0 = warp * 2; fi0 < FRAMES + 1; fi0 += WARPS * 2) {
        const int fi = fi0 + hw, t = tf0 + fi;
        const bool active = fi < FRAMES + 1, has = active && t < T;   // per half warp; both halves run the transform
        // Z[k] = E[k] + i O[k]; E = (X[k] + conj X[M-k])/2 ; O = (X[k] - conj X[M-k])/2 * conj(W512^k)
        for (int k = hl; k < M; k += 16) {
            float2 z = make_float2(0.f, 0.f);
            if (has) {
                float2 xk = xin[k * XS + fi];
                float2 xm = xin[(M - k) * XS + fi];
                if (k == 0) { xk.y = 0.f; xm.y = 0.f; }  // c2r ignores Im of the DC and Nyquist bins
                float2 E = make_float2(0.5f * (xk.x + xm.x), 0.5f * (xk.y - xm.y));
                float2 D = make_float2(0.5f * (xk.x - xm.x), 0.5f * (xk.y + xm.y));
                float2 w = tb->tw[k];
                w.y = -w.y;
                float2 O = cmul(D, w);
                z = make_float2(E.x - O.y, E.y + O.x);  // E + i*O
            }
            buf[k] = z;
        }
        __syncwarp();
        half_warp_fft256<true>(buf, tb->tw, hl);
        if (active) {
            float* fr = frames + fi * NFFT;
            for (int m = hl; m < M; m += 16) {
                float2 z = buf[m];
                fr[2 * m] = z.x * (1.0f / 256.0f) * tb->win[2 * m];
                fr[2 * m + 1] = z.y * (1.0f / 256.0f) * tb->win[2 * m + 1];
            }
        }
        __syncwarp();
    }
    __syncthreads();
    float* out = wave + (size_t)b * length;
    for (int idx = threadIdx.x; idx < FRAMES * HOP; idx += blockDim.x) {
        int hbi = idx / HOP, o = idx % HOP;
        int hb = hb0 + hbi;
        long long n = (long long)hb * HOP + o - NFFT / 2;
        if (n >= length) continue;
        float val = 0.f, env = 0.f;
        if (hb <= T - 1) {  // frame t = hb, first half
            val += frames[(hbi + 1) * NFFT + o];
            env += tb->win[o] * tb->win[o];
        }
        if (hb - 1 <= T - 1) {  // frame t = hb-1, second half
            val += frames[hbi * NFFT + HOP + o];
            env += tb->win[HOP + o] * tb->win[HOP + o];
        }
        // samples beyond the signal support (n >= hop*(T-1)) are zero padding, as torch.istft(length=) does
        out[n] = (n < (long long)HOP * (T - 1)) ? val / env : 0.f;
    }
}

}  // namespace

extern "C" int nppc_stft_mri(const float* wave, int B, int L, int n_fft, int hop, float* mag, float* real,
                             float* imag, void* stream) {
    NPPC_CHECK_ARG(n_fft == NFFT && hop == HOP, "nppc_stft_mri: only n_fft=512/hop=256 is built (got %d/%d)", n_fft, hop);
    NPPC_CHECK_ARG(B > 0 && L > NFFT / 2, "nppc_stft_mri: need B>0 and L>%d for reflect padding (B=%d L=%d)", NFFT / 2, B, L);
    NPPC_CHECK_ARG(wave && mag && real && imag, "nppc_stft_mri: null pointer");
    int T = 1 + L / HOP;
    NPPC_CHECK_ARG(ensure_tables((cudaStream_t)stream) == 0, "nppc_stft_mri: table init failed");
    size_t smem = sizeof(SmemTables) + sizeof(float2) * FRAMES * SFB + sizeof(float) * (FRAMES + 1) * HOP;
    NPPC_CUDA_OK(cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(nppc::cdiv(T, FRAMES), B);
    stft_kernel<<<grid, WARPS * 32, smem, (cudaStream_t)stream>>>(wave, L, T, mag, real, imag);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_istft(const float* real, const float* imag, int B, int T, int n_fft, int hop, int length,
                          float* wave, void* stream) {
    NPPC_CHECK_ARG(n_fft == NFFT && hop == HOP, "nppc_istft: only n_fft=512/hop=256 is built (got %d/%d)", n_fft, hop);
    NPPC_CHECK_ARG(B > 0 && T > 0 && length > 0, "nppc_istft: bad sizes B=%d T=%d length=%d", B, T, length);
    NPPC_CHECK_ARG(real && imag && wave, "nppc_istft: null pointer");
    NPPC_CHECK_ARG(ensure_tables((cudaStream_t)stream) == 0, "nppc_istft: table init failed");
    size_t smem = sizeof(SmemTables) + sizeof(float2) * 2 * WARPS * FBUF + sizeof(float) * (FRAMES + 1) * NFFT +
                  sizeof(float2) * NBIN * (FRAMES + 2);
    NPPC_CUDA_OK(cudaFuncSetAttribute(istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nblocks = nppc::cdiv(length, HOP);  // hop-blocks hb = 1 .. ceil((length+256)/256)-1
    dim3 grid(nppc::cdiv(nblocks, FRAMES), B);
    istft_kernel<<<grid, WARPS * 32, smem, (cudaStream_t)stream>>>(real, imag, T, length, wave);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}
