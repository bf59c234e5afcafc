// a1 STFT front end and a15 iSTFT back end (n_fft = win = 512, hop = 256).
//
// HBM-bound kernels: one CTA transforms 16 consecutive frames of one utterance so that the [B,F,T]
// (T contiguous) planes are read/written in 64-byte runs; the 512-point real FFT of each frame is done by one
// warp as a 256-point complex radix-2 FFT in shared memory (even/odd packing), twiddles and the periodic hann
// window staged in shared memory once per CTA.
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

// In-place 256-point complex FFT on `buf` (input in bit-reversed order), forward (e^{-i..}) or inverse (e^{+i..}),
// unnormalised. tw[k] = exp(-2*pi*i*k/512), k < 256. Executed by one full warp.
template <bool INVERSE>
__device__ __forceinline__ void warp_fft256(float2* buf, const float2* tw, int lane) {
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const int half = 1 << s;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int j = lane + 32 * q;
            int grp = j >> s, pos = j & (half - 1);
            int i0 = (grp << (s + 1)) + pos, i1 = i0 + half;
            float2 w = tw[pos << (8 - s)];
            if (INVERSE) w.y = -w.y;
            float2 a = buf[i0];
            float2 b = cmul(buf[i1], w);
            buf[i0] = make_float2(a.x + b.x, a.y + b.y);
            buf[i1] = make_float2(a.x - b.x, a.y - b.y);
        }
        __syncwarp();
    }
}

__device__ __forceinline__ int bitrev8(int m) { return (int)(__brev((unsigned)m) >> 24); }

struct __align__(16) SmemTables {
    float2 tw[M];     // exp(-2 pi i k / 512)
    float win[NFFT];  // periodic hann
};

__device__ SmemTables g_tables;   // filled once per process by init_tables_kernel
__global__ void init_tables_kernel() {
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
        float s, c;
        sincospif(-(float)k / 256.0f, &s, &c);
        g_tables.tw[k] = make_float2(c, s);
    }
    for (int n = threadIdx.x; n < NFFT; n += blockDim.x) g_tables.win[n] = 0.5f - 0.5f * cospif((float)n / 256.0f);
}
__device__ __forceinline__ void fill_tables(SmemTables* tb) {
    const float4* src = reinterpret_cast<const float4*>(&g_tables);
    float4* dst = reinterpret_cast<float4*>(tb);
    for (int i = threadIdx.x; i < (int)(sizeof(SmemTables) / 16); i += blockDim.x) dst[i] = src[i];
}
int ensure_tables(cudaStream_t s) {
    static bool done = false;
    if (!done) {
        init_tables_kernel<<<1, 256, 0, s>>>();
        if (cudaGetLastError() != cudaSuccess) return -1;
        done = true;
    }
    return 0;
}

// grid: (ceil(T/32), B), block 256. dynamic smem: tables + per-warp FFT buffers + output tile.
__global__ void __launch_bounds__(WARPS * 32) stft_kernel(const float* __restrict__ wave, int L, int T,
                                                         float* __restrict__ mag, float* __restrict__ re,
                                                         float* __restrict__ im) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemTables* tb = reinterpret_cast<SmemTables*>(smem_raw);
    float2* fftbuf = reinterpret_cast<float2*>(smem_raw + sizeof(SmemTables));       // [WARPS][M]
    float2* tile = fftbuf + WARPS * M;                                               // [NBIN][FRAMES+1]
    const int b = blockIdx.y, t0 = blockIdx.x * FRAMES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    fill_tables(tb);
    __syncthreads();
    const float* x = wave + (size_t)b * L;
    float2* buf = fftbuf + warp * M;
    for (int fi = warp; fi < FRAMES; fi += WARPS) {
        int t = t0 + fi;
        if (t >= T) break;  // warp-uniform
        long long base = (long long)t * HOP - NFFT / 2;
        // load 512 windowed samples (reflect padding, no edge repeat), pack even/odd into complex, bit-reversed
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            int m = lane + 32 * q;
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int n = 2 * m + e;
                long long p = base + n;
                if (p < 0) p = -p;
                if (p > L - 1) p = 2LL * (L - 1) - p;
                v[e] = x[p] * tb->win[n];
            }
            buf[bitrev8(m)] = make_float2(v[0], v[1]);
        }
        __syncwarp();
        warp_fft256<false>(buf, tb->tw, lane);
        // real-FFT recovery: X[k] = E[k] + W512^k O[k], E = (Z[k]+conj Z[M-k])/2, O = (Z[k]-conj Z[M-k])/(2i)
        for (int k = lane; k <= M; k += 32) {
            float2 zk = buf[k & (M - 1)];
            float2 zm = buf[(M - k) & (M - 1)];
            float2 E = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
            float2 D = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));  // (Zk - conj Zm)/2
            float2 O = make_float2(D.y, -D.x);                                     // D / i
            float2 w = (k < M) ? tb->tw[k] : make_float2(-1.0f, 0.0f);
            float2 wo = cmul(w, O);
            tile[k * (FRAMES + 1) + fi] = make_float2(E.x + wo.x, E.y + wo.y);
        }
        __syncwarp();
    }
    __syncthreads();
    const size_t plane = (size_t)b * NBIN * T;
    for (int idx = threadIdx.x; idx < NBIN * FRAMES; idx += blockDim.x) {
        int k = idx / FRAMES, fi = idx % FRAMES;
        int t = t0 + fi;
        if (t < T) {
            float2 v = tile[k * (FRAMES + 1) + fi];
            size_t o = plane + (size_t)k * T + t;
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
    float2* fftbuf = reinterpret_cast<float2*>(smem_raw + sizeof(SmemTables));  // [WARPS][M]
    float* frames = reinterpret_cast<float*>(fftbuf + WARPS * M);                // [FRAMES+1][NFFT] windowed
    float2* xin = reinterpret_cast<float2*>(frames + (FRAMES + 1) * NFFT);       // [NBIN][FRAMES+2] spectrum tile
    constexpr int XS = FRAMES + 2;
    const int b = blockIdx.y, hb0 = 1 + blockIdx.x * FRAMES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
    float2* buf = fftbuf + warp * M;
    for (int fi = warp; fi < FRAMES + 1; fi += WARPS) {
        int t = tf0 + fi;
        float* fr = frames + fi * NFFT;
        if (t >= T) {  // warp-uniform
            for (int n = lane; n < NFFT; n += 32) fr[n] = 0.f;
            continue;
        }
        // Z[k] = E[k] + i O[k]; E = (X[k] + conj X[M-k])/2 ; O = (X[k] - conj X[M-k])/2 * conj(W512^k)
        for (int k = lane; k < M; k += 32) {
            float2 xk = xin[k * XS + fi];
            float2 xm = xin[(M - k) * XS + fi];
            if (k == 0) { xk.y = 0.f; xm.y = 0.f; }  // c2r ignores Im of the DC and Nyquist bins
            float2 E = make_float2(0.5f * (xk.x + xm.x), 0.5f * (xk.y - xm.y));
            float2 D = make_float2(0.5f * (xk.x - xm.x), 0.5f * (xk.y + xm.y));
            float2 w = tb->tw[k];
            w.y = -w.y;
            float2 O = cmul(D, w);
            buf[bitrev8(k)] = make_float2(E.x - O.y, E.y + O.x);  // E + i*O
        }
        __syncwarp();
        warp_fft256<true>(buf, tb->tw, lane);
        for (int m = lane; m < M; m += 32) {
            float2 z = buf[m];
            fr[2 * m] = z.x * (1.0f / 256.0f) * tb->win[2 * m];
            fr[2 * m + 1] = z.y * (1.0f / 256.0f) * tb->win[2 * m + 1];
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
    size_t smem = sizeof(SmemTables) + sizeof(float2) * WARPS * M + sizeof(float2) * NBIN * (FRAMES + 1);
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
    size_t smem = sizeof(SmemTables) + sizeof(float2) * WARPS * M + sizeof(float) * (FRAMES + 1) * NFFT +
                  sizeof(float2) * NBIN * (FRAMES + 2);
    NPPC_CUDA_OK(cudaFuncSetAttribute(istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nblocks = nppc::cdiv(length, HOP);  // hop-blocks hb = 1 .. ceil((length+256)/256)-1
    dim3 grid(nppc::cdiv(nblocks, FRAMES), B);
    istft_kernel<<<grid, WARPS * 32, smem, (cudaStream_t)stream>>>(real, imag, T, length, wave);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}
