// Row N3 (GPU data preparation): batched versions of the reference's per-item CPU dataset code.
//   mix_with_snr      AudioDataset._normalize_audio + _mix_with_snr (dataset/audio_dataset.py:92-158): RMS-normalise the clean
//                     signal to target dBFS, scale the noise to the requested SNR, mix, and prevent clipping (peak > 0.99).
//   time_to_spec_mask AudioInpaintingDataset.time_to_spec_mask (dataset/audio_dataset_inpainting.py:223-251, a Python loop over
//                     frames in the reference): a frame is kept (1) iff every sample under its (centred) window is unmasked.
// HBM-bound streaming kernels: fp64 per-sample sums, float-as-int atomicMax for the peak, grids sized from the SM count.
#include "common.cuh"

namespace {
constexpr int TPB = 256;

// stats[b] = (sum clean^2, sum noise^2)
__global__ void __launch_bounds__(TPB) mix_stats_kernel(const float* __restrict__ clean, const float* __restrict__ noise, int L,
                                                       double* __restrict__ stats) {
    __shared__ double red[32];
    const int b = blockIdx.y;
    const float* c = clean + (size_t)b * L;
    const float* n = noise + (size_t)b * L;
    float sc = 0.f, sn = 0.f;
    double dc = 0.0, dn = 0.0;
    int cnt = 0;
#pragma unroll 4
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
        const float a = c[i], q = n[i];
        sc += a * a;
        sn += q * q;
        if (++cnt == 16) { dc += sc; dn += sn; sc = sn = 0.f; cnt = 0; }
    }
    dc += sc; dn += sn;
    dc = nppc::block_sum(dc, red);
    dn = nppc::block_sum(dn, red);
    if (threadIdx.x == 0) {
        atomicAdd(&stats[2 * b], dc);
        atomicAdd(&stats[2 * b + 1], dn);
    }
}

// clean' = clean * gain; noisy = clean' + noise * scale; peak[b] = max |noisy| (float bits, non-negative -> integer max)
__global__ void __launch_bounds__(TPB) mix_apply_kernel(const float* __restrict__ clean, const float* __restrict__ noise, int L,
                                                       const double* __restrict__ stats, const float* __restrict__ snr_db,
                                                       const float* __restrict__ target_db, float* __restrict__ noisy,
                                                       float* __restrict__ clean_out, unsigned int* __restrict__ peak) {
    __shared__ float red[32];
    const int b = blockIdx.y;
    // audio_dataset.py:104-108 (fp32 arithmetic as the reference: rms -> dB -> gain)
    const float rms = sqrtf((float)(stats[2 * b] / L));
    const float gain = powf(10.0f, (target_db[b] - 20.0f * log10f(rms + 1e-8f)) / 20.0f);
    const float clean_power = gain * gain * (float)(stats[2 * b] / L);      // mean((gain*clean)^2)
    const float noise_power = (float)(stats[2 * b + 1] / L);
    const float snr_lin = powf(10.0f, snr_db[b] / 10.0f);
    const float scale = sqrtf(clean_power / (snr_lin * noise_power + 1e-8f));   // :144-145
    const size_t off = (size_t)b * L;
    float m = 0.f;
#pragma unroll 4
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
        const float c = clean[off + i] * gain;
        const float v = c + noise[off + i] * scale;
        clean_out[off + i] = c;
        noisy[off + i] = v;
        m = fmaxf(m, fabsf(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
        atomicMax(&peak[b], __float_as_uint(m));
    }
}

// clipping prevention (:152-156): if peak > 0.99 both signals are scaled by 0.99 / peak
__global__ void __launch_bounds__(TPB) mix_clip_kernel(float* __restrict__ noisy, float* __restrict__ clean_out, int L,
                                                      const unsigned int* __restrict__ peak) {
    const int b = blockIdx.y;
    const float pk = __uint_as_float(peak[b]);
    if (!(pk > 0.99f)) return;
    const float f = 0.99f / pk;
    const size_t off = (size_t)b * L;
#pragma unroll 4
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
        noisy[off + i] *= f;
        clean_out[off + i] *= f;
    }
}

// one warp per (b, frame): out = 1 if min over the frame's window of mask_time == 1, else 0
__global__ void __launch_bounds__(TPB) spec_mask_kernel(const float* __restrict__ mask_time, int L, int T_frames, int win, int hop,
                                                       int center, float* __restrict__ out, int total) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= total) return;
    const int b = wid / T_frames, t = wid - b * T_frames;
    int start = t * hop - (center ? win / 2 : 0);
    int end = start + win;
    start = start < 0 ? 0 : start;
    end = end > L ? L : end;
    float mn = 1.0f;
    for (int i = start + lane; i < end; i += 32) mn = fminf(mn, mask_time[(size_t)b * L + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if (lane == 0) out[wid] = (end > start && mn == 1.0f) ? 1.0f : 0.0f;
}

int per_sample_grid(long long n, int B) {
    long long g = (n + TPB * 4 - 1) / (TPB * 4);
    long long per = ((long long)nppc::sm_count() * 8 + B - 1) / B;
    if (g > per) g = per;
    return (int)(g < 1 ? 1 : g);
}
}  // namespace

extern "C" size_t nppc_mix_scratch_bytes(int B) { return (size_t)(B > 0 ? B : 0) * (2 * sizeof(double) + sizeof(unsigned int)) + 16; }

extern "C" int nppc_mix_with_snr(const float* clean, const float* noise, int B, int L, const float* snr_db, const float* target_db,
                                 void* scratch, float* noisy, float* clean_out, void* stream) {
    NPPC_CHECK_ARG(clean && noise && snr_db && target_db && scratch && noisy && clean_out && B > 0 && L > 0 && B <= 65535,
                   "nppc_mix_with_snr: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    double* stats = (double*)scratch;
    unsigned int* peak = (unsigned int*)(stats + 2 * (size_t)B);
    NPPC_CUDA_OK(cudaMemsetAsync(scratch, 0, nppc_mix_scratch_bytes(B), s));
    dim3 grid(per_sample_grid(L, B), B);
    mix_stats_kernel<<<grid, TPB, 0, s>>>(clean, noise, L, stats);
    mix_apply_kernel<<<grid, TPB, 0, s>>>(clean, noise, L, stats, snr_db, target_db, noisy, clean_out, peak);
    mix_clip_kernel<<<grid, TPB, 0, s>>>(noisy, clean_out, L, peak);
    NPPC_COUNT_LAUNCH(3);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_time_to_spec_mask(const float* mask_time, int B, int L, int T_frames, int win_length, int hop_length,
                                      int center, float* out, void* stream) {
    NPPC_CHECK_ARG(mask_time && out && B > 0 && L > 0 && T_frames > 0 && win_length > 0 && hop_length > 0,
                   "nppc_time_to_spec_mask: bad arguments");
    const int total = B * T_frames;
    spec_mask_kernel<<<nppc::cdiv((long long)total * 32, TPB), TPB, 0, (cudaStream_t)stream>>>(mask_time, L, T_frames, win_length,
                                                                                             hop_length, center, out, total);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}
