// Row N4: the inpainting UNet's 3x3 convolutions (nppc_audio/inpainting/networks/tmp_utils.py:8-35 double_conv,
// unet.py:247-290) as an IMPLICIT GEMM on tcgen05: activations NHWC fp16, one CTA = a tile of 8 x 16 output pixels (M = 128)
// x BN output channels; the K loop walks 9 taps x (C_in / 64) channel blocks, and the A tile of tap (ky, kx) is simply the
// TMA box of the input shifted by (ky - 1, kx - 1) — the 4-D tensor map zero-fills the halo, so there is no im2col buffer,
// no padding copy and no border branch.  Up to two input tensors (the decoder's cat([skip, upsampled]) is never
// materialised: the K loop just continues into the second tensor).  Epilogue: + bias (eval-mode BatchNorm folded on the
// host) -> LeakyReLU -> fp16 -> swizzled staging -> 4-D TMA store (clipped at the image border).
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace {
using namespace nppc::tc;

constexpr int TH = 8, TW = 16;          // output tile: 8 rows x 16 columns = 128 pixels
constexpr int BM = 128, BK = 64, STAGES = 4, NTH = 192;

template <int BN>
struct ConvSmem {
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE = A_BYTES + B_BYTES;
    static constexpr int C_OFF = STAGES * STAGE;                 // staging: BN/64 boxes of [128 pixels][64 channels] fp16
    static constexpr int BAR_OFF = C_OFF + BM * BN * 2;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(NTH) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_x0, const __grid_constant__ CUtensorMap tmap_x1,
                                                         const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_y,
                                                         const float* __restrict__ bias, int c0_blocks, int c1_blocks, int tiles_w,
                                                         int tiles_h, float slope) {
    using S = ConvSmem<BN>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_x0); tma_prefetch_desc(&tmap_x1); tma_prefetch_desc(&tmap_w); tma_prefetch_desc(&tmap_y);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<BN>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    int tile = blockIdx.x;
    const int tw = tile % tiles_w; tile /= tiles_w;
    const int th = tile % tiles_h;
    const int b = tile / tiles_h;
    const int w0 = tw * TW, h0 = th * TH, n_blk = blockIdx.y;
    const int cb_total = c0_blocks + c1_blocks, kb_total = 9 * cb_total;
    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3, kx = tap - ky * 3;
                for (int cb = 0; cb < cb_total; ++cb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* sa = smem + stage * S::STAGE;
                    mbar_arrive_expect_tx(&full[stage], S::STAGE);
                    if (cb < c0_blocks) tma_load_4d(sa, &tmap_x0, &full[stage], cb * BK, w0 + kx - 1, h0 + ky - 1, b);
                    else tma_load_4d(sa, &tmap_x1, &full[stage], (cb - c0_blocks) * BK, w0 + kx - 1, h0 + ky - 1, b);
                    tma_load_2d(sa + S::A_BYTES, &tmap_w, &full[stage], (tap * cb_total + cb) * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < kb_total; ++kb) {
                mbar_wait(&full[stage], phase);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(smem + stage * S::STAGE);
                const uint64_t da = umma_desc_k128(sa), db = umma_desc_k128(sa + S::A_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                umma_commit(&empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(acc_full);
        }
    } else {
        const int lg = warp & 3, rloc = lg * 32 + lane;        // pixel of the tile = TMEM lane
        mbar_wait(acc_full, 0);
        tcgen05_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16);
        unsigned char* cs = smem + S::C_OFF;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            tmem_ld32(t_row + c * 32, v);
            tmem_wait_ld();
            uint32_t packed[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float a = __uint_as_float(v[2 * i]) + __ldg(bias + n_blk * BN + c * 32 + 2 * i);
                float d = __uint_as_float(v[2 * i + 1]) + __ldg(bias + n_blk * BN + c * 32 + 2 * i + 1);
                a = a >= 0.f ? a : slope * a;
                d = d >= 0.f ? d : slope * d;
                const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(d, -65504.f), 65504.f));
                packed[i] = *reinterpret_cast<const uint32_t*>(&h);
            }
            // staging = SWIZZLE_128B box [128 pixels][64 channels]: 16-byte chunk q of pixel r lives at q ^ (r & 7)
            unsigned char* box = cs + (c >> 1) * (BM * 128) + rloc * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int chunk = ((c & 1) * 4 + q) ^ (rloc & 7);
                *reinterpret_cast<uint4*>(box + chunk * 16) = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            }
        }
        tcgen05_fence_before();
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
#pragma unroll
            for (int cb = 0; cb < BN / 64; ++cb) tma_store_4d(&tmap_y, cs + cb * (BM * 128), n_blk * BN + cb * 64, w0, h0, b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<BN>(tmem_base);
    }
}

// weights [Cout][Cin][3][3] fp32 (BatchNorm already folded) -> fp16 [Cout][9][C0p + C1p], tap-major, the two input tensors'
// channels each padded to a multiple of 64 (zeros); input channel ci < C0 belongs to tensor 0, the rest to tensor 1
__global__ void pack_conv_w_kernel(const float* __restrict__ w, int Cout, int C0, int C1, int C0p, int C1p, __half* __restrict__ out) {
    const int Kp = 9 * (C0p + C1p);
    const long long n = (long long)Cout * Kp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i / Kp), k = (int)(i - (long long)co * Kp);
        const int tap = k / (C0p + C1p), c = k - tap * (C0p + C1p);
        int ci = -1;
        if (c < C0p) { if (c < C0) ci = c; }
        else if (c - C0p < C1) ci = C0 + (c - C0p);
        float v = 0.f;
        if (ci >= 0) v = w[((size_t)co * (C0 + C1) + ci) * 9 + tap];
        out[i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    }
}

// NCHW fp32 [B][C][H][W] -> NHWC fp16 [B][H][W][Cp] (channels >= C zero); thread = (pixel, 8-channel group): 16-byte stores
__global__ void __launch_bounds__(256) nchw_to_nhwc_f16_kernel(const float* __restrict__ x, int C, int HW, int Cp, __half* __restrict__ y) {
    const int b = blockIdx.y, G = Cp / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)HW * G; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i / G), g = (int)(i - (long long)p * G);
        __half h[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = g * 8 + e;
            h[e] = __float2half_rn(c < C ? x[((size_t)b * C + c) * HW + p] : 0.f);
        }
        *reinterpret_cast<uint4*>(y + ((size_t)b * HW + p) * Cp + g * 8) = *reinterpret_cast<const uint4*>(h);
    }
}

// MaxPool2d(2) on NHWC fp16 (tmp_utils.py:45 nn.MaxPool2d(2)): [B][H][W][C] -> [B][H/2][W/2][C]; thread = (out pixel, 8 channels)
__global__ void __launch_bounds__(256) maxpool2x2_nhwc_kernel(const __half* __restrict__ x, int H, int W, int C, __half* __restrict__ y) {
    const int b = blockIdx.y, Ho = H / 2, Wo = W / 2, G = C / 8;
    const long long n = (long long)Ho * Wo * G;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % G);
        const long long po = i / G;
        const int wo = (int)(po % Wo), ho = (int)(po / Wo);
        const __half* base = x + (((size_t)b * H + 2 * ho) * W + 2 * wo) * C + g * 8;
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(base)), c = __ldg(reinterpret_cast<const uint4*>(base + C));
        const uint4 d = __ldg(reinterpret_cast<const uint4*>(base + (size_t)W * C)), e = __ldg(reinterpret_cast<const uint4*>(base + (size_t)W * C + C));
        uint4 r;
        const __half2 *pa = reinterpret_cast<const __half2*>(&a), *pc = reinterpret_cast<const __half2*>(&c),
                      *pd = reinterpret_cast<const __half2*>(&d), *pe = reinterpret_cast<const __half2*>(&e);
        __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
        for (int q = 0; q < 4; ++q) pr[q] = __hmax2(__hmax2(pa[q], pc[q]), __hmax2(pd[q], pe[q]));
        *reinterpret_cast<uint4*>(y + (((size_t)b * Ho + ho) * Wo + wo) * C + g * 8) = r;
    }
}

// nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True) followed by the zero pad to the skip tensor's size
// (tmp_utils.py:59-82), NHWC fp16: in [B][h][w][C] -> out [B][H][W][C]; the 2h x 2w upsampled image sits at (py, px).
// thread = (out pixel, 8 channels); interpolation in fp32 like ATen's opmath, one rounding to fp16.
__global__ void __launch_bounds__(256) upsample2x_pad_nhwc_kernel(const __half* __restrict__ x, int h, int w, int C, int H, int W, int py,
                                                                  int px, __half* __restrict__ y) {
    const int b = blockIdx.y, G = C / 8, uh = 2 * h, uw = 2 * w;
    const float ry = uh > 1 ? (float)(h - 1) / (float)(uh - 1) : 0.f, rx = uw > 1 ? (float)(w - 1) / (float)(uw - 1) : 0.f;
    const long long n = (long long)H * W * G;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % G);
        const long long po = i / G;
        const int X = (int)(po % W), Y = (int)(po / W);
        const int uy = Y - py, ux = X - px;
        uint4 r = make_uint4(0, 0, 0, 0);
        if (uy >= 0 && uy < uh && ux >= 0 && ux < uw) {
            const float sy = ry * uy, sx = rx * ux;
            const int y0 = (int)sy, x0 = (int)sx;
            const int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
            const float ly = sy - y0, lx = sx - x0;
            const __half* base = x + (size_t)b * h * w * C + g * 8;
            const uint4 v00 = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y0 * w + x0) * C)), v01 = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y0 * w + x1) * C));
            const uint4 v10 = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y1 * w + x0) * C)), v11 = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y1 * w + x1) * C));
            const __half *a = reinterpret_cast<const __half*>(&v00), *c = reinterpret_cast<const __half*>(&v01),
                         *d = reinterpret_cast<const __half*>(&v10), *e = reinterpret_cast<const __half*>(&v11);
            __half* pr = reinterpret_cast<__half*>(&r);
            const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                pr[q] = __float2half_rn(w00 * __half2float(a[q]) + w01 * __half2float(c[q]) + w10 * __half2float(d[q]) + w11 * __half2float(e[q]));
        }
        *reinterpret_cast<uint4*>(y + (((size_t)b * H + Y) * W + X) * C + g * 8) = r;
    }
}

// 1x1 output convolution (unet.py outc): NHWC fp16 [B*HW][Cin] -> NCHW fp32 [B][Cout][HW], Cout <= 16.  8 lanes share a pixel
// (each loads 16 bytes of its row: a warp reads 4 rows = 512 contiguous bytes when Cin = 64), partial sums meet by shuffles.
__global__ void __launch_bounds__(256) conv1x1_out_kernel(const __half* __restrict__ x, int Cin, int HW, const float* __restrict__ w,
                                                          const float* __restrict__ bias, int Cout, float* __restrict__ y) {
    extern __shared__ float ws[];   // [Cout][Cin]
    for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) ws[i] = w[i];
    __syncthreads();
    const int b = blockIdx.y, sub = threadIdx.x & 7;
    const int p = blockIdx.x * 32 + (threadIdx.x >> 3);
    float acc[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = 0.f;
    if (p < HW) {
        const __half* xr = x + ((size_t)b * HW + p) * Cin;
        for (int c0 = sub * 8; c0 < Cin; c0 += 64) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(xr + c0));
            const __half* hp = reinterpret_cast<const __half*>(&raw);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float v = __half2float(hp[e]);
#pragma unroll
                for (int o = 0; o < 16; ++o) if (o < Cout) acc[o] = fmaf(v, ws[o * Cin + c0 + e], acc[o]);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < 16; ++o) if (o < Cout) {
        float v = acc[o];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (sub == (o & 7) && p < HW) y[((size_t)b * Cout + o) * HW + p] = v + bias[o];
    }
}

template <int BN>
int launch_conv(const void* x0, int C0p, const void* x1, int C1p, const void* w2, const float* bias, void* y, int B, int H, int W, int Cout,
                float slope, cudaStream_t s) {
    CUtensorMap tx0, tx1, tw, ty;
    int rc = make_tmap_f16_nhwc(&tx0, x0, C0p, W, H, B, 64, TW, TH);
    if (rc) return rc;
    if (x1) { if ((rc = make_tmap_f16_nhwc(&tx1, x1, C1p, W, H, B, 64, TW, TH))) return rc; }
    else tx1 = tx0;
    const int Kp = 9 * (C0p + C1p);
    if ((rc = make_tmap_bf16_2d(&tw, w2, Cout, Kp, (uint64_t)Kp * 2, BN, 64))) return rc;
    if ((rc = make_tmap_f16_nhwc(&ty, y, Cout, W, H, B, 64, TW, TH))) return rc;
    using S = ConvSmem<BN>;
    NPPC_CUDA_OK(cudaFuncSetAttribute(conv3x3_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
    conv3x3_tc_kernel<BN><<<dim3((unsigned)(tiles_w * tiles_h * B), Cout / BN), NTH, S::TOTAL, s>>>(tx0, tx1, tw, ty, bias, C0p / 64, C1p / 64,
                                                                                                 tiles_w, tiles_h, slope);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}
}  // namespace

extern "C" int nppc_conv3x3_pack_weights(const float* w, int Cout, int C0, int C1, void* out, void* stream) {
    NPPC_CHECK_ARG(w && out && Cout > 0 && C0 > 0 && C1 >= 0, "nppc_conv3x3_pack_weights: bad arguments");
    const int C0p = (C0 + 63) / 64 * 64, C1p = (C1 + 63) / 64 * 64;
    pack_conv_w_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(w, Cout, C0, C1, C0p, C1p, (__half*)out);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_conv3x3_tc(const void* x0, int C0p, const void* x1, int C1p, const void* w_packed, const float* bias, void* y, int B,
                               int H, int W, int Cout, float negative_slope, void* stream) {
    NPPC_CHECK_ARG(x0 && w_packed && bias && y, "nppc_conv3x3_tc: null pointer");
    NPPC_CHECK_ARG(B > 0 && H > 0 && W > 0 && C0p > 0 && C0p % 64 == 0 && C1p >= 0 && C1p % 64 == 0 && Cout % 64 == 0 && Cout > 0,
                   "nppc_conv3x3_tc: channel counts must be multiples of 64 (C0p=%d C1p=%d Cout=%d)", C0p, C1p, Cout);
    NPPC_CHECK_ARG((x1 != nullptr) == (C1p > 0), "nppc_conv3x3_tc: x1 and C1p must come together");
    NPPC_CHECK_ARG((long long)B * ((H + TH - 1) / TH) * ((W + TW - 1) / TW) < (1LL << 31), "nppc_conv3x3_tc: too many tiles");
    cudaStream_t s = (cudaStream_t)stream;
    if (Cout % 128 == 0) return launch_conv<128>(x0, C0p, x1, C1p, w_packed, bias, y, B, H, W, Cout, negative_slope, s);
    return launch_conv<64>(x0, C0p, x1, C1p, w_packed, bias, y, B, H, W, Cout, negative_slope, s);
}

extern "C" int nppc_nchw_to_nhwc_f16(const float* x, int B, int C, int H, int W, int Cp, void* y, void* stream) {
    NPPC_CHECK_ARG(x && y && B > 0 && C > 0 && H > 0 && W > 0 && Cp >= C && B <= 65535, "nppc_nchw_to_nhwc_f16: bad arguments");
    NPPC_CHECK_ARG(Cp % 8 == 0, "nppc_nchw_to_nhwc_f16: Cp must be a multiple of 8");
    nchw_to_nhwc_f16_kernel<<<dim3(nppc::cdiv((long long)H * W * (Cp / 8), 1024), B), 256, 0, (cudaStream_t)stream>>>(x, C, H * W, Cp, (__half*)y);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_conv1x1_out(const void* x, int B, int HW, int Cin, const float* w, const float* bias, int Cout, float* y, void* stream) {
    NPPC_CHECK_ARG(x && w && bias && y && B > 0 && HW > 0 && Cin % 8 == 0 && Cout >= 1 && Cout <= 16 && B <= 65535,
                   "nppc_conv1x1_out: bad arguments (Cin %% 8 == 0, Cout <= 16)");
    conv1x1_out_kernel<<<dim3(nppc::cdiv(HW, 32), B), 256, sizeof(float) * Cout * Cin, (cudaStream_t)stream>>>((const __half*)x, Cin, HW, w, bias,
                                                                                                              Cout, y);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_maxpool2x2_nhwc(const void* x, int B, int H, int W, int C, void* y, void* stream) {
    NPPC_CHECK_ARG(x && y && B > 0 && H >= 2 && W >= 2 && C % 8 == 0 && B <= 65535, "nppc_maxpool2x2_nhwc: bad arguments");
    const long long n = (long long)(H / 2) * (W / 2) * (C / 8);
    maxpool2x2_nhwc_kernel<<<dim3(nppc::cdiv(n, 512), B), 256, 0, (cudaStream_t)stream>>>((const __half*)x, H, W, C, (__half*)y);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_upsample2x_pad_nhwc(const void* x, int B, int h, int w, int C, int H, int W, void* y, void* stream) {
    NPPC_CHECK_ARG(x && y && B > 0 && h > 0 && w > 0 && C % 8 == 0 && H >= 2 * h && W >= 2 * w && B <= 65535,
                   "nppc_upsample2x_pad_nhwc: bad arguments (target must be at least 2h x 2w)");
    const int dy = H - 2 * h, dx = W - 2 * w;   // tmp_utils.py:76-82: F.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
    const long long n = (long long)H * W * (C / 8);
    upsample2x_pad_nhwc_kernel<<<dim3(nppc::cdiv(n, 512), B), 256, 0, (cudaStream_t)stream>>>((const __half*)x, h, w, C, H, W, dy / 2, dx / 2, (__half*)y);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}
