// a7, implementation 2: STEPWISE tensor-core LSTM (2 layers + fc) for generic (I, H) with saved state and a hand-written
// backward (BPTT) — the training step of the PC head (nppc_audio/trainer.py:100-106,234-317 runs nn.LSTM fwd + bwd,
// audio_zen/model/module/sequence_model.py:113-123), the tensor-core path for shapes the persistent kernel of lstm_tc.cu is
// not built for (e.g. the original FullSubNet full-band LSTM 257 -> 512), and — with split-precision operands — an
// fp32-accurate tensor-core LSTM for utterances whose normaliser means cancel (DESIGN.md "Conditioning").
//
// One launch per (layer, time step): a multi-segment TN GEMM on tcgen05 (TMA-fed 3-stage ring, TMEM accumulator) whose
// EPILOGUE is the LSTM cell:
//   forward  : acc[128 rows x (32 units x 4 gates)] = x_t W_ih^T + h_{t-1} W_hh^T  -> gates, c_t, h_t (fp16), saved for BPTT
//   backward : acc[128 rows x 64 units] = dZ_{t+1} W_hh + dZ^{above}_t W_ih^{above}  (or dy_t W_fc for the top layer)
//              -> dZ_t (fp16, loss-scaled), dc carry
// "Multi-segment" = the K loop walks a list of (A array, W array) pairs, so the input projection, the recurrent product, the
// layer-to-layer coupling and the hi / lo halves of split-precision operands are all just more K blocks of one accumulator.
// After the time loop the weight gradients are three big GEMMs dW = dZ^T [x | h_prev] over all T' R rows, with BOTH operands
// MN-major (row-major [rows][cols] arrays consumed "transposed" by tcgen05 through MN-major SW128 descriptors), split-K over
// CTAs into fp32 partials that a deterministic reduction folds into the nn.LSTM parameter layout.
// Gate column packing (weights, bias, saved gates, dZ):  p = n*128 + gate*32 + uu  <->  nn.LSTM row gate*H + n*32 + uu.
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace {
using namespace nppc::tc;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 3;
constexpr int NTH = 192;          // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int MAX_A = 4, MAX_W = 4, MAX_SEG = 6;

struct SegArgs {
    CUtensorMap amap[MAX_A];      // A arrays [rows][K] fp16 (box 128 x 64, SW128)
    CUtensorMap wmap[MAX_W];      // W arrays [N][K] fp16 (box BN x 64, SW128)
    int nseg;
    int seg_a[MAX_SEG], seg_w[MAX_SEG], seg_kb[MAX_SEG];
    int a_row0[MAX_A];            // row offset of this launch inside each A array (time step * R_stride)
};

template <int BN>
struct StepSmem {
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE = A_BYTES + B_BYTES;
    static constexpr int BAR_OFF = STAGES * STAGE;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <bool PRECISE> __device__ __forceinline__ float act_tanh(float x) { return PRECISE ? tanhf(x) : tanh_fast(x); }
template <bool PRECISE> __device__ __forceinline__ float act_sigm(float x) {
    return PRECISE ? 1.0f / (1.0f + expf(-x)) : fmaf(0.5f, tanh_fast(0.5f * x), 0.5f);
}

// ---- shared main loop: producer + MMA issuer; returns with the accumulator complete (acc_full) for the epilogue warps ----
template <int BN>
__device__ __forceinline__ void seg_mainloop(const SegArgs& g, unsigned char* smem, uint64_t* full, uint64_t* empty,
                                             uint64_t* acc_full, uint32_t tmem_base, int m_blk, int n_blk, int warp, int lane) {
    using S = StepSmem<BN>;
    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int sgi = 0; sgi < g.nseg; ++sgi) {
                const CUtensorMap* ma = &g.amap[g.seg_a[sgi]];
                const CUtensorMap* mw = &g.wmap[g.seg_w[sgi]];
                const int arow = g.a_row0[g.seg_a[sgi]] + m_blk * BM;
                for (int kb = 0; kb < g.seg_kb[sgi]; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* sa = smem + stage * S::STAGE;
                    mbar_arrive_expect_tx(&full[stage], S::STAGE);
                    tma_load_2d(sa, ma, &full[stage], kb * BK, arow);
                    tma_load_2d(sa + S::A_BYTES, mw, &full[stage], kb * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
            int stage = 0; uint32_t phase = 0;
            int total = 0;
            for (int sgi = 0; sgi < g.nseg; ++sgi) total += g.seg_kb[sgi];
            for (int kb = 0; kb < total; ++kb) {
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
    }
}

template <int BN>
__device__ __forceinline__ void seg_setup(const SegArgs& g, unsigned char*& smem, uint64_t*& full, uint64_t*& empty,
                                          uint64_t*& acc_full, uint32_t& tmem_base, int warp, int lane) {
    using S = StepSmem<BN>;
    extern __shared__ unsigned char smem_raw[];
    smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    empty = full + STAGES;
    acc_full = empty + STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < MAX_A; ++i) tma_prefetch_desc(&g.amap[i]);
        for (int i = 0; i < MAX_W; ++i) tma_prefetch_desc(&g.wmap[i]);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    tmem_base = *tmem_ptr;
}

template <int BN>
__device__ __forceinline__ void seg_teardown(uint32_t tmem_base, int warp) {
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_base);
    }
}

// =====================================================================================================================
// forward step.  grid (RS / 128, H / 32).  Tile columns: [gate][32 units].
//   c_prev / c_out : [RS][H] fp32 (rows of this step);  h_out (and h_lo) : [RS][H] fp16;  gates : [RS][4H] fp16 packed, or NULL
// =====================================================================================================================
template <bool PRECISE>
__global__ void __launch_bounds__(NTH) lstm_step_fwd_kernel(const __grid_constant__ SegArgs g, const float* __restrict__ bias_p,
                                                            int H, const float* __restrict__ c_prev, float* __restrict__ c_out,
                                                            __half* __restrict__ h_out, __half* __restrict__ h_lo,
                                                            __half* __restrict__ gates) {
    constexpr int BN = 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* smem; uint64_t *full, *empty, *acc_full; uint32_t tmem_base;
    seg_setup<BN>(g, smem, full, empty, acc_full, tmem_base, warp, lane);
    const int m_blk = blockIdx.x, n_blk = blockIdx.y;
    seg_mainloop<BN>(g, smem, full, empty, acc_full, tmem_base, m_blk, n_blk, warp, lane);
    if (warp >= 2) {
        const int lg = warp & 3;                               // TMEM lane group this warp may touch
        const int row = m_blk * BM + lg * 32 + lane;
        const float* bp = bias_p + n_blk * 128;
        const size_t hoff = (size_t)row * H + n_blk * 32;
        // c_{t-1} of this thread's 32 units is fetched while the MMAs run (a dependent global load per chunk inside the loop
        // was the longest part of the CTA's life)
        float4 cpre[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) cpre[q] = __ldg(reinterpret_cast<const float4*>(c_prev + hoff) + q);
        mbar_wait(acc_full, 0);
        tcgen05_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {                          // 8 units at a time
            uint32_t vi[8], vf[8], vg[8], vo[8];
            tmem_ld8(t_row + q * 8, vi);
            tmem_ld8(t_row + 32 + q * 8, vf);
            tmem_ld8(t_row + 64 + q * 8, vg);
            tmem_ld8(t_row + 96 + q * 8, vo);
            tmem_wait_ld();
            const float cp[8] = {cpre[2 * q].x, cpre[2 * q].y, cpre[2 * q].z, cpre[2 * q].w,
                                 cpre[2 * q + 1].x, cpre[2 * q + 1].y, cpre[2 * q + 1].z, cpre[2 * q + 1].w};
            float cn[8], hn[8], gi[8], gf[8], gg[8], go[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                gi[u] = act_sigm<PRECISE>(__uint_as_float(vi[u]) + __ldg(bp + q * 8 + u));
                gf[u] = act_sigm<PRECISE>(__uint_as_float(vf[u]) + __ldg(bp + 32 + q * 8 + u));
                gg[u] = act_tanh<PRECISE>(__uint_as_float(vg[u]) + __ldg(bp + 64 + q * 8 + u));
                go[u] = act_sigm<PRECISE>(__uint_as_float(vo[u]) + __ldg(bp + 96 + q * 8 + u));
                cn[u] = fmaf(gf[u], cp[u], gi[u] * gg[u]);
                hn[u] = go[u] * act_tanh<PRECISE>(cn[u]);
            }
            *reinterpret_cast<float4*>(c_out + hoff + q * 8) = make_float4(cn[0], cn[1], cn[2], cn[3]);
            *reinterpret_cast<float4*>(c_out + hoff + q * 8 + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
            auto pack8 = [](const float* v) {
                __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]), c = __floats2half2_rn(v[4], v[5]),
                        d = __floats2half2_rn(v[6], v[7]);
                return make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b), *reinterpret_cast<uint32_t*>(&c),
                                  *reinterpret_cast<uint32_t*>(&d));
            };
            const uint4 hh = pack8(hn);
            *reinterpret_cast<uint4*>(h_out + hoff + q * 8) = hh;
            if (h_lo) {
                float lo[8];
                const __half* hp = reinterpret_cast<const __half*>(&hh);
#pragma unroll
                for (int u = 0; u < 8; ++u) lo[u] = hn[u] - __half2float(hp[u]);
                *reinterpret_cast<uint4*>(h_lo + hoff + q * 8) = pack8(lo);
            }
            if (gates) {
                __half* gp = gates + (size_t)row * 4 * H + n_blk * 128 + q * 8;
                *reinterpret_cast<uint4*>(gp) = pack8(gi);
                *reinterpret_cast<uint4*>(gp + 32) = pack8(gf);
                *reinterpret_cast<uint4*>(gp + 64) = pack8(gg);
                *reinterpret_cast<uint4*>(gp + 96) = pack8(go);
            }
        }
    }
    seg_teardown<BN>(tmem_base, warp);
}

// =====================================================================================================================
// backward step.  grid (RS / 128, H / 64).  acc = dh_t (loss-scaled) for 64 units; writes dZ_t [RS][4H] fp16 packed,
// updates the dc carry [RS][H] fp32 in place (every element is owned by exactly one thread).
// =====================================================================================================================
__global__ void __launch_bounds__(NTH) lstm_step_bwd_kernel(const __grid_constant__ SegArgs g, int H,
                                                            const __half* __restrict__ gates, const float* __restrict__ c_t,
                                                            const float* __restrict__ c_prev, float* __restrict__ dc_carry,
                                                            __half* __restrict__ dz, int first /* t == T'-1: dc carry = 0 */) {
    constexpr int BN = 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* smem; uint64_t *full, *empty, *acc_full; uint32_t tmem_base;
    seg_setup<BN>(g, smem, full, empty, acc_full, tmem_base, warp, lane);
    const int m_blk = blockIdx.x, n_blk = blockIdx.y;
    seg_mainloop<BN>(g, smem, full, empty, acc_full, tmem_base, m_blk, n_blk, warp, lane);
    if (warp >= 2) {
        const int lg = warp & 3;
        const int row = m_blk * BM + lg * 32 + lane;
        // operands of one 8-unit chunk; chunk q+1 is loaded while chunk q is computed, chunk 0 while the MMAs run
        struct Ops { uint4 ri, rf, rg, ro; float4 ct0, ct1, cp0, cp1, dc0, dc1; };
        auto load_ops = [&](int q) {
            Ops o;
            const int u0 = n_blk * 64 + q * 8;
            const size_t hoff = (size_t)row * H + u0;
            const size_t goff = (size_t)row * 4 * H + (u0 >> 5) * 128 + (u0 & 31);
            o.ri = __ldg(reinterpret_cast<const uint4*>(gates + goff));
            o.rf = __ldg(reinterpret_cast<const uint4*>(gates + goff + 32));
            o.rg = __ldg(reinterpret_cast<const uint4*>(gates + goff + 64));
            o.ro = __ldg(reinterpret_cast<const uint4*>(gates + goff + 96));
            o.ct0 = __ldg(reinterpret_cast<const float4*>(c_t + hoff));
            o.ct1 = __ldg(reinterpret_cast<const float4*>(c_t + hoff + 4));
            o.cp0 = __ldg(reinterpret_cast<const float4*>(c_prev + hoff));
            o.cp1 = __ldg(reinterpret_cast<const float4*>(c_prev + hoff + 4));
            if (first) {
                o.dc0 = make_float4(0.f, 0.f, 0.f, 0.f);
                o.dc1 = o.dc0;
            } else {
                o.dc0 = *reinterpret_cast<const float4*>(dc_carry + hoff);
                o.dc1 = *reinterpret_cast<const float4*>(dc_carry + hoff + 4);
            }
            return o;
        };
        Ops nxt = load_ops(0);
        mbar_wait(acc_full, 0);
        tcgen05_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16);
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {                          // 8 units at a time: units u0 .. u0+7
            const Ops cur = nxt;
            if (q + 1 < 8) nxt = load_ops(q + 1);
            const int u0 = n_blk * 64 + q * 8;
            uint32_t vd[8];
            tmem_ld8(t_row + q * 8, vd);
            tmem_wait_ld();
            const size_t hoff = (size_t)row * H + u0;
            const size_t goff = (size_t)row * 4 * H + (u0 >> 5) * 128 + (u0 & 31);
            const __half *pi = reinterpret_cast<const __half*>(&cur.ri), *pf = reinterpret_cast<const __half*>(&cur.rf),
                         *pg = reinterpret_cast<const __half*>(&cur.rg), *po = reinterpret_cast<const __half*>(&cur.ro);
            const float ct[8] = {cur.ct0.x, cur.ct0.y, cur.ct0.z, cur.ct0.w, cur.ct1.x, cur.ct1.y, cur.ct1.z, cur.ct1.w};
            const float cpv[8] = {cur.cp0.x, cur.cp0.y, cur.cp0.z, cur.cp0.w, cur.cp1.x, cur.cp1.y, cur.cp1.z, cur.cp1.w};
            float dcc[8] = {cur.dc0.x, cur.dc0.y, cur.dc0.z, cur.dc0.w, cur.dc1.x, cur.dc1.y, cur.dc1.z, cur.dc1.w};
            float zi[8], zf[8], zg[8], zo[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float dh = __uint_as_float(vd[u]);
                const float i = __half2float(pi[u]), f = __half2float(pf[u]), gg = __half2float(pg[u]), o = __half2float(po[u]);
                const float tc = tanhf(ct[u]);
                const float dc = fmaf(dh * o, 1.0f - tc * tc, dcc[u]);
                zo[u] = dh * tc * o * (1.0f - o);
                zi[u] = dc * gg * i * (1.0f - i);
                zf[u] = dc * cpv[u] * f * (1.0f - f);
                zg[u] = dc * i * (1.0f - gg * gg);
                dcc[u] = dc * f;
            }
            *reinterpret_cast<float4*>(dc_carry + hoff) = make_float4(dcc[0], dcc[1], dcc[2], dcc[3]);
            *reinterpret_cast<float4*>(dc_carry + hoff + 4) = make_float4(dcc[4], dcc[5], dcc[6], dcc[7]);
            auto pack8 = [](const float* v) {
                __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]), c = __floats2half2_rn(v[4], v[5]),
                        d = __floats2half2_rn(v[6], v[7]);
                return make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b), *reinterpret_cast<uint32_t*>(&c),
                                  *reinterpret_cast<uint32_t*>(&d));
            };
            __half* zp = dz + goff;
            *reinterpret_cast<uint4*>(zp) = pack8(zi);
            *reinterpret_cast<uint4*>(zp + 32) = pack8(zf);
            *reinterpret_cast<uint4*>(zp + 64) = pack8(zg);
            *reinterpret_cast<uint4*>(zp + 96) = pack8(zo);
        }
    }
    seg_teardown<BN>(tmem_base, warp);
}

// =====================================================================================================================
// plain multi-segment GEMM with an fp32 epilogue: C[rows][ldc] (first ncols columns) = acc * (*mul).  grid (rows/128, N/128).
// Used for dxs = dZ0 W_ih0 over all T' R rows (the gradient that flows on into the sub-band packer).
// =====================================================================================================================
__global__ void __launch_bounds__(NTH) seg_gemm_f32_kernel(const __grid_constant__ SegArgs g, float* __restrict__ C, int ldc,
                                                           int ncols, const float* __restrict__ mul) {
    constexpr int BN = 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* smem; uint64_t *full, *empty, *acc_full; uint32_t tmem_base;
    seg_setup<BN>(g, smem, full, empty, acc_full, tmem_base, warp, lane);
    const int m_blk = blockIdx.x, n_blk = blockIdx.y;
    seg_mainloop<BN>(g, smem, full, empty, acc_full, tmem_base, m_blk, n_blk, warp, lane);
    if (warp >= 2) {
        const int lg = warp & 3;
        const size_t row = (size_t)m_blk * BM + lg * 32 + lane;
        mbar_wait(acc_full, 0);
        tcgen05_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16);
        const float s = mul ? *mul : 1.0f;
#pragma unroll 1
        for (int q = 0; q < BN / 8; ++q) {
            const int c0 = n_blk * BN + q * 8;
            if (c0 >= ncols) break;
            uint32_t v[8];
            tmem_ld8(t_row + q * 8, v);
            tmem_wait_ld();
            float* dst = C + row * ldc + c0;
            if (c0 + 8 <= ncols && (ldc & 3) == 0) {
                *reinterpret_cast<float4*>(dst) = make_float4(__uint_as_float(v[0]) * s, __uint_as_float(v[1]) * s, __uint_as_float(v[2]) * s, __uint_as_float(v[3]) * s);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(__uint_as_float(v[4]) * s, __uint_as_float(v[5]) * s, __uint_as_float(v[6]) * s, __uint_as_float(v[7]) * s);
            } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) if (c0 + u < ncols) dst[u] = __uint_as_float(v[u]) * s;
            }
        }
    }
    seg_teardown<BN>(tmem_base, warp);
}

// =====================================================================================================================
// dW = A^T B over all rows:  P[split][Mo][No] (fp32 partial) = sum_{r in split} A[r][m] * B[r][n]
//   A [rows][Mo] fp16, B [rows][No] fp16, both row-major = MN-major operands: the TMA box is [64 rows (k)][64 cols] with the
//   128-byte swizzle, i.e. the canonical MN-major SW128 atom (8 k-rows x 128 B = 1024 B; next 8 k: SBO = 1024 B; next 64
//   columns: LBO = one box = 8192 B); a K = 16 step advances the start address by 2 atoms (2048 B).
//   grid (Mo / 128, No / 64, splits).  BM = 128 (two A boxes per stage), BN = 64.
// =====================================================================================================================
constexpr int ATB_STAGES = 4;
template <int BN> struct AtbSmem {
    static constexpr int STAGE = 2 * 8192 + (BN / 64) * 8192;     // two A boxes + BN/64 B boxes of [64 k][64 cols] fp16
    static constexpr int TOTAL = ATB_STAGES * STAGE + 256 + 1024;
};

__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // leading byte offset: next 64 MN elements
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset: next 8 K rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}

template <int BN>
__global__ void __launch_bounds__(NTH) gemm_atb_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                                                       long long rows, int Mo, int No, float* __restrict__ P) {
    constexpr int ATB_STAGE = AtbSmem<BN>::STAGE;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + ATB_STAGES * ATB_STAGE);
    uint64_t* empty = full + ATB_STAGES;
    uint64_t* acc_full = empty + ATB_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int i = 0; i < ATB_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<BN>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int m_blk = blockIdx.x, n_blk = blockIdx.y, sp = blockIdx.z, nsp = gridDim.z;
    const long long kb_total = rows / BK;
    const long long kb0 = kb_total * sp / nsp, kb1 = kb_total * (sp + 1) / nsp;
    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (long long kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                unsigned char* sa = smem + stage * ATB_STAGE;
                mbar_arrive_expect_tx(&full[stage], ATB_STAGE);
                tma_load_2d(sa, &tmap_a, &full[stage], m_blk * BM, (int)(kb * BK));
                tma_load_2d(sa + 8192, &tmap_a, &full[stage], m_blk * BM + 64, (int)(kb * BK));
#pragma unroll
                for (int bx = 0; bx < BN / 64; ++bx)
                    tma_load_2d(sa + 16384 + bx * 8192, &tmap_b, &full[stage], n_blk * BN + bx * 64, (int)(kb * BK));
                if (++stage == ATB_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // D f32, A/B fp16, BOTH MN-major (bits 15, 16), M = 128, N = BN
            constexpr uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            for (long long kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[stage], phase);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(smem + stage * ATB_STAGE);
                const uint64_t da = umma_desc_mn128(sa, 8192), db = umma_desc_mn128(sa + 16384, 8192);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                    umma_bf16(tmem_base, da + (uint64_t)(k * 2048 >> 4), db + (uint64_t)(k * 2048 >> 4), idesc, (kb > kb0) || k != 0);
                umma_commit(&empty[stage]);
                if (++stage == ATB_STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(acc_full);
        }
    } else {
        const int lg = warp & 3;
        const int m = m_blk * BM + lg * 32 + lane;
        if (kb1 > kb0) {
            mbar_wait(acc_full, 0);
            tcgen05_fence_after();
        }
        const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16);
        float* dst = P + ((size_t)sp * Mo + m) * No + n_blk * BN;
#pragma unroll 1
        for (int q = 0; q < BN / 8; ++q) {
            uint32_t v[8];
            if (kb1 > kb0) {
                tmem_ld8(t_row + q * 8, v);
                tmem_wait_ld();
            } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = 0u;
            }
            *reinterpret_cast<float4*>(dst + q * 8) = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
            *reinterpret_cast<float4*>(dst + q * 8 + 4) = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<BN>(tmem_base);
    }
}

// ---- small kernels ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int packed_to_row(int p, int H) {   // p = n*128 + gate*32 + uu -> gate*H + n*32 + uu
    const int n = p >> 7, gate = (p >> 5) & 3, uu = p & 31;
    return gate * H + n * 32 + uu;
}
// fp32 W [4H][K] (nn.LSTM rows) -> fp16 [4H][KP] packed rows (zero-padded K), optional lo = fp16(w - hi)
__global__ void pack_rows_kernel(const float* __restrict__ w, int H, int K, int KP, __half* __restrict__ hi, __half* __restrict__ lo) {
    const long long n = (long long)4 * H * KP;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i / KP), k = (int)(i - (long long)p * KP);
        const float v = k < K ? fminf(fmaxf(w[(size_t)packed_to_row(p, H) * K + k], -65504.f), 65504.f) : 0.f;
        const __half h = __float2half_rn(v);
        hi[i] = h;
        if (lo) lo[i] = __float2half_rn(v - __half2float(h));
    }
}
// fp32 W [4H][K] -> fp16 W^T [NP rows (k, zero rows beyond K)][4H] with packed columns
__global__ void pack_cols_T_kernel(const float* __restrict__ w, int H, int K, int NP, __half* __restrict__ out) {
    const long long n = (long long)NP * 4 * H;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / (4 * H)), p = (int)(i - (long long)k * 4 * H);
        out[i] = __float2half_rn(k < K ? fminf(fmaxf(w[(size_t)packed_to_row(p, H) * K + k], -65504.f), 65504.f) : 0.f);
    }
}
__global__ void pack_bias_kernel(const float* __restrict__ b_ih, const float* __restrict__ b_hh, int H, float* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < 4 * H) { const int r = packed_to_row(p, H); out[p] = b_ih[r] + b_hh[r]; }
}
// fc_w [O][H] -> fp16 W_fc^T [H][64] (columns >= O zero)
__global__ void pack_fcT_kernel(const float* __restrict__ w, int O, int H, __half* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < H * 64) { const int u = i >> 6, o = i & 63; out[i] = __float2half_rn(o < O ? w[(size_t)o * H + u] : 0.f); }
}
// fp32 xs [n] -> hi / lo fp16 (the split-precision input of the precise mode)
__global__ void split_f32_kernel(const float* __restrict__ x, long long n, __half* __restrict__ hi, __half* __restrict__ lo) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = fminf(fmaxf(x[i], -65504.f), 65504.f);
        const __half h = __float2half_rn(v);
        hi[i] = h;
        lo[i] = __float2half_rn(v - __half2float(h));
    }
}

// y[r][o][t] = b[o] + sum_u W[o][u] (h[t][r][u] (+ h_lo)), one warp per (t, r), 8 units (one 16-byte load) per lane and round;
// the fc weights sit in shared memory as [O][H]: a lane reads the 8 weights of its units as two conflict-free float4.  O <= 24
__global__ void __launch_bounds__(256) fc_fwd_kernel(const __half* __restrict__ h, const __half* __restrict__ h_lo, int R, int RS, int Tp, int H,
                                                     const float* __restrict__ w, const float* __restrict__ b, int O, float* __restrict__ y) {
    extern __shared__ __align__(16) float ws[];       // [O][H]
    for (int i = threadIdx.x; i < H * O; i += blockDim.x) ws[i] = w[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (wid >= (long long)Tp * R) return;
    const int t = (int)(wid / R), r = (int)(wid - (long long)t * R);
    const __half* hr = h + ((size_t)t * RS + r) * H;
    const __half* lr = h_lo ? h_lo + ((size_t)t * RS + r) * H : nullptr;
    float acc[24];
#pragma unroll
    for (int o = 0; o < 24; ++o) acc[o] = 0.f;
    for (int u0 = lane * 8; u0 < H; u0 += 256) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(hr + u0));
        const __half* hp = reinterpret_cast<const __half*>(&raw);
        float hv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) hv[e] = __half2float(hp[e]);
        if (lr) {
            const uint4 rl = __ldg(reinterpret_cast<const uint4*>(lr + u0));
            const __half* lp = reinterpret_cast<const __half*>(&rl);
#pragma unroll
            for (int e = 0; e < 8; ++e) hv[e] += __half2float(lp[e]);
        }
#pragma unroll
        for (int o = 0; o < 24; ++o) if (o < O) {
            const float4 w0 = *reinterpret_cast<const float4*>(ws + o * H + u0), w1 = *reinterpret_cast<const float4*>(ws + o * H + u0 + 4);
            acc[o] += hv[0] * w0.x + hv[1] * w0.y + hv[2] * w0.z + hv[3] * w0.w + hv[4] * w1.x + hv[5] * w1.y + hv[6] * w1.z + hv[7] * w1.w;
        }
    }
#pragma unroll
    for (int o = 0; o < 24; ++o) if (o < O) {
        const float v = nppc::warp_sum(acc[o]);
        if (lane == 0) y[((size_t)r * O + o) * Tp + t] = v + b[o];
    }
}

// loss scale: *scale = 2^k with max|dy| * 2^k in [32, 64)  (fp16 gradients; 1 when dy is all zero), *inv = 1 / *scale
__global__ void absmax_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ bits) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(bits, __float_as_uint(m));
}
__global__ void scale_from_max_kernel(const unsigned int* __restrict__ bits, float* __restrict__ scale) {
    const float m = __uint_as_float(*bits);
    float s = 1.0f;
    if (m > 0.f && m < 3.0e38f) {
        int e;
        frexpf(m, &e);              // m = f * 2^e, f in [0.5, 1)
        s = ldexpf(1.0f, 6 - e);    // m * s in [32, 64)
    }
    scale[0] = s;
    scale[1] = 1.0f / s;
}
// dy [R][O][Tp] fp32 -> dyp [Tp][RS][64] fp16 (scaled; columns >= O and rows >= R zero)
__global__ void pack_dy_kernel(const float* __restrict__ dy, int R, int RS, int Tp, int O, const float* __restrict__ scale,
                               __half* __restrict__ dyp) {
    const long long n = (long long)Tp * RS * 64;
    const float s = scale[0];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(i & 63);
        const long long tr = i >> 6;
        const int t = (int)(tr / RS), r = (int)(tr - (long long)t * RS);
        float v = 0.f;
        if (o < O && r < R) v = dy[((size_t)r * O + o) * Tp + t] * s;
        dyp[i] = __float2half_rn(v);
    }
}
// column sums of a fp16 matrix [rows][C] in two deterministic stages: part[chunk][C] fp32
__global__ void __launch_bounds__(256) colsum_partial_kernel(const __half* __restrict__ x, long long rows, int C, float* __restrict__ part) {
    const long long per = (rows + gridDim.x - 1) / gridDim.x, r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
    for (int c = threadIdx.x * 2; c < C; c += 512) {
        float s0 = 0.f, s1 = 0.f;
        for (long long r = r0; r < r1; ++r) {
            const float2 v = __half22float2(*reinterpret_cast<const __half2*>(x + (size_t)r * C + c));
            s0 += v.x; s1 += v.y;
        }
        part[(size_t)blockIdx.x * C + c] = s0;
        part[(size_t)blockIdx.x * C + c + 1] = s1;
    }
}
// bias gradient: g_b_ih[row(p)] = g_b_hh[row(p)] = inv_scale * sum_chunks part[chunk][p]
__global__ void bias_grad_kernel(const float* __restrict__ part, int chunks, int H, const float* __restrict__ scale,
                                 float* __restrict__ g_b_ih, float* __restrict__ g_b_hh) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= 4 * H) return;
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += (double)part[(size_t)c * 4 * H + p];
    const float v = (float)(s * (double)scale[1]);
    const int r = packed_to_row(p, H);
    if (g_b_ih) g_b_ih[r] = v;
    if (g_b_hh) g_b_hh[r] = v;
}
// weight gradient from the split-K partials: grad[row(p)][k] = inv_scale * sum_s P[s][p][k]  (k < K; P has No >= K columns)
__global__ void wgrad_reduce_kernel(const float* __restrict__ P, int splits, int H, int No, int K, const float* __restrict__ scale,
                                    float* __restrict__ grad) {
    const long long n = (long long)4 * H * K;
    const float inv = scale[1];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i / K), k = (int)(i - (long long)p * K);
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += P[((size_t)sp * 4 * H + p) * No + k];
        grad[(size_t)packed_to_row(p, H) * K + k] = s * inv;
    }
}
// fc weight gradient: g_fc_w[o][u] = inv_scale * sum_s P[s][u][o]   (P: [splits][H][64])
__global__ void fcgrad_reduce_kernel(const float* __restrict__ P, int splits, int H, int O, const float* __restrict__ scale,
                                     float* __restrict__ grad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= O * H) return;
    const int o = i / H, u = i - o * H;
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += P[((size_t)sp * H + u) * 64 + o];
    grad[i] = s * scale[1];
}
// fc bias gradient: g_fc_b[o] = sum_{r,t} dy[r][o][t]: 64 chunk partials per o (fp64), then a fixed-order sum (deterministic)
__global__ void __launch_bounds__(256) fcbias_partial_kernel(const float* __restrict__ dy, int R, int O, int Tp, double* __restrict__ part) {
    __shared__ double red[32];
    const int o = blockIdx.y;
    const int per = (R + gridDim.x - 1) / gridDim.x, r0 = blockIdx.x * per, r1 = min(R, r0 + per);
    double s = 0.0;
    for (long long i = (long long)r0 * Tp + threadIdx.x; i < (long long)r1 * Tp; i += blockDim.x) {
        const int r = (int)(i / Tp), t = (int)(i - (long long)r * Tp);
        s += (double)dy[((size_t)r * O + o) * Tp + t];
    }
    s = nppc::block_sum(s, red);
    if (threadIdx.x == 0) part[o * gridDim.x + blockIdx.x] = s;
}
__global__ void fcbias_finish_kernel(const double* __restrict__ part, int chunks, int O, float* __restrict__ grad) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= O) return;
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += part[o * chunks + c];
    grad[o] = (float)s;
}

// ---- host helpers ----------------------------------------------------------------------------------------------------
struct Bump {   // carve 256-byte aligned regions out of the caller's workspace
    uintptr_t p, end;
    template <typename T> T* take(size_t n) {
        p = (p + 255) & ~(uintptr_t)255;
        T* r = reinterpret_cast<T*>(p);
        p += n * sizeof(T);
        return r;
    }
};
inline int r128(int x) { return (x + 127) / 128 * 128; }
inline int r64(int x) { return (x + 63) / 64 * 64; }

struct Layout {     // every buffer of one forward (+ backward) call inside the workspace
    // packed weights
    __half *wx[2], *wx_lo[2], *wh[2], *wh_lo[2];   // [4H][KPl], [4H][H]
    float* bias[2];
    __half *whT[2], *wx1T, *wx0T, *wfcT;           // backward operands
    // state
    __half *hseq[2], *hlo[2];                      // [(Tp+1) RS][H], zero first block
    float* cseq[2];                                // train: [(Tp+1) RS][H]; infer: [2 RS][H] ping-pong
    __half* gates[2];                              // train: [Tp RS][4H]
    __half *xhi, *xlo;                             // precise: split input [Tp RS][KP]
    // backward
    __half *dz[2], *dyp;
    float *dcc, *scale, *part, *P;
    unsigned int* maxbits;
    size_t bytes;
};

Layout carve(void* ws, int I, int H, int O, int RS, int Tp, int KP, int train, int precise) {
    Layout L{};
    Bump b{(uintptr_t)ws, 0};
    const size_t rows = (size_t)Tp * RS, rows1 = (size_t)(Tp + 1) * RS;
    for (int l = 0; l < 2; ++l) {
        const int Kl = l == 0 ? KP : H;
        L.wx[l] = b.take<__half>((size_t)4 * H * Kl);
        L.wh[l] = b.take<__half>((size_t)4 * H * H);
        L.wx_lo[l] = precise ? b.take<__half>((size_t)4 * H * Kl) : nullptr;
        L.wh_lo[l] = precise ? b.take<__half>((size_t)4 * H * H) : nullptr;
        L.bias[l] = b.take<float>((size_t)4 * H);
        L.hseq[l] = b.take<__half>(rows1 * H);
        L.hlo[l] = precise ? b.take<__half>(rows1 * H) : nullptr;
        L.cseq[l] = b.take<float>((train ? rows1 : (size_t)2 * RS) * H);
        L.gates[l] = train ? b.take<__half>(rows * 4 * H) : nullptr;
    }
    L.xhi = precise ? b.take<__half>(rows * KP) : nullptr;
    L.xlo = precise ? b.take<__half>(rows * KP) : nullptr;
    if (train) {
        for (int l = 0; l < 2; ++l) {
            L.whT[l] = b.take<__half>((size_t)r128(H) * 4 * H);
            L.dz[l] = b.take<__half>(rows * 4 * H);
        }
        L.wx1T = b.take<__half>((size_t)r128(H) * 4 * H);
        L.wx0T = b.take<__half>((size_t)r128(KP) * 4 * H);
        L.wfcT = b.take<__half>((size_t)r128(H) * 64);
        L.dyp = b.take<__half>(rows * 64);
        L.dcc = b.take<float>((size_t)RS * H);
        L.scale = b.take<float>(4);
        L.maxbits = b.take<unsigned int>(4);
        L.part = b.take<float>((size_t)1024 * 4 * H);
        const size_t pmax = (size_t)16 * 4 * H * (size_t)(H > 64 ? H : 64);
        L.P = b.take<float>(pmax);
    }
    L.bytes = (b.p - (uintptr_t)ws) + 512;
    return L;
}

int tmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    return make_tmap_bf16_2d(m, base, rows, cols, cols * 2, box_rows, 64);
}

template <typename K>
int allow_smem(K kern, int smem_bytes) {   // once per call site group, not per launch (the time loop launches ~1000 kernels)
    NPPC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    return NPPC_OK;
}
template <typename K, typename... Args>
int launch_seg(K kern, int smem_bytes, dim3 grid, cudaStream_t s, Args... args) {
    kern<<<grid, NTH, smem_bytes, s>>>(args...);
    NPPC_COUNT_LAUNCH(1);
    return NPPC_OK;
}

int run_atb(const __half* A, int Mo, const __half* B, int No, long long rows, int splits, float* P, cudaStream_t s) {
    CUtensorMap ta, tb;
    int rc = tmap(&ta, A, (uint64_t)rows, (uint64_t)Mo, 64);   // box [64 rows][64 cols]
    if (rc) return rc;
    rc = tmap(&tb, B, (uint64_t)rows, (uint64_t)No, 64);
    if (rc) return rc;
    if (No % 128 == 0) {   // 128 x 128 tiles: half the operand bytes per FLOP through L2 / shared memory
        NPPC_CUDA_OK(cudaFuncSetAttribute(gemm_atb_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtbSmem<128>::TOTAL));
        gemm_atb_kernel<128><<<dim3(Mo / 128, No / 128, splits), NTH, AtbSmem<128>::TOTAL, s>>>(ta, tb, rows, Mo, No, P);
    } else {
        NPPC_CUDA_OK(cudaFuncSetAttribute(gemm_atb_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtbSmem<64>::TOTAL));
        gemm_atb_kernel<64><<<dim3(Mo / 128, No / 64, splits), NTH, AtbSmem<64>::TOTAL, s>>>(ta, tb, rows, Mo, No, P);
    }
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

int check_shapes(const nppc_lstm_weights* w, int R, int RS, int Tp, int KP) {
    NPPC_CHECK_ARG(w && w->I > 0 && w->H > 0 && w->O > 0, "nppc_lstm_step: bad weights struct");
    NPPC_CHECK_ARG(w->H % 64 == 0, "nppc_lstm_step: hidden size must be a multiple of 64 (got %d)", w->H);
    NPPC_CHECK_ARG(w->O <= 24, "nppc_lstm_step: output size %d > 24", w->O);
    NPPC_CHECK_ARG(KP % 64 == 0 && KP >= w->I, "nppc_lstm_step: KP must be a multiple of 64 and >= I (KP=%d I=%d)", KP, w->I);
    NPPC_CHECK_ARG(R > 0 && Tp > 0 && RS % 128 == 0 && RS >= R, "nppc_lstm_step: R_stride must be a multiple of 128 and >= R");
    NPPC_CHECK_ARG((long long)(Tp + 1) * RS < (1LL << 31), "nppc_lstm_step: T'*R too large");
    for (int l = 0; l < 2; ++l)
        NPPC_CHECK_ARG(w->w_ih[l] && w->w_hh[l] && w->b_ih[l] && w->b_hh[l], "nppc_lstm_step: null weight pointer");
    NPPC_CHECK_ARG(w->fc_w && w->fc_b, "nppc_lstm_step: null fc pointer");
    return NPPC_OK;
}
}  // namespace

extern "C" size_t nppc_lstm_step_workspace_bytes(int I, int H, int O, int R_stride, int Tp, int KP, int train, int precise) {
    (void)I;
    return carve(nullptr, I, H, O, R_stride, Tp, KP, train, precise).bytes + 1024;
}

extern "C" int nppc_lstm_step_forward(const nppc_lstm_weights* w, const void* xs, int xs_is_f32, int R, int R_stride, int Tp, int KP,
                                      int train, int precise, void* workspace, size_t workspace_bytes, float* y, void* stream) {
    int rc = check_shapes(w, R, R_stride, Tp, KP);
    if (rc) return rc;
    NPPC_CHECK_ARG(xs && workspace && y, "nppc_lstm_step_forward: null pointer");
    NPPC_CHECK_ARG(!precise || xs_is_f32, "nppc_lstm_step_forward: the precise mode takes the fp32 packed input (split into hi / lo here)");
    NPPC_CHECK_ARG(precise || !xs_is_f32, "nppc_lstm_step_forward: the fast mode takes the fp16 packed input");
    const int H = w->H, RS = R_stride;
    NPPC_CHECK_ARG(workspace_bytes >= nppc_lstm_step_workspace_bytes(w->I, H, w->O, RS, Tp, KP, train, precise), "nppc_lstm_step_forward: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const Layout L = carve(workspace, w->I, H, w->O, RS, Tp, KP, train, precise);
    const size_t rows = (size_t)Tp * RS, rows1 = (size_t)(Tp + 1) * RS;
    // ---- weights (repacked every call: they change with every optimiser step, and the kernels are microseconds) ----
    for (int l = 0; l < 2; ++l) {
        const int Kin = l == 0 ? w->I : H, Kl = l == 0 ? KP : H;
        pack_rows_kernel<<<128, 256, 0, s>>>(w->w_ih[l], H, Kin, Kl, L.wx[l], L.wx_lo[l]);
        pack_rows_kernel<<<128, 256, 0, s>>>(w->w_hh[l], H, H, H, L.wh[l], L.wh_lo[l]);
        pack_bias_kernel<<<nppc::cdiv(4 * H, 256), 256, 0, s>>>(w->b_ih[l], w->b_hh[l], H, L.bias[l]);
        NPPC_CUDA_OK(cudaMemsetAsync(L.hseq[l], 0, (size_t)RS * H * 2, s));           // h_{-1} = 0
        if (precise) NPPC_CUDA_OK(cudaMemsetAsync(L.hlo[l], 0, (size_t)RS * H * 2, s));
        NPPC_CUDA_OK(cudaMemsetAsync(L.cseq[l], 0, (size_t)RS * H * 4, s));            // c_{-1} = 0
    }
    NPPC_COUNT_LAUNCH(6);
    const __half* xh = (const __half*)xs;
    if (precise) {
        split_f32_kernel<<<592, 256, 0, s>>>((const float*)xs, (long long)rows * KP, L.xhi, L.xlo);
        NPPC_COUNT_LAUNCH(1);
        xh = L.xhi;
    }
    NPPC_LAUNCH_OK();
    if ((rc = allow_smem(lstm_step_fwd_kernel<true>, StepSmem<128>::TOTAL)) || (rc = allow_smem(lstm_step_fwd_kernel<false>, StepSmem<128>::TOTAL))) return rc;
    // ---- time loop, layer by layer (layer 1 reads the finished h sequence of layer 0) ----
    for (int l = 0; l < 2; ++l) {
        SegArgs g{};
        const int Kl = l == 0 ? KP : H;
        const __half* a_in = l == 0 ? xh : L.hseq[0] + (size_t)RS * H;                  // input sequence, row block t
        const __half* a_in_lo = l == 0 ? L.xlo : (precise ? L.hlo[0] + (size_t)RS * H : nullptr);
        if ((rc = tmap(&g.amap[0], a_in, rows, Kl, 128))) return rc;
        if ((rc = tmap(&g.amap[1], L.hseq[l], rows1, H, 128))) return rc;               // h_{t-1} = row block t of the shifted buffer
        if ((rc = tmap(&g.wmap[0], L.wx[l], 4 * H, Kl, 128))) return rc;
        if ((rc = tmap(&g.wmap[1], L.wh[l], 4 * H, H, 128))) return rc;
        if (precise) {
            if ((rc = tmap(&g.amap[2], a_in_lo, rows, Kl, 128))) return rc;
            if ((rc = tmap(&g.amap[3], L.hlo[l], rows1, H, 128))) return rc;
            if ((rc = tmap(&g.wmap[2], L.wx_lo[l], 4 * H, Kl, 128))) return rc;
            if ((rc = tmap(&g.wmap[3], L.wh_lo[l], 4 * H, H, 128))) return rc;
            g.nseg = 6;   // x Wx + x_lo Wx + x Wx_lo + h Wh + h_lo Wh + h Wh_lo
            const int sa[6] = {0, 2, 0, 1, 3, 1}, sw[6] = {0, 0, 2, 1, 1, 3}, kb[6] = {Kl / 64, Kl / 64, Kl / 64, H / 64, H / 64, H / 64};
            for (int i = 0; i < 6; ++i) { g.seg_a[i] = sa[i]; g.seg_w[i] = sw[i]; g.seg_kb[i] = kb[i]; }
        } else {
            for (int i = 2; i < MAX_A; ++i) { g.amap[i] = g.amap[0]; g.wmap[i] = g.wmap[0]; }
            g.nseg = 2;
            g.seg_a[0] = 0; g.seg_w[0] = 0; g.seg_kb[0] = Kl / 64;
            g.seg_a[1] = 1; g.seg_w[1] = 1; g.seg_kb[1] = H / 64;
        }
        const dim3 grid(RS / 128, H / 32);
        for (int t = 0; t < Tp; ++t) {
            for (int i = 0; i < MAX_A; ++i) g.a_row0[i] = t * RS;
            const float* cprev = train ? L.cseq[l] + (size_t)t * RS * H : L.cseq[l] + (size_t)(t & 1) * RS * H;
            float* cout = train ? L.cseq[l] + (size_t)(t + 1) * RS * H : L.cseq[l] + (size_t)((t + 1) & 1) * RS * H;
            __half* hout = L.hseq[l] + (size_t)(t + 1) * RS * H;
            __half* hlo = precise ? L.hlo[l] + (size_t)(t + 1) * RS * H : nullptr;
            __half* gt = train ? L.gates[l] + (size_t)t * RS * 4 * H : nullptr;
            if (precise) rc = launch_seg(lstm_step_fwd_kernel<true>, StepSmem<128>::TOTAL, grid, s, g, (const float*)L.bias[l], H, cprev, cout, hout, hlo, gt);
            else rc = launch_seg(lstm_step_fwd_kernel<false>, StepSmem<128>::TOTAL, grid, s, g, (const float*)L.bias[l], H, cprev, cout, hout, hlo, gt);
            if (rc) return rc;
        }
        NPPC_LAUNCH_OK();
    }
    fc_fwd_kernel<<<nppc::cdiv((long long)Tp * R, 8), 256, sizeof(float) * H * w->O, s>>>(L.hseq[1] + (size_t)RS * H, precise ? L.hlo[1] + (size_t)RS * H : nullptr, R, RS, Tp, H,
                                                                 w->fc_w, w->fc_b, w->O, y);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

extern "C" int nppc_lstm_step_backward(const nppc_lstm_weights* w, const void* xs, int R, int R_stride, int Tp, int KP, void* workspace,
                                       size_t workspace_bytes, const float* dy, const nppc_lstm_grads* gr, float* dxs, void* stream) {
    int rc = check_shapes(w, R, R_stride, Tp, KP);
    if (rc) return rc;
    NPPC_CHECK_ARG(xs && workspace && dy && gr, "nppc_lstm_step_backward: null pointer");
    const int H = w->H, RS = R_stride, O = w->O;
    NPPC_CHECK_ARG(H % 128 == 0, "nppc_lstm_step_backward: hidden size must be a multiple of 128 (got %d)", H);
    NPPC_CHECK_ARG(workspace_bytes >= nppc_lstm_step_workspace_bytes(w->I, H, O, RS, Tp, KP, 1, 0), "nppc_lstm_step_backward: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const Layout L = carve(workspace, w->I, H, O, RS, Tp, KP, 1, 0);
    const size_t rows = (size_t)Tp * RS, rows1 = (size_t)(Tp + 1) * RS;
    const __half* xh = (const __half*)xs;
    // ---- loss scale, packed dy, transposed weights ----
    NPPC_CUDA_OK(cudaMemsetAsync(L.maxbits, 0, 4, s));
    absmax_kernel<<<592, 256, 0, s>>>(dy, (long long)R * O * Tp, L.maxbits);
    scale_from_max_kernel<<<1, 1, 0, s>>>(L.maxbits, L.scale);
    pack_dy_kernel<<<1184, 256, 0, s>>>(dy, R, RS, Tp, O, L.scale, L.dyp);
    for (int l = 0; l < 2; ++l) pack_cols_T_kernel<<<128, 256, 0, s>>>(w->w_hh[l], H, H, r128(H), L.whT[l]);
    pack_cols_T_kernel<<<128, 256, 0, s>>>(w->w_ih[1], H, H, r128(H), L.wx1T);
    pack_cols_T_kernel<<<128, 256, 0, s>>>(w->w_ih[0], H, w->I, r128(KP), L.wx0T);
    NPPC_CUDA_OK(cudaMemsetAsync(L.wfcT, 0, (size_t)r128(H) * 64 * 2, s));
    pack_fcT_kernel<<<nppc::cdiv(H * 64, 256), 256, 0, s>>>(w->fc_w, O, H, L.wfcT);
    NPPC_COUNT_LAUNCH(8);
    NPPC_LAUNCH_OK();
    if ((rc = allow_smem(lstm_step_bwd_kernel, StepSmem<64>::TOTAL)) || (rc = allow_smem(seg_gemm_f32_kernel, StepSmem<128>::TOTAL))) return rc;
    // ---- BPTT: layer 1 (top) first, then layer 0 with layer 1's dZ as the gradient from above ----
    for (int l = 1; l >= 0; --l) {
        SegArgs g{};
        if ((rc = tmap(&g.amap[0], L.dz[l], rows, 4 * H, 128))) return rc;                       // dZ_{t+1} of this layer
        if ((rc = tmap(&g.wmap[0], L.whT[l], r128(H), 4 * H, 64))) return rc;
        if (l == 1) {
            if ((rc = tmap(&g.amap[1], L.dyp, rows, 64, 128))) return rc;                        // dy_t W_fc
            if ((rc = tmap(&g.wmap[1], L.wfcT, r128(H), 64, 64))) return rc;
        } else {
            if ((rc = tmap(&g.amap[1], L.dz[1], rows, 4 * H, 128))) return rc;                   // dZ1_t W_ih1
            if ((rc = tmap(&g.wmap[1], L.wx1T, r128(H), 4 * H, 64))) return rc;
        }
        for (int i = 2; i < MAX_A; ++i) { g.amap[i] = g.amap[0]; g.wmap[i] = g.wmap[0]; }
        g.nseg = 2;
        g.seg_a[0] = 1; g.seg_w[0] = 1; g.seg_kb[0] = l == 1 ? 1 : 4 * H / 64;                   // from above, this step
        g.seg_a[1] = 0; g.seg_w[1] = 0;                                                         // recurrent, from step t+1
        const dim3 grid(RS / 128, H / 64);
        for (int t = Tp - 1; t >= 0; --t) {
            g.seg_kb[1] = t == Tp - 1 ? 0 : 4 * H / 64;
            g.a_row0[1] = t * RS;
            g.a_row0[0] = (t == Tp - 1 ? t : t + 1) * RS;
            rc = launch_seg(lstm_step_bwd_kernel, StepSmem<64>::TOTAL, grid, s, g, H, (const __half*)(L.gates[l] + (size_t)t * RS * 4 * H),
                            (const float*)(L.cseq[l] + (size_t)(t + 1) * RS * H), (const float*)(L.cseq[l] + (size_t)t * RS * H), L.dcc,
                            L.dz[l] + (size_t)t * RS * 4 * H, (int)(t == Tp - 1));
            if (rc) return rc;
        }
        NPPC_LAUNCH_OK();
    }
    // ---- weight / bias gradients: dW = dZ^T [input | h_prev] over all rows (MN-major operands, split-K partials) ----
    const int chunks = 592;
    for (int l = 0; l < 2; ++l) {
        const int Kin = l == 0 ? w->I : H, Kl = l == 0 ? KP : H;
        const __half* a_in = l == 0 ? xh : L.hseq[0] + (size_t)RS * H;
        int splits = Kl <= 64 ? 12 : 4;
        if ((rc = run_atb(L.dz[l], 4 * H, a_in, Kl, (long long)rows, splits, L.P, s))) return rc;
        wgrad_reduce_kernel<<<296, 256, 0, s>>>(L.P, splits, H, Kl, Kin, L.scale, gr->w_ih[l]);
        splits = 4;
        if ((rc = run_atb(L.dz[l], 4 * H, L.hseq[l], H, (long long)rows, splits, L.P, s))) return rc;   // h_{t-1}: un-shifted start
        wgrad_reduce_kernel<<<296, 256, 0, s>>>(L.P, splits, H, H, H, L.scale, gr->w_hh[l]);
        colsum_partial_kernel<<<chunks, 256, 0, s>>>(L.dz[l], (long long)rows, 4 * H, L.part);
        bias_grad_kernel<<<nppc::cdiv(4 * H, 256), 256, 0, s>>>(L.part, chunks, H, L.scale, gr->b_ih[l], gr->b_hh[l]);
        NPPC_COUNT_LAUNCH(4);
    }
    {   // fc: g_fc_w^T [H][64] = h1^T dyp
        const int splits = 16;
        if ((rc = run_atb(L.hseq[1] + (size_t)RS * H, H, L.dyp, 64, (long long)rows, splits, L.P, s))) return rc;
        fcgrad_reduce_kernel<<<nppc::cdiv(O * H, 256), 256, 0, s>>>(L.P, splits, H, O, L.scale, gr->fc_w);
        fcbias_partial_kernel<<<dim3(64, O), 256, 0, s>>>(dy, R, O, Tp, reinterpret_cast<double*>(L.part));
        fcbias_finish_kernel<<<1, 32, 0, s>>>(reinterpret_cast<const double*>(L.part), 64, O, gr->fc_b);
        NPPC_COUNT_LAUNCH(3);
    }
    if (dxs) {   // gradient w.r.t. the packed input: dZ0 W_ih0 -> [Tp RS][KP] fp32 (unscaled)
        SegArgs g{};
        if ((rc = tmap(&g.amap[0], L.dz[0], rows, 4 * H, 128))) return rc;
        if ((rc = tmap(&g.wmap[0], L.wx0T, r128(KP), 4 * H, 128))) return rc;
        for (int i = 1; i < MAX_A; ++i) { g.amap[i] = g.amap[0]; g.wmap[i] = g.wmap[0]; }
        g.nseg = 1;
        g.seg_a[0] = 0; g.seg_w[0] = 0; g.seg_kb[0] = 4 * H / 64;
        rc = launch_seg(seg_gemm_f32_kernel, StepSmem<128>::TOTAL, dim3((unsigned)(rows / 128), r128(KP) / 128), s, g, dxs, KP, KP, (const float*)(L.scale + 1));
        if (rc) return rc;
    }
    NPPC_LAUNCH_OK();
    (void)rows1;
    return NPPC_OK;
}

// dW-style GEMM exposed for tests and for the 1x1-convolution weight gradients of the TCN:  C[Mo][No] fp32 = A^T B,
// A [rows][Mo], B [rows][No] fp16 row-major; rows % 64 == 0, Mo % 128 == 0, No % 64 == 0.  `partials` holds splits*Mo*No floats.
__global__ void atb_sum_kernel(const float* __restrict__ P, int splits, long long n, float* __restrict__ C) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += P[(size_t)sp * n + i];
        C[i] = s;
    }
}
extern "C" int nppc_gemm_f16_atb(const void* A, const void* B, long long rows, int Mo, int No, int splits, float* partials, float* C,
                                 void* stream) {
    NPPC_CHECK_ARG(A && B && partials && C, "nppc_gemm_f16_atb: null pointer");
    NPPC_CHECK_ARG(rows > 0 && rows % 64 == 0 && Mo % 128 == 0 && No % 64 == 0 && splits >= 1 && splits <= 64,
                   "nppc_gemm_f16_atb: need rows %% 64 == 0, Mo %% 128 == 0, No %% 64 == 0, 1 <= splits <= 64");
    NPPC_CHECK_ARG(rows < (1LL << 31), "nppc_gemm_f16_atb: too many rows");
    cudaStream_t s = (cudaStream_t)stream;
    int rc = run_atb((const __half*)A, Mo, (const __half*)B, No, rows, splits, partials, s);
    if (rc) return rc;
    atb_sum_kernel<<<296, 256, 0, s>>>(partials, splits, (long long)Mo * No, C);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}
