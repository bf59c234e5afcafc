// a7, implementation 1: bf16 tcgen05 tensor-core LSTM (2 layers + fc), H = 384.
//
// Per layer:  (1) input projection for ALL steps as one big GEMM (gemm_tc.cu):  Zx[T'*R, 4H] = X * W_ih^T + (b_ih+b_hh)
//             (2) persistent recurrent kernel: one CTA owns 128 sequences for all T' steps.
//
// Recurrent kernel, per step t and per gate chunk j (32 hidden units = 128 gate columns, 12 chunks):
//   tensor core : acc[128 x 128] (TMEM cols 0..127) = h_{t-1}[128 x 384] (smem, bf16, K-major SW128) * W_hh[chunk j]^T
//                 W_hh streams from L2 through a 6-stage TMA ring (it is shared by every CTA and stays L2-resident)
//   epilogue    : 8 warps; thread = (row, 16-hidden-unit half): acc + Zx -> sigmoid/tanh -> c (fp32, resident in TMEM
//                 cols 128..511) -> h ; h_t is written (bf16) to the time-major h sequence in global memory, which is
//                 both the next layer's GEMM operand and this CTA's own A operand for step t+1 (re-loaded by TMA, so
//                 L2 acts as the double buffer that shared memory has no room for); last layer: fc partial sums.
// Gate columns are permuted at pack time so that a thread's 4 gates x 16 units are contiguous in Zx and in TMEM:
//   packed column = chunk*128 + half*64 + gate*16 + u   <->   nn.LSTM row gate*H + chunk*32 + half*16 + u.
#include <stdlib.h>
#include <string.h>
#include <cuda_fp16.h>
#include "lstm_plan.cuh"
#include "tc_common.cuh"

namespace nppc {
int gemm_16bit_tn(const void* A, const void* W, const float* bias, void* C, long long M, int N, int K, int f16,
                  cudaStream_t s);
}

namespace {
using namespace nppc::tc;

constexpr int H = 384;
constexpr int H4 = 4 * H;
constexpr int ROWS = 128;           // sequences per CTA
constexpr int CH = 32;              // hidden units per chunk
constexpr int NCHUNK = H / CH;      // 12
constexpr int NSLAB = H / 64;       // 6 K-slabs of 64
constexpr int SLAB_BYTES = ROWS * 64 * 2;   // 16 KB (A slab and W stage have the same shape: 128 x 64 bf16)
constexpr int NTHREADS = 384;       // warps 0-3: TMA-W, MMA, TMEM alloc, TMA-A ; warps 4-11: epilogue
constexpr int OPMAX = 24;

constexpr int HST_BYTES = ROWS * CH * 2;   // one chunk of h_t for the CTA's rows: [128][32] bf16 = 8 KB

struct RecSmem {
    static constexpr int NST = 6;                                        // W ring stages
    static constexpr int A_OFF = 0;                                      // h_{t-1}: 6 slabs [128][64] bf16, SW128
    static constexpr int W_OFF = NSLAB * SLAB_BYTES;                     // W_hh ring
    static constexpr int HST_OFF = W_OFF + NST * SLAB_BYTES;             // 2 x [128][32] fp16 staging for the TMA store of h_t
    static constexpr int X_OFF = HST_OFF + 2 * HST_BYTES;                // layer 0 (fused input projection): x_t [128][64] fp16
    static constexpr int BAR_OFF = X_OFF + SLAB_BYTES;
#ifdef NPPC_REC_TRACE
    static constexpr int TRACE_OFF = BAR_OFF + 256;
    static constexpr int TOTAL = TRACE_OFF + 4 * 12 * 16 * 8 + 1024;
#else
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
#endif
    static_assert(TOTAL <= 227 * 1024, "shared memory budget exceeded");
};

#ifdef NPPC_REC_TRACE
__device__ long long g_trace[4 * 12 * 16];
#define TRACE(slot) do { if ((t) >= 5 && (t) < 9 && (threadIdx.x & 31) == 0) trace_s[(((t) - 5) * 12 + (j)) * 16 + (slot)] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

// CL = cluster size: the CL CTAs of a cluster share ONE W_hh stream — slab s is fetched from L2 by CTA (s % CL) and
// multicast into every CTA's ring.
// Buffers are time-major with RS (multiple of 128) rows per step:  zx in the interleaved layout written by the GEMM
// ([m_blk][chunk][warp-quad][half][piece][lane][8 bf16], m_blk = t*tiles + tile), hseq row-major [T'*RS][H].
// FUSE_X (layer 0): the K=64 input projection is fused — x_t is a 7th A slab (TMA from the packed time-major input),
// W_ih a 7th weight slab per chunk, and the epilogue adds the bias instead of streaming pre-activations from HBM.
// FUSE_FC (last layer): y_{t-1} = W_fc h_{t-1} + b is a 13th, 16-column mini-chunk issued at the start of step t (h_{t-1} is the
// A operand that was just reloaded), plus one pseudo-step after the last; its 128 x 16 fp32 result is read from TMEM cols 0..15.
template <int CL, bool FUSE_X, bool FUSE_FC>
__global__ void __launch_bounds__(NTHREADS, 1)
lstm_rec_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
                const __grid_constant__ CUtensorMap tmap_hst, const __grid_constant__ CUtensorMap tmap_x,
                const __grid_constant__ CUtensorMap tmap_wx, const __grid_constant__ CUtensorMap tmap_fc,
                const uint4* __restrict__ zx, const float* __restrict__ bias, int RS, int Tp,
                const float* __restrict__ fc_b, int O, int R, float* __restrict__ y) {
    constexpr int NST = RecSmem::NST;
    constexpr int NK = NSLAB + (FUSE_X ? 1 : 0);   // weight slabs per chunk
    constexpr int FC_SLAB_BYTES = 16 * 64 * 2;     // [16 outputs][64 k] fp16
    const int Tx = Tp + (FUSE_FC ? 1 : 0);         // steps incl. the fc-only pseudo-step
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + RecSmem::BAR_OFF);
    uint64_t* w_empty = w_full + NST;
    uint64_t* a_full = w_empty + NST;      // [NSLAB]
    uint64_t* acc_full = a_full + NSLAB;
    uint64_t* acc_empty = acc_full + 1;
    uint64_t* h_ready = acc_empty + 1;
    uint64_t* x_full = h_ready + 1;
    uint64_t* x_free = x_full + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(x_free + 1);
#ifdef NPPC_REC_TRACE
    long long* trace_s = reinterpret_cast<long long*>(smem + RecSmem::TRACE_OFF);
    for (int i = threadIdx.x; i < 4 * 12 * 16; i += NTHREADS) trace_s[i] = 0;
#endif

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = RS / ROWS;
    const int tile = blockIdx.x;          // CTAs beyond `tiles` (cluster padding) redo the last tile's loads, store nothing
    const bool live = tile < tiles;
    const int tile_c = live ? tile : tiles - 1;
    const int row0 = tile_c * ROWS;

    // zero the A operand (h_{-1} = 0)
    for (int i = threadIdx.x; i < NSLAB * SLAB_BYTES / 16; i += NTHREADS)
        reinterpret_cast<uint4*>(smem + RecSmem::A_OFF)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_h);
        tma_prefetch_desc(&tmap_hst);
        if (FUSE_X) { tma_prefetch_desc(&tmap_x); tma_prefetch_desc(&tmap_wx); }
        if (FUSE_FC) tma_prefetch_desc(&tmap_fc);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], CL); }
        for (int i = 0; i < NSLAB; ++i) mbar_init(&a_full[i], 1);
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 256);
        mbar_init(h_ready, 1);
        mbar_init(x_full, 1);
        mbar_init(x_free, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // peers' barriers must be initialised before any multicast / remote arrive
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1);

    if (warp == 0) {
        // ---- W_hh producer: 72 slabs per step, independent of t (runs ahead across step boundaries) ----
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            uint32_t s = 0;
            for (int t = 0; t < Tx; ++t) {
                if (FUSE_FC && t >= 1)
                    for (int k = 0; k < NSLAB; ++k, ++s) {   // fc weights for y_{t-1}
                        mbar_wait(&w_empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&w_full[stage], FC_SLAB_BYTES);
                        unsigned char* dst = smem + RecSmem::W_OFF + stage * SLAB_BYTES;
                        if (CL == 1) tma_load_2d(dst, &tmap_fc, &w_full[stage], k * 64, 0);
                        else if (s % CL == crank) tma_load_2d_mcast(dst, &tmap_fc, &w_full[stage], k * 64, 0, CMASK);
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                if (t >= Tp) break;
                for (int j = 0; j < NCHUNK; ++j)
                    for (int k = 0; k < NK; ++k, ++s) {
                        mbar_wait(&w_empty[stage], phase ^ 1);   // every CTA of the cluster has released this slot
                        mbar_arrive_expect_tx(&w_full[stage], SLAB_BYTES);
                        unsigned char* dst = smem + RecSmem::W_OFF + stage * SLAB_BYTES;
                        const CUtensorMap* tm = (FUSE_X && k == NSLAB) ? &tmap_wx : &tmap_w;
                        const int kc = (FUSE_X && k == NSLAB) ? 0 : k * 64;
                        if (CL == 1) tma_load_2d(dst, tm, &w_full[stage], kc, j * 128);
                        else if (s % CL == crank) tma_load_2d_mcast(dst, tm, &w_full[stage], kc, j * 128, CMASK);
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 3) {
        // ---- A producer: reload h_t (TMA-stored by the epilogue) as the A operand of step t+1 ----
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        if (lane == 0) {
            for (int k = 0; k < NSLAB; ++k) mbar_arrive(&a_full[k]);  // step 0: zeros already in place
            if (FUSE_X) {
                mbar_arrive_expect_tx(x_full, SLAB_BYTES);
                tma_load_2d(smem + RecSmem::X_OFF, &tmap_x, x_full, 0, row0);
            }
            for (int t = 1; t < Tx; ++t) {
                if (FUSE_X && t < Tp) {   // x_t: needs only the previous step's x MMAs to be done (independent of h)
                    mbar_wait(x_free, (t - 1) & 1);
                    mbar_arrive_expect_tx(x_full, SLAB_BYTES);
                    tma_load_2d(smem + RecSmem::X_OFF, &tmap_x, x_full, 0, t * RS + row0);
                }
                mbar_wait(h_ready, (t - 1) & 1);
                for (int k = 0; k < NSLAB; ++k) {
                    mbar_arrive_expect_tx(&a_full[k], SLAB_BYTES);
                    tma_load_2d(smem + RecSmem::A_OFF + k * SLAB_BYTES, &tmap_h, &a_full[k], k * 64, (t - 1) * RS + row0);
                }
            }
        }
    } else if (warp == 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    } else if (warp == 1) {
        // ---- MMA issuer: the whole warp runs the (warp-uniform) control flow so descriptors stay in uniform registers;
        //      one elected lane issues tcgen05.mma / tcgen05.commit ----
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        {
            constexpr uint32_t idesc = umma_idesc_f16(ROWS, 128);
            int stage = 0; uint32_t phase = 0;
            uint32_t it = 0;  // chunk counter for the acc_empty parity
            const uint32_t a_base = smem_u32(smem + RecSmem::A_OFF);
            const uint32_t w_base = smem_u32(smem + RecSmem::W_OFF);
            const uint32_t x_base = smem_u32(smem + RecSmem::X_OFF);
            for (int t = 0; t < Tx; ++t) {
                if (FUSE_FC && t >= 1) {   // y_{t-1}: [128 x 16] = h_{t-1} * W_fc^T into TMEM cols 0..15
                    constexpr uint32_t idesc_fc = umma_idesc_f16(ROWS, 16);
                    mbar_wait(acc_empty, (it & 1) ^ 1);
                    tcgen05_fence_after();
                    for (int k = 0; k < NSLAB; ++k) {
                        mbar_wait(&a_full[k], t & 1);
                        mbar_wait(&w_full[stage], phase);
                        tcgen05_fence_after();
                        const uint64_t da = umma_desc_k128(a_base + k * SLAB_BYTES);
                        const uint64_t db = umma_desc_k128(w_base + stage * SLAB_BYTES);
                        if (elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_base, da + 2 * kk, db + 2 * kk, idesc_fc, (k | kk) != 0);
                            if (CL == 1) umma_commit(&w_empty[stage]);
                            else umma_commit_mcast(&w_empty[stage], CMASK);
                            if (k == NSLAB - 1) umma_commit(acc_full);
                        }
                        __syncwarp();
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                    ++it;
                }
                if (t >= Tp) break;
                for (int j = 0; j < NCHUNK; ++j, ++it) {
                    mbar_wait(acc_empty, (it & 1) ^ 1);
                    TRACE(0);
                    tcgen05_fence_after();
                    for (int k = 0; k < NK; ++k) {
                        const bool is_x = FUSE_X && k == NSLAB;
                        if (j == 0) {
                            if (is_x) mbar_wait(x_full, t & 1);
                            else mbar_wait(&a_full[k], t & 1);
                        }
                        mbar_wait(&w_full[stage], phase);
                        if (k == 0) TRACE(1);
                        if (k == 3) TRACE(2);
                        if (k == 5) TRACE(3);
                        tcgen05_fence_after();
                        const uint64_t da = umma_desc_k128(is_x ? x_base : a_base + k * SLAB_BYTES);
                        const uint64_t db = umma_desc_k128(w_base + stage * SLAB_BYTES);
                        if (elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_base, da + 2 * kk, db + 2 * kk, idesc, (k | kk) != 0);
                            if (CL == 1) umma_commit(&w_empty[stage]);
                            else umma_commit_mcast(&w_empty[stage], CMASK);
                            if (k == NK - 1) umma_commit(acc_full);
                            if (is_x && j == NCHUNK - 1) umma_commit(x_free);   // x_t consumed by every chunk of this step
                        }
                        __syncwarp();
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                    TRACE(4);
                }
            }
        }
    } else if (warp >= 4) {
        // ---- epilogue: thread = (row = 32*(warp%4) + lane, half = (warp-4)/4) ----
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const int ew = warp & 3, half = (warp - 4) >> 2;
        const int rloc = ew * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(ew * 32) << 16);
        const uint32_t t_acc = t_lane + half * 64;
        unsigned char* hst = smem + RecSmem::HST_OFF;
        // c_0 = 0
        {
            uint32_t z[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = 0u;
            for (int j = 0; j < NCHUNK; ++j) tmem_st16(t_lane + 128 + j * CH + half * 16, z);
            tmem_wait_st();
        }
        uint32_t it = 0;
        // software-pipelined Zx stream: piece q of this thread's 64 pre-activations (i,f,g,o x 16 units, fp16) for chunk
        // (t, j) is the uint4 at ((((t*tiles + tile)*12 + j)*4 + ew)*2 + half)*256 + q*32 + lane  (layout written by the
        // GEMM epilogue) -> each warp-wide load reads 512 contiguous bytes.  Chunk it+1 is requested while it is computed.
        auto zx_ptr = [&](int t, int j) -> const uint4* {
            return zx + (((((size_t)t * tiles + tile_c) * NCHUNK + j) * 4 + ew) * 2 + half) * 256 + lane;
        };
        uint4 zraw[8];
        if (!FUSE_X) {
            const uint4* zp = zx_ptr(0, 0);
#pragma unroll
            for (int q = 0; q < 8; ++q) zraw[q] = __ldg(zp + q * 32);
        }
        const bool valid = live && (row0 + rloc) < R;
        for (int t = 0; t < Tx; ++t) {
            if (FUSE_FC && t >= 1) {   // fc mini-chunk: y[row][o][t-1] from TMEM cols 0..15 (half 0 threads own the rows)
                mbar_wait(acc_full, it & 1);
                tcgen05_fence_after();
                uint32_t yv[16];
                if (half == 0) {
                    tmem_ld16(t_lane, yv);
                    tmem_wait_ld();
                }
                tcgen05_fence_before();
                mbar_arrive(acc_empty);
                if (half == 0 && valid) {
                    float* dst = y + (size_t)(row0 + rloc) * O * Tp + (t - 1);
#pragma unroll
                    for (int o = 0; o < 16; ++o)
                        if (o < O) dst[(size_t)o * Tp] = __uint_as_float(yv[o]) + __ldg(fc_b + o);
                }
                ++it;
            }
            if (t >= Tp) break;
#pragma unroll 1
            for (int j = 0; j < NCHUNK; ++j, ++it) {
                mbar_wait(acc_full, it & 1);
                if (threadIdx.x == 128) TRACE(5);
                if (threadIdx.x == 352) TRACE(10);
                tcgen05_fence_after();
                uint32_t gi[16], gf[16], gg[16], go[16], cc[16];
                tmem_ld16(t_acc + 0, gi);
                tmem_ld16(t_acc + 16, gf);
                tmem_ld16(t_acc + 32, gg);
                tmem_ld16(t_acc + 48, go);
                tmem_ld16(t_lane + 128 + j * CH + half * 16, cc);
                tmem_wait_ld();
                tcgen05_fence_before();
                mbar_arrive(acc_empty);  // accumulator is in registers: the next chunk's MMAs may start
                if (threadIdx.x == 128) TRACE(6);
                if (threadIdx.x == 352) TRACE(11);
                uint4 zcur[8];
                if (!FUSE_X) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) zcur[q] = zraw[q];
                    int jn = j + 1, tn = t;
                    if (jn == NCHUNK) { jn = 0; tn = t + 1; }
                    if (tn < Tp) {
                        const uint4* zp = zx_ptr(tn, jn);
#pragma unroll
                        for (int q = 0; q < 8; ++q) zraw[q] = __ldg(zp + q * 32);
                    }
                }
                const __half2* zh = reinterpret_cast<const __half2*>(zcur);  // zh[gate*8 + u/2] = fp16 pair (u, u+1)
                const float* bj = bias + j * 128 + half * 64;                // fused: warp-uniform (broadcast) bias loads
                float hv[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    float ai, af, ag, ao;
                    if (FUSE_X) {
                        ai = __ldg(bj + u); af = __ldg(bj + 16 + u); ag = __ldg(bj + 32 + u); ao = __ldg(bj + 48 + u);
                    } else {
                        const float2 pi = __half22float2(zh[u >> 1]), pf = __half22float2(zh[8 + (u >> 1)]);
                        const float2 pg = __half22float2(zh[16 + (u >> 1)]), po = __half22float2(zh[24 + (u >> 1)]);
                        ai = (u & 1) ? pi.y : pi.x; af = (u & 1) ? pf.y : pf.x;
                        ag = (u & 1) ? pg.y : pg.x; ao = (u & 1) ? po.y : po.x;
                    }
                    float zi = __uint_as_float(gi[u]) + ai;
                    float zf = __uint_as_float(gf[u]) + af;
                    float zg = __uint_as_float(gg[u]) + ag;
                    float zo = __uint_as_float(go[u]) + ao;
                    float c = sigmoid_fast(zf) * __uint_as_float(cc[u]) + sigmoid_fast(zi) * tanh_fast(zg);
                    cc[u] = __float_as_uint(c);
                    hv[u] = sigmoid_fast(zo) * tanh_fast(c);
                }
                tmem_st16(t_lane + 128 + j * CH + half * 16, cc);
                if (threadIdx.x == 128) TRACE(7);
                if (threadIdx.x == 352) TRACE(12);
                // h_t chunk -> staging tile [128 rows][32 units] bf16 -> one TMA store per chunk (full-line writes)
                if (threadIdx.x == 128) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("bar.sync 2, 256;" ::: "memory");   // staging buffer (j & 1) is free
                {
                    uint32_t hp[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        __half2 p = __floats2half2_rn(hv[2 * u], hv[2 * u + 1]);
                        hp[u] = *reinterpret_cast<uint32_t*>(&p);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(hst + (j & 1) * HST_BYTES + rloc * (CH * 2) + half * 32);
                    dst[0] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                    dst[1] = make_uint4(hp[4], hp[5], hp[6], hp[7]);
                }
                fence_proxy_async_smem();
                asm volatile("bar.sync 3, 256;" ::: "memory");
                if (threadIdx.x == 128 && live) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmap_hst),
                                 "r"(smem_u32(hst + (j & 1) * HST_BYTES)), "r"(j * CH), "r"(t * RS + row0)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (threadIdx.x == 128) TRACE(8);
                if (threadIdx.x == 352) TRACE(13);
            }
            // end of step: h_t fully written (async proxy) -> let the A producer reload it for step t+1
            tmem_wait_st();
            if (threadIdx.x == 128) {
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                fence_proxy_async_all();
                mbar_arrive(h_ready);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
#ifdef NPPC_REC_TRACE
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < 4 * 12 * 16; i += NTHREADS) g_trace[i] = trace_s[i];
#endif
    if (CL > 1) cluster_sync_all();  // no CTA may exit while peers can still multicast into it / arrive on its barriers
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// fp32 [4H][K] (nn.LSTM row order) -> bf16 [4H][KP] with permuted rows, zero-padded K
__global__ void pack_w_kernel(const float* __restrict__ w, int K, int KP, __half* __restrict__ out) {
    long long n = (long long)H4 * KP;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int p = (int)(i / KP), k = (int)(i - (long long)p * KP);
        int chunk = p >> 7, half = (p >> 6) & 1, gate = (p >> 4) & 3, u = p & 15;
        int src = gate * H + chunk * CH + half * 16 + u;
        out[i] = __float2half_rn(k < K ? fminf(fmaxf(w[(size_t)src * K + k], -65504.f), 65504.f) : 0.f);
    }
}
__global__ void pack_fc_kernel(const float* __restrict__ w, int O, __half* __restrict__ out) {  // [O][H] f32 -> [16][H] fp16
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 16 * H) out[i] = __float2half_rn(i / H < O ? fminf(fmaxf(w[i], -65504.f), 65504.f) : 0.f);
}
__global__ void pack_b_kernel(const float* __restrict__ b, float* __restrict__ out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < H4) {
        int chunk = p >> 7, half = (p >> 6) & 1, gate = (p >> 4) & 3, u = p & 15;
        out[p] = b[gate * H + chunk * CH + half * 16 + u];
    }
}

// fc_output_layer over the h sequence of the last layer: y[row][o][t] = b[o] + sum_k W[o][k] * h[t][row][k]
// (sequence_model.py:79,119).  One CTA = 64 (t,row) rows staged in shared memory by coalesced 16-byte loads; warp w
// covers K-quarter (w & 3) of rows 32*(w >> 2) + lane, so weight reads are warp-broadcast float4; the four partial
// sums per row meet in shared memory.  One pass over hseq.
constexpr int FC_ROWS = 64;
constexpr int FC_TG = 8;     // frames per CTA: each (row, o) pair writes 8 consecutive floats
template <int OQ>   // OQ = ceil(O / 4)
__global__ void __launch_bounds__(256) lstm_fc_kernel(const __half* __restrict__ hseq, int R, int RS, int Tp,
                                                      const float* __restrict__ fc_w, const float* __restrict__ fc_b, int O,
                                                      float* __restrict__ y) {
    extern __shared__ __align__(16) unsigned char fsm[];
    constexpr int HP = H + 8;                         // padded row (fp16 elements): conflict-free row stride
    constexpr int KQ = H / 4;                         // 96 k per warp
    __half* hs = reinterpret_cast<__half*>(fsm);      // [FC_ROWS][HP]
    float4* wt = reinterpret_cast<float4*>(fsm + FC_ROWS * HP * 2);                  // [H][OQ] transposed, zero-padded
    float* part = reinterpret_cast<float*>(fsm + FC_ROWS * HP * 2 + H * OQ * 16);    // [4][FC_ROWS][OQ*4]
    const int t0 = blockIdx.y * FC_TG;
    const int row0 = blockIdx.x * FC_ROWS;
    for (int i = threadIdx.x; i < H * OQ * 4; i += blockDim.x) {
        int k = i / (OQ * 4), o = i - k * (OQ * 4);
        reinterpret_cast<float*>(wt)[i] = o < O ? fc_w[(size_t)o * H + k] : 0.f;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kq = warp & 3, r = (warp >> 2) * 32 + lane;
    const __half2* hr = reinterpret_cast<const __half2*>(hs + r * HP + kq * KQ);
    const float4* wq = wt + (size_t)kq * KQ * OQ;
    constexpr int NP = (FC_ROWS * OQ * 4 + 255) / 256;   // (row, o) pairs finished by each thread
    float res[NP][FC_TG];
    for (int tt = 0; tt < FC_TG; ++tt) {
        const int t = t0 + tt;
        if (t >= Tp) break;
        __syncthreads();   // previous tile fully consumed
        for (int i = threadIdx.x; i < FC_ROWS * (H / 8); i += blockDim.x) {
            int rr = i / (H / 8), c = i - rr * (H / 8);
            uint4 v = make_uint4(0, 0, 0, 0);
            if (row0 + rr < R) v = __ldg(reinterpret_cast<const uint4*>(hseq + ((size_t)t * RS + row0 + rr) * H) + c);
            *reinterpret_cast<uint4*>(hs + rr * HP + c * 8) = v;
        }
        __syncthreads();
        float acc[OQ * 4];
#pragma unroll
        for (int o = 0; o < OQ * 4; ++o) acc[o] = 0.f;
#pragma unroll 4
        for (int k2 = 0; k2 < KQ / 2; ++k2) {
            float2 hv = __half22float2(hr[k2]);
#pragma unroll
            for (int q = 0; q < OQ; ++q) {
                float4 w0 = wq[(2 * k2) * OQ + q], w1 = wq[(2 * k2 + 1) * OQ + q];
                acc[4 * q + 0] = fmaf(hv.x, w0.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(hv.x, w0.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(hv.x, w0.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(hv.x, w0.w, acc[4 * q + 3]);
                acc[4 * q + 0] = fmaf(hv.y, w1.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(hv.y, w1.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(hv.y, w1.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(hv.y, w1.w, acc[4 * q + 3]);
            }
        }
#pragma unroll
        for (int o = 0; o < OQ * 4; ++o) part[((size_t)kq * FC_ROWS + r) * (OQ * 4) + o] = acc[o];
        __syncthreads();
#pragma unroll
        for (int pi = 0; pi < NP; ++pi) {
            int i = threadIdx.x + pi * 256;
            float v = 0.f;
            if (i < FC_ROWS * OQ * 4) {
#pragma unroll
                for (int q = 0; q < 4; ++q) v += part[(size_t)q * FC_ROWS * (OQ * 4) + i];
            }
#pragma unroll
            for (int u = 0; u < FC_TG; ++u) if (u == tt) res[pi][u] = v;
        }
    }
    // each (row, o) pair now owns FC_TG consecutive frames: contiguous 32-byte runs in y [R][O][Tp]
#pragma unroll
    for (int pi = 0; pi < NP; ++pi) {
        int i = threadIdx.x + pi * 256;
        if (i >= FC_ROWS * OQ * 4) continue;
        int rr = i / (OQ * 4), o = i - rr * (OQ * 4);
        if (o >= O || row0 + rr >= R) continue;
        float* dst = y + ((size_t)(row0 + rr) * O + o) * Tp + t0;
        const float bias = fc_b[o];
#pragma unroll
        for (int u = 0; u < FC_TG; ++u)
            if (t0 + u < Tp) dst[u] = res[pi][u] + bias;
    }
}

template <int OQ>
int launch_fc(const __half* hseq, int R, int RS, int Tp, const float* w, const float* b, int O, float* y, cudaStream_t s) {
    size_t fsm = (size_t)FC_ROWS * (H + 8) * 2 + (size_t)H * OQ * 16 + (size_t)4 * FC_ROWS * OQ * 16;
    NPPC_CUDA_OK(cudaFuncSetAttribute(lstm_fc_kernel<OQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
    lstm_fc_kernel<OQ><<<dim3(nppc::cdiv(R, FC_ROWS), nppc::cdiv(Tp, FC_TG)), 256, fsm, s>>>(hseq, R, RS, Tp, w, b, O, y);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

struct FcArgs {
    const CUtensorMap* tfc;
    const float* fc_b;
    int O, R;
    float* y;
};

template <int CL, bool FUSE_X, bool FUSE_FC>
int launch_rec_cl(const CUtensorMap& tw, const CUtensorMap& th, const CUtensorMap& thst, const CUtensorMap& tx,
                  const CUtensorMap& twx, const void* zx, const float* bias, int RS, int Tp, const FcArgs& fc,
                  cudaStream_t s) {
    auto kern = lstm_rec_kernel<CL, FUSE_X, FUSE_FC>;
    NPPC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, RecSmem::TOTAL));
    int tiles = RS / ROWS;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nppc::cdiv(tiles, CL) * CL));  // padded CTAs keep the cluster protocol, store nothing
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = RecSmem::TOTAL;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NPPC_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tw, th, thst, tx, twx, *fc.tfc, (const uint4*)zx, bias, RS, Tp, fc.fc_b, fc.O,
                                    fc.R, fc.y));
    NPPC_COUNT_LAUNCH(1);
    return NPPC_OK;
}

int rec_cluster_size() {
    static int cl = -1;
    if (cl < 0) {
        const char* e = getenv("NPPC_LSTM_CLUSTER");
        cl = e ? atoi(e) : 2;
        if (cl != 1 && cl != 2 && cl != 4) cl = 2;
    }
    return cl;
}

template <bool FUSE_X, bool FUSE_FC>
int launch_rec(const CUtensorMap& tw, const CUtensorMap& th, const CUtensorMap& thst, const CUtensorMap& tx,
               const CUtensorMap& twx, const void* zx, const float* bias, int RS, int Tp, const FcArgs& fc, cudaStream_t s) {
    switch (rec_cluster_size()) {
        case 1: return launch_rec_cl<1, FUSE_X, FUSE_FC>(tw, th, thst, tx, twx, zx, bias, RS, Tp, fc, s);
        case 2: return launch_rec_cl<2, FUSE_X, FUSE_FC>(tw, th, thst, tx, twx, zx, bias, RS, Tp, fc, s);
        default: return launch_rec_cl<4, FUSE_X, FUSE_FC>(tw, th, thst, tx, twx, zx, bias, RS, Tp, fc, s);
    }
}

}  // namespace

namespace nppc {

size_t lstm_workspace_tc(const nppc_lstm_plan* p, int R, int Tp) {
    (void)p;
    size_t rows = (size_t)Tp * (size_t)(cdiv(R, ROWS) * ROWS);
    return rows * H4 * 2 + rows * H * 2 + 1024;
}

int lstm_plan_pack_tc(nppc_lstm_plan* p, const float* w_ih0, const float* w_hh0, const float* w_ih1, const float* w_hh1,
                      cudaStream_t s) {
    if (p->H != H) return NPPC_OK;  // tensor-core path is built for H = 384 only; impl 0 covers other sizes
    p->KP0 = ((p->I + 63) / 64) * 64;
    const float* wih[2] = {w_ih0, w_ih1};
    const float* whh[2] = {w_hh0, w_hh1};
    for (int l = 0; l < 2; ++l) {
        int Kin = l == 0 ? p->I : H, KP = l == 0 ? p->KP0 : H;
        NPPC_CUDA_OK(cudaMalloc(&p->wp_ih[l], sizeof(__nv_bfloat16) * (size_t)H4 * KP));
        NPPC_CUDA_OK(cudaMalloc(&p->wp_hh[l], sizeof(__nv_bfloat16) * (size_t)H4 * H));
        NPPC_CUDA_OK(cudaMalloc(&p->bias_p[l], sizeof(float) * H4));
        pack_w_kernel<<<256, 256, 0, s>>>(wih[l], Kin, KP, (__half*)p->wp_ih[l]);
        pack_w_kernel<<<256, 256, 0, s>>>(whh[l], H, H, (__half*)p->wp_hh[l]);
        pack_b_kernel<<<cdiv(H4, 256), 256, 0, s>>>(p->bias[l], p->bias_p[l]);
    }
    NPPC_CUDA_OK(cudaMalloc(&p->wp_fc, sizeof(__half) * 16 * H));
    pack_fc_kernel<<<cdiv(16 * H, 256), 256, 0, s>>>(p->fc_w, p->O, (__half*)p->wp_fc);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

void lstm_plan_free_tc(nppc_lstm_plan* p) {
    cudaFree(p->wp_fc);
    for (int l = 0; l < 2; ++l) {
        cudaFree(p->wp_ih[l]); cudaFree(p->wp_hh[l]); cudaFree(p->bias_p[l]);
    }
}

int lstm_forward_tc(const nppc_lstm_plan* p, const void* xs, int R, int RS, int Tp, int KP, void* ws, size_t ws_bytes,
                    float* y, cudaStream_t s) {
    NPPC_CHECK_ARG(p->H == H && p->wp_hh[0], "nppc_lstm_forward(impl 1): the tcgen05 path is built for H=384 (got %d)", p->H);
    NPPC_CHECK_ARG(KP == p->KP0, "nppc_lstm_forward(impl 1): KP must be %d (got %d)", p->KP0, KP);
    NPPC_CHECK_ARG(p->O <= OPMAX, "nppc_lstm_forward(impl 1): output size %d > %d", p->O, OPMAX);
    NPPC_CHECK_ARG(RS == cdiv(R, ROWS) * ROWS, "nppc_lstm_forward(impl 1): R_stride must be R rounded up to %d (got %d for R=%d)",
                   ROWS, RS, R);
    NPPC_CHECK_ARG(ws_bytes >= lstm_workspace_tc(p, R, Tp), "nppc_lstm_forward: workspace too small");
    NPPC_CHECK_ARG((long long)Tp * RS < (1LL << 31), "nppc_lstm_forward(impl 1): T'*R too large");
    const size_t rows = (size_t)Tp * RS;
    uintptr_t base = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
    __nv_bfloat16* zx = (__nv_bfloat16*)base;
    __nv_bfloat16* hseq = zx + rows * H4;
    const long long M = (long long)rows;
    CUtensorMap tw[2], th, thst;
    for (int l = 0; l < 2; ++l) {
        int rc = tc::make_tmap_bf16_2d(&tw[l], p->wp_hh[l], H4, H, H * 2, 128, 64);
        if (rc) return rc;
    }
    int rc = tc::make_tmap_bf16_2d(&th, hseq, (uint64_t)M, H, H * 2, ROWS, 64);
    if (rc) return rc;
    rc = tc::make_tmap_bf16_2d(&thst, hseq, (uint64_t)M, H, H * 2, ROWS, CH, 0);
    if (rc) return rc;
    // layer 0: input projection (K = 64) fused into the recurrent kernel unless NPPC_LSTM_FUSE_X=0
    static const bool fuse_x = !(getenv("NPPC_LSTM_FUSE_X") && atoi(getenv("NPPC_LSTM_FUSE_X")) == 0);
    CUtensorMap tx, twx;
    rc = tc::make_tmap_bf16_2d(&tx, xs, (uint64_t)M, (uint64_t)KP, (uint64_t)KP * 2, ROWS, 64);
    if (rc) return rc;
    rc = tc::make_tmap_bf16_2d(&twx, p->wp_ih[0], H4, (uint64_t)KP, (uint64_t)KP * 2, 128, 64);
    if (rc) return rc;
    static const bool fuse_fc = !(getenv("NPPC_LSTM_FUSE_FC") && atoi(getenv("NPPC_LSTM_FUSE_FC")) == 0);
    CUtensorMap tfc;
    rc = tc::make_tmap_bf16_2d(&tfc, p->wp_fc, 16, H, H * 2, 16, 64);
    if (rc) return rc;
    const FcArgs no_fc{&tfc, nullptr, 0, R, nullptr};
    if (fuse_x && KP == 64) {
        rc = launch_rec<true, false>(tw[0], th, thst, tx, twx, nullptr, p->bias_p[0], RS, Tp, no_fc, s);
    } else {
        rc = gemm_16bit_tn(xs, p->wp_ih[0], p->bias_p[0], zx, M, H4, KP, 2, s);
        if (rc) return rc;
        rc = launch_rec<false, false>(tw[0], th, thst, tx, twx, zx, nullptr, RS, Tp, no_fc, s);
    }
    if (rc) return rc;
    // layer 1 (+ fc)
    rc = gemm_16bit_tn(hseq, p->wp_ih[1], p->bias_p[1], zx, M, H4, H, 2, s);
    if (rc) return rc;
    if (fuse_fc && p->O <= 16) {   // fc fused into the last layer's recurrent kernel (13th mini-chunk per step)
        const FcArgs fc{&tfc, p->fc_b, p->O, R, y};
        return launch_rec<false, true>(tw[1], th, thst, tx, twx, zx, nullptr, RS, Tp, fc, s);
    }
    rc = launch_rec<false, false>(tw[1], th, thst, tx, twx, zx, nullptr, RS, Tp, no_fc, s);
    if (rc) return rc;
    switch ((p->O + 3) / 4) {
        case 1: return launch_fc<1>((const __half*)hseq, R, RS, Tp, p->fc_w, p->fc_b, p->O, y, s);
        case 2: return launch_fc<2>((const __half*)hseq, R, RS, Tp, p->fc_w, p->fc_b, p->O, y, s);
        case 3: return launch_fc<3>((const __half*)hseq, R, RS, Tp, p->fc_w, p->fc_b, p->O, y, s);
        case 4: return launch_fc<4>((const __half*)hseq, R, RS, Tp, p->fc_w, p->fc_b, p->O, y, s);
        case 5: return launch_fc<5>((const __half*)hseq, R, RS, Tp, p->fc_w, p->fc_b, p->O, y, s);
        case 6: return launch_fc<6>((const __half*)hseq, R, RS, Tp, p->fc_w, p->fc_b, p->O, y, s);
        default: break;
    }
    set_error("nppc_lstm_forward(impl 1): output size %d > 24 not built", p->O);
    return NPPC_ERR_UNSUPPORTED;
}

}  // namespace nppc


#ifdef NPPC_REC_TRACE
extern "C" int nppc_debug_rec_trace(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, g_trace, sizeof(g_trace)) == cudaSuccess ? 0 : -2;
}
#endif
