// LSTM layer-1 input projection for ALL steps:  Zx[T'*R, 4H] = H0[T'*R, H] * W_ih^T + (b_ih + b_hh), fp16 in / fp16 out,
// H = 384, written in the pre-activation layout the recurrent kernel's epilogue reads (lstm_tc.cu).
// (the W_ih half of nn.LSTM's gate GEMM, sequence_model.py:118)
//
// CTA-pair GEMM (tcgen05.mma.cta_group::2, M = 256 per pair, N = 256 per tile):
//   * the pair's A operand (its 2 x 128 rows x K = 384, 96 KB per CTA) is loaded ONCE per m-block and stays resident in
//     shared memory for all six 256-column n-tiles; each CTA streams only ITS half of every weight tile (128 of 256
//     columns) through a 7-stage TMA ring.  The single-CTA 128 x 256 GEMM (gemm_tc.cu) moves 192 B/clk through shared
//     memory at full tensor rate (operand reads + TMA writes of both operands + the C staging) and is bound by the
//     128 B/clk port; this kernel needs ~94 B/clk.
//   * two TMEM accumulator stages (2 x 256 columns): the epilogue of tile i overlaps the MMAs of tile i+1.
//   * epilogue: 8 warps, warp = (lane quadrant, 128-column chunk).  In the Zx layout
//       uint4 index ((((m_blk*12 + chunk)*4 + quad)*2 + half)*8 + piece)*32 + lane
//     a warp-wide 16-byte store is 512 contiguous bytes, so the tile goes registers -> global directly: no shared-memory
//     staging, no TMA store.  TMEM loads are double-buffered against the convert + store of the previous 32 columns.
//   * A slabs are released one by one during the m-block's last n-tile, so the next m-block's A streams in behind the MMAs.
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace {
using namespace nppc::tc;

constexpr int ZK = 384, ZN = 1536;
constexpr int NSL = ZK / 64;          // 6 K slabs
constexpr int NTILE = ZN / 256;       // 6 n-tiles
constexpr int SLAB = 128 * 64 * 2;    // 16 KB: [128 rows][64 k] fp16, SW128 (A slab and B stage have the same shape)
constexpr int NST = 7;
constexpr int NTHREADS = 384;         // warps 0-3: B producer, MMA, TMEM alloc, A producer ; warps 4-11: epilogue

struct ZxSmem {
    static constexpr int A_OFF = 0;
    static constexpr int B_OFF = NSL * SLAB;
    static constexpr int BIAS_OFF = B_OFF + NST * SLAB;
    static constexpr int BAR_OFF = BIAS_OFF + ZN * 4;
    static constexpr int TOTAL = BAR_OFF + 512 + 1024;
    static_assert(TOTAL <= 227 * 1024, "shared memory budget exceeded");
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
gemm_zx_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const float* __restrict__ bias, uint4* __restrict__ zx, int m_blocks) {
    using S = ZxSmem;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* bias_s = reinterpret_cast<float*>(smem + S::BIAS_OFF);
    uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);   // leader: both CTAs' weight halves have landed
    uint64_t* w_empty = w_full + NST;                                    // both: the MMAs reading ring slot s are done
    uint64_t* a_full = w_empty + NST;                                    // leader [NSL]
    uint64_t* a_free = a_full + NSL;                                     // both [NSL]: last MMA reading A slab k is done
    uint64_t* tfull = a_free + NSL;                                      // both [2]
    uint64_t* tempty = tfull + 2;                                        // leader [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const int n_pairs = gridDim.x >> 1, pair0 = blockIdx.x >> 1;
    const int num_mp = (m_blocks + 1) >> 1;   // m-block pairs

    for (int i = threadIdx.x; i < ZN; i += NTHREADS) bias_s[i] = bias ? bias[i] : 0.f;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&w_full[i], 2); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < NSL; ++i) { mbar_init(&a_full[i], 2); mbar_init(&a_free[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_pair<512>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ---- B producer (both CTAs): this CTA's 128 columns of every 256-column weight tile, one K slab per stage ----
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t wf0 = mapa_u32(smem_u32(w_full), 0);
            for (int mp = pair0; mp < num_mp; mp += n_pairs)
                for (int nt = 0; nt < NTILE; ++nt)
                    for (int k = 0; k < NSL; ++k) {
                        mbar_wait(&w_empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx_cluster(wf0 + stage * 8, SLAB);
                        tma_load_2d_pair(smem + S::B_OFF + stage * SLAB, &tmap_b, wf0 + stage * 8, k * 64,
                                         nt * 256 + (int)crank * 128);
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 3) {
        // ---- A producer (both CTAs): this CTA's 128 rows, slab by slab as the previous m-block releases them ----
        if (lane == 0) {
            const uint32_t af0 = mapa_u32(smem_u32(a_full), 0);
            int it = 0;
            for (int mp = pair0; mp < num_mp; mp += n_pairs, ++it) {
                int m_blk = 2 * mp + (int)crank;
                if (m_blk >= m_blocks) m_blk = m_blocks - 1;   // padding CTA of the last pair: valid loads, no stores
                for (int k = 0; k < NSL; ++k) {
                    if (it > 0) mbar_wait(&a_free[k], (it - 1) & 1);
                    mbar_arrive_expect_tx_cluster(af0 + k * 8, SLAB);
                    tma_load_2d_pair(smem + S::A_OFF + k * SLAB, &tmap_a, af0 + k * 8, k * 64, m_blk * 128);
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer (leader CTA only) ----
        if (crank == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(256, 256);
            int stage = 0; uint32_t phase = 0;
            uint32_t tc = 0;
            int it = 0;
            const uint32_t a_base = smem_u32(smem + S::A_OFF), b_base = smem_u32(smem + S::B_OFF);
            for (int mp = pair0; mp < num_mp; mp += n_pairs, ++it)
                for (int nt = 0; nt < NTILE; ++nt, ++tc) {
                    const uint32_t as = tc & 1;
                    mbar_wait(&tempty[as], ((tc >> 1) & 1) ^ 1);
                    tcgen05_fence_after();
                    for (int k = 0; k < NSL; ++k) {
                        if (nt == 0) mbar_wait(&a_full[k], it & 1);
                        mbar_wait(&w_full[stage], phase);
                        tcgen05_fence_after();
                        const uint64_t da = umma_desc_k128(a_base + k * SLAB);
                        const uint64_t db = umma_desc_k128(b_base + stage * SLAB);
                        if (elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                umma_f16_pair(tmem_base + as * 256, da + 2 * kk, db + 2 * kk, idesc, (k | kk) != 0);
                            umma_commit_pair(&w_empty[stage], 3);
                            if (nt == NTILE - 1) umma_commit_pair(&a_free[k], 3);
                            if (k == NSL - 1) umma_commit_pair(&tfull[as], 3);
                        }
                        __syncwarp();
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                }
        }
    } else if (warp >= 4) {
        // ---- epilogue (both CTAs): warp = (lane quadrant ew, 128-column chunk ch of the tile) ----
        const int ew = warp & 3, ch = (warp - 4) >> 2;
        const uint32_t te0 = mapa_u32(smem_u32(tempty), 0);
        uint32_t tc = 0;
        for (int mp = pair0; mp < num_mp; mp += n_pairs) {
            const int m_blk = 2 * mp + (int)crank;
            const bool live = m_blk < m_blocks;
            for (int nt = 0; nt < NTILE; ++nt, ++tc) {
                const uint32_t as = tc & 1;
                mbar_wait(&tfull[as], (tc >> 1) & 1);
                tcgen05_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + as * 256 + ch * 128;
                const float* bs = bias_s + nt * 256 + ch * 128;
                // destination of this warp's [32 rows][128 cols]: 16 pieces of 512 contiguous bytes
                uint4* dst = zx + ((((size_t)(live ? m_blk : 0) * (ZN / 128) + nt * 2 + ch) * 4 + ew) * 2) * 256 + lane;
                uint32_t va[32], vb[32];
                auto flush = [&](const uint32_t (&v)[32], int c) {   // columns 32c .. 32c+31 -> pieces (half = c>>1, q = 4*(c&1)..)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t pk[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float a = __uint_as_float(v[8 * q + 2 * i]) + bs[c * 32 + 8 * q + 2 * i];
                            const float b = __uint_as_float(v[8 * q + 2 * i + 1]) + bs[c * 32 + 8 * q + 2 * i + 1];
                            // saturating: fp16 has the mantissa the LSTM pre-activations need, not the range
                            __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
                            pk[i] = *reinterpret_cast<uint32_t*>(&h);
                        }
                        if (live) dst[((c >> 1) * 8 + (c & 1) * 4 + q) * 32] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                };
                auto wait32 = [&](uint32_t (&v)[32]) {
                    asm volatile("tcgen05.wait::ld.sync.aligned;"
                                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                                 :: "memory");
                };
                tmem_ld32(t_row, va);
                wait32(va);
                tmem_ld32(t_row + 32, vb);
                flush(va, 0);
                wait32(vb);
                tmem_ld32(t_row + 64, va);
                flush(vb, 1);
                wait32(va);
                tmem_ld32(t_row + 96, vb);
                flush(va, 2);
                wait32(vb);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(te0 + as * 8);   // accumulator stage is in registers
                flush(vb, 3);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc_pair<512>(tmem_base);
    }
}

}  // namespace

namespace nppc {
// A [M][384] fp16 (M % 128 == 0), W [1536][384] fp16 (packed gate-column order), bias [1536] f32 -> zx in the LSTM layout
int gemm_zx_pair(const void* A, const void* W, const float* bias, void* zx, long long M, cudaStream_t s) {
    NPPC_CHECK_ARG(A && W && zx && M > 0 && M % 128 == 0, "gemm_zx_pair: need M %% 128 == 0 (M=%lld)", M);
    NPPC_CHECK_ARG((M / 128) * (ZN / 128) * 256 < (1LL << 31), "gemm_zx_pair: M too large");
    CUtensorMap ta, tb;
    int rc = tc::make_tmap_bf16_2d(&ta, A, (uint64_t)M, ZK, ZK * 2, 128, 64);
    if (rc) return rc;
    rc = tc::make_tmap_bf16_2d(&tb, W, ZN, ZK, ZK * 2, 128, 64);
    if (rc) return rc;
    NPPC_CUDA_OK(cudaFuncSetAttribute(gemm_zx_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ZxSmem::TOTAL));
    const int m_blocks = (int)(M / 128);
    int pairs = (m_blocks + 1) / 2;
    const int max_pairs = sm_count() / 2;
    if (pairs > max_pairs) pairs = max_pairs;
    gemm_zx_pair_kernel<<<pairs * 2, NTHREADS, ZxSmem::TOTAL, s>>>(ta, tb, bias, (uint4*)zx, m_blocks);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}
}  // namespace nppc
