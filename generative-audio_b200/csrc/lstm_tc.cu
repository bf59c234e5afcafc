// a7, implementation 1: fp16 tcgen05 tensor-core LSTM (2 layers + fc), H = 384.
//
// Layer 0:  ONE persistent recurrent kernel; the K = 64 input projection is fused (x_t is a 7th A slab).
// Layer 1:  input projection for ALL steps as one big GEMM (gemm_tc.cu): Zx[T'*R, 4H] = H0 * W_ih^T + (b_ih+b_hh), fp16,
//           then the persistent recurrent kernel, which also computes the fc output layer (16-column mini-chunk per step).
// The recurrent kernel (CTA pair, cta_group::2) is described at lstm_rec_kernel below.
#include <stdlib.h>
#include <string.h>
#include <cuda_fp16.h>
#include "lstm_plan.cuh"
#include "tc_common.cuh"

namespace nppc {
int gemm_16bit_tn(const void* A, const void* W, const float* bias, void* C, long long M, int N, int K, int f16,
                  cudaStream_t s);
int gemm_zx_pair(const void* A, const void* W, const float* bias, void* zx, long long M, cudaStream_t s);
}

namespace {
using namespace nppc::tc;

constexpr int H = 384;
constexpr int H4 = 4 * H;
constexpr int ROWS = 128;           // sequences per CTA
constexpr int CH = 32;              // hidden units per chunk
constexpr int NCHUNK = H / CH;      // 12
constexpr int NSLAB = H / 64;       // 6 K-slabs of 64
constexpr int SLAB_BYTES = ROWS * 64 * 2;   // 16 KB (A slab and W stage have the same shape: 128 x 64 fp16)
constexpr int NTHREADS = 384;       // warps 0-3: TMA-W, MMA, TMEM alloc, TMA-A ; warps 4-11: epilogue
constexpr int OPMAX = 24;

constexpr int HST_BYTES = ROWS * CH * 2;   // one chunk of h_t for the CTA's rows: [128][32] fp16 = 8 KB

#ifdef NPPC_LSTM_ABLATE
#define NPPC_DBG(dbg, bit) ((dbg) & (bit))
#else
#define NPPC_DBG(dbg, bit) false
#endif

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
// sigmoid(z) = 0.5 tanh(z / 2) + 0.5; the 1/2 is folded into the packed i, f, o rows of W_ih, W_hh and the bias (exact:
// a power-of-two scale), so the MMAs already deliver z / 2 for those gates
__device__ __forceinline__ float sigmoid_half_arg(float hx) { return fmaf(0.5f, tanh_fast(hx), 0.5f); }

// tcgen05.wait::ld that also names the destination registers, so no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}

// =====================================================================================================================
// CTA-pair persistent recurrent kernel.  Two CTAs (cluster of 2, ranks 0/1) own 2 x 128 sequences for all T' steps and
// run ONE tcgen05.mma.cta_group::2 of M = 256 per K-step: each CTA supplies its own 128 rows of h_{t-1} (A, shared
// memory, K-major SW128) and only HALF of the weight tile (64 of the chunk's 128 gate columns), so per SM the weight
// stream through shared memory (TMA write + operand read) is halved.
//   per step t, per chunk (32 hidden units = 128 gate columns, 12 chunks, walked in a per-pair rotated order):
//     tensor core : acc[stage][256 x 128] = h_{t-1} * W_hh[chunk]^T  (+ x_t * W_ih[chunk]^T when the K=64 layer-0 input
//                   projection is fused, + a 16-column fc mini-chunk per step for the last layer)
//     epilogue    : 8 warps, thread = (row, 16-unit half).  The half's 64 accumulator columns are ordered
//                   [4-unit block][gate][unit] so that one 16-column tcgen05.ld brings i,f,g,o of 4 units: the load of
//                   block b+1 is in flight while block b goes through the MUFU (tanh) pipe — TMEM reads (64 B/clk/SM)
//                   and the 5 tanh per unit are the two floors of this kernel and now overlap.
//                   c lives in TMEM as packed fp16 pairs (192 columns; measured cost at the LSTM output 6.3e-4 ->
//                   7.1e-4 max-rel vs fp32), which leaves room for the second accumulator stage.
//   TMEM columns: [0,128) acc stage 0 | [128,256) acc stage 1 | [256,448) c (fp16x2) | [448,464) fc accumulator.
//   mbarriers   : w_full / a_full / x_full / acc_empty / fc_empty live in the LEADER (both CTAs' producers and epilogues
//                 arrive remotely, CTA-scope semantics: a cluster-scope release costs ~700 cycles per arrive);
//                 w_empty / acc_full / fc_full / x_free are signalled in both CTAs by multicast tcgen05.commit.
// Packed gate column (weights, bias, Zx):  chunk*128 + half*64 + blk*16 + gate*4 + uu  <->  nn.LSTM row
//                 gate*H + chunk*32 + half*16 + blk*4 + uu.
// Zx (layer >= 1 pre-activations, fp16) is read in the layout the GEMM epilogue writes:
//   uint4 index ((((t*tiles + tile)*12 + chunk)*4 + warp-quad)*2 + half)*256 + piece*32 + lane, piece q = columns 8q..8q+7
//   of the half -> block b = pieces 2b (i,f) and 2b+1 (g,o); every warp-wide load reads 512 contiguous bytes.
template <bool FUSE_X>
struct RecSmem {
    static constexpr int WST_BYTES = 2 * 64 * 64 * 2;                    // ring stage: this CTA's half of TWO K slabs (16 KB)
#ifdef NPPC_REC_TRACE
    static constexpr int NST = FUSE_X ? 5 : 6;
#else
    static constexpr int NST = FUSE_X ? 6 : 7;
#endif
    static constexpr int A_OFF = 0;                                      // h_{t-1}: 6 slabs [128][64] fp16, SW128
    static constexpr int W_OFF = NSLAB * SLAB_BYTES;
    static constexpr int HST_OFF = W_OFF + NST * WST_BYTES;              // 2 x [128 rows][32 units] fp16 staging for the TMA store of h_t
    static constexpr int X_OFF = HST_OFF + 2 * HST_BYTES;                // layer 0: x_t [128][64] fp16
    static constexpr int BAR_OFF = X_OFF + (FUSE_X ? SLAB_BYTES : 0);
#ifdef NPPC_REC_TRACE
    static constexpr int TRACE_OFF = BAR_OFF + 512;
    static constexpr int TOTAL = TRACE_OFF + 4 * 12 * 16 * 8 + 1024;
#else
    static constexpr int TOTAL = BAR_OFF + 512 + 1024;
#endif
    static_assert(TOTAL <= 227 * 1024, "shared memory budget exceeded");
    static constexpr int C_COL = 256, FC_COL = 448;
};

template <bool FUSE_X, bool FUSE_FC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
lstm_rec_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
                const __grid_constant__ CUtensorMap tmap_hst, const __grid_constant__ CUtensorMap tmap_x,
                const __grid_constant__ CUtensorMap tmap_wx, const __grid_constant__ CUtensorMap tmap_fc,
                const uint4* __restrict__ zx, const float* __restrict__ bias, int RS, int Tp,
                const float* __restrict__ fc_b, int O, int R, float* __restrict__ y, int dbg, int xkk) {
    using S = RecSmem<FUSE_X>;
    constexpr int NST = S::NST;
    constexpr int NP = NSLAB / 2;                  // slab pairs per chunk (one ring stage each)
    constexpr int NK = NP + (FUSE_X ? 1 : 0);      // ring stages per chunk; with FUSE_X the x slab comes FIRST (k = 0)
    constexpr int HALF_SLAB = 64 * 64 * 2;         // this CTA's [64 cols][64 k] half of one weight slab
    constexpr int FC_HALF = 8 * 64 * 2;            // this CTA's 8 of the 16 fc rows, one K slab
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* w_empty = w_full + NST;
    uint64_t* a_full = w_empty + NST;      // [NSLAB]
    uint64_t* acc_full = a_full + NSLAB;   // [2]
    uint64_t* acc_empty = acc_full + 2;    // [2]
    uint64_t* fc_full = acc_empty + 2;
    uint64_t* fc_empty = fc_full + 1;
    uint64_t* h_stored = fc_empty + 1;     // [NSLAB] CTA-local: h_t slab is in global memory (both chunks, all 8 warps)
    uint64_t* a_free = h_stored + NSLAB;   // every MMA of the step has completed: the A operand may be overwritten
    uint64_t* staged = a_free + 1;         // [2] CTA-local: all 8 epilogue warps have written their h tile into staging buffer b
    uint64_t* stage_free = staged + 2;     // [2] CTA-local: the TMA store has read staging buffer b
    uint64_t* late_written = stage_free + 2;   // CTA-local: the late slab of the A operand has been written by all 8 warps
    uint64_t* x_full = late_written + 1;
    uint64_t* x_free = x_full + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(x_free + 1);
#ifdef NPPC_REC_TRACE
    long long* trace_s = reinterpret_cast<long long*>(smem + S::TRACE_OFF);
    for (int i = threadIdx.x; i < 4 * 12 * 16; i += NTHREADS) trace_s[i] = 0;
#endif

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = RS / ROWS;
    const int tile = blockIdx.x;          // the odd CTA of the last pair may be padding: it redoes the last tile, stores nothing
    const bool live = tile < tiles;
    const int tile_c = live ? tile : tiles - 1;
    const int row0 = tile_c * ROWS;
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;
    // Every pair walks the 12 chunks in its own rotated order (so the pairs, which run in lock-step, do not all fetch the
    // same weight lines at the same moment) and consumes the K slabs in the order the previous step PRODUCED them: the slab
    // finished last (positions 10, 11) is needed last, so the reload of h_t overlaps the tail of step t.
    const int krot = (blockIdx.x >> 1) % NP;
    const int rot = 4 * krot;
    auto chunk_of = [&](int j) { int c = j + rot; return c >= NCHUNK ? c - NCHUNK : c; };
    auto pair_of = [&](int k) { int c = k + krot; return c >= NP ? c - NP : c; };
    // The slab produced LAST in a step (positions 10, 11) skips the global round trip: the epilogue writes it straight into
    // the A operand (its region is free once the accumulator of position 11 has been seen) and arrives on a_full itself.
    const int late_slab = chunk_of(NCHUNK - 1) >> 1;

    for (int i = threadIdx.x; i < NSLAB * SLAB_BYTES / 16; i += NTHREADS)   // h_{-1} = 0
        reinterpret_cast<uint4*>(smem + S::A_OFF)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_h);
        tma_prefetch_desc(&tmap_hst);
        if (FUSE_X) { tma_prefetch_desc(&tmap_x); tma_prefetch_desc(&tmap_wx); }
        if (FUSE_FC) tma_prefetch_desc(&tmap_fc);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&w_full[i], 2); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < NSLAB; ++i) mbar_init(&a_full[i], 2);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 16); }
        mbar_init(fc_full, 1);
        mbar_init(fc_empty, 8);
        for (int i = 0; i < NSLAB; ++i) mbar_init(&h_stored[i], 2);
        for (int i = 0; i < 2; ++i) { mbar_init(&staged[i], 8); mbar_init(&stage_free[i], 1); }
        mbar_init(late_written, 8);
        mbar_init(a_free, 1);
        mbar_init(x_full, 2);
        mbar_init(x_free, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_pair<512>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer's barriers are initialised and its TMEM allocated before any remote arrive / pair MMA
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ---- weight producer (both CTAs): this CTA's 64-column half of two K slabs per TMA, counted on the leader's barrier ----
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t wf0 = mapa_u32(smem_u32(w_full), 0);
            for (int t = 0; t < Tp + (FUSE_FC ? 1 : 0); ++t) {
                if (FUSE_FC && t >= 1)
                    for (int k = 0; k < NP; ++k) {
                        mbar_wait(&w_empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx_cluster(wf0 + stage * 8, 2 * FC_HALF);
                        tma_load_3d_pair(smem + S::W_OFF + stage * S::WST_BYTES, &tmap_fc, wf0 + stage * 8, 0, (int)crank * 8,
                                         pair_of(k) * 2);
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                if (t >= Tp) break;
                for (int j = 0; j < NCHUNK; ++j)
                    for (int k = 0; k < NK; ++k) {
                        mbar_wait(&w_empty[stage], phase ^ 1);
                        if (k == 0) TRACE(9);
                        if (k == NK - 1) TRACE(14);
                        unsigned char* dst = smem + S::W_OFF + stage * S::WST_BYTES;
                        const int wrow = chunk_of(j) * 128 + (int)crank * 64;
                        if (NPPC_DBG(dbg, 2)) {
                            mbar_arrive_cluster(wf0 + stage * 8);
                        } else if (FUSE_X && k == 0) {
                            mbar_arrive_expect_tx_cluster(wf0 + stage * 8, HALF_SLAB);
                            tma_load_2d_pair(dst, &tmap_wx, wf0 + stage * 8, 0, wrow);
                        } else {
                            mbar_arrive_expect_tx_cluster(wf0 + stage * 8, 2 * HALF_SLAB);
                            tma_load_3d_pair(dst, &tmap_w, wf0 + stage * 8, 0, wrow, pair_of(FUSE_X ? k - 1 : k) * 2);
                        }
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 3) {
        // ---- A producer (both CTAs): reload this CTA's h_t rows as the A operand of step t+1 ----
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        if (lane == 0) {
            const uint32_t af0 = mapa_u32(smem_u32(a_full), 0);
            const uint32_t xf0 = mapa_u32(smem_u32(x_full), 0);
            for (int k = 0; k < NSLAB; ++k) mbar_arrive_cluster(af0 + k * 8);  // step 0: zeros already in place
            if (FUSE_X) {
                mbar_arrive_expect_tx_cluster(xf0, SLAB_BYTES);
                tma_load_2d_pair(smem + S::X_OFF, &tmap_x, xf0, 0, row0);
            }
            for (int t = 1; t < Tp + (FUSE_FC ? 1 : 0); ++t) {
                if (FUSE_X && t < Tp) {   // x_t: needs only the previous step's x MMAs to be done (independent of h)
                    mbar_wait(x_free, (t - 1) & 1);
                    mbar_arrive_expect_tx_cluster(xf0, SLAB_BYTES);
                    tma_load_2d_pair(smem + S::X_OFF, &tmap_x, xf0, 0, t * RS + row0);
                }
                mbar_wait(a_free, (t - 1) & 1);
                for (int kk = 0; kk < NSLAB; ++kk) {   // in the order the slabs were produced = the order they are consumed
                    const int k = pair_of(kk >> 1) * 2 + (kk & 1);
                    if (k == late_slab) continue;
                    mbar_wait(&h_stored[k], (t - 1) & 1);
                    if (NPPC_DBG(dbg, 8)) { mbar_arrive_cluster(af0 + k * 8); continue; }
                    mbar_arrive_expect_tx_cluster(af0 + k * 8, SLAB_BYTES);
                    tma_load_2d_pair(smem + S::A_OFF + k * SLAB_BYTES, &tmap_h, af0 + k * 8, k * 64, (t - 1) * RS + row0);
                }
            }
        }
    } else if (warp == 2) {
        // ---- store warp (both CTAs): everything slow about getting h_t out of the SM lives here, off the epilogue warps'
        //      (Proxy ordering: the epilogue warps' generic st.shared are released by their mbarrier arrive, acquired by this
        //      thread's wait, and only then does THIS thread's fence.proxy.async order them before the TMA store / the MMA
        //      that read the same bytes through the async proxy.  tests: bitwise determinism over many CTA pairs.)
        //      critical path: the generic->async proxy fence (a MEMBAR that in an epilogue warp would also wait for its
        //      in-flight Zx loads), the TMA store of the staged [128 x 32] tile, its read / write completion, and the
        //      hand-over of the late slab to the MMA issuer.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        if (lane == 0) {
            const uint32_t al0 = mapa_u32(smem_u32(&a_full[late_slab]), 0);
            unsigned char* hst = smem + S::HST_OFF;
            for (int t = 0; t < Tp; ++t)
                for (int j = 0; j < NCHUNK; ++j) {
                    const uint32_t b = j & 1, use = (uint32_t)t * (NCHUNK / 2) + (j >> 1);
                    if (j == NCHUNK - 1) {   // positions 10, 11 are in the A operand: publish them to the tensor core first
                        mbar_wait(late_written, t & 1);
                        fence_proxy_async_smem();
                        mbar_arrive_cluster(al0);
                    }
                    mbar_wait(&staged[b], use & 1);
                    fence_proxy_async_smem();
                    if (live) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmap_hst),
                                     "r"(smem_u32(hst + b * HST_BYTES)), "r"(chunk_of(j) * CH), "r"(t * RS + row0)
                                     : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    mbar_arrive(&stage_free[b]);
                    // completion of the PREVIOUS position's store (a chunk period old: never stalls long); async-proxy write
                    // -> async-proxy read by the reload, so no proxy fence is needed
                    if (j >= 1) {
                        asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
                        mbar_arrive(&h_stored[chunk_of(j - 1) >> 1]);
                    }
                    if (j == NCHUNK - 1) {
                        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                        mbar_arrive(&h_stored[late_slab]);
                    }
                }
        }
    } else if (warp == 1) {
        // ---- MMA issuer (leader CTA only): the warp runs the warp-uniform control flow, one elected lane issues ----
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        if (leader) {
            constexpr uint32_t idesc = umma_idesc_f16(2 * ROWS, 128);
            constexpr uint32_t idesc_fc = umma_idesc_f16(2 * ROWS, 16);
            int stage = 0; uint32_t phase = 0;
            const uint32_t a_base = smem_u32(smem + S::A_OFF);
            const uint32_t w_base = smem_u32(smem + S::W_OFF);
            const uint32_t x_base = smem_u32(smem + S::X_OFF);
            for (int t = 0; t < Tp + (FUSE_FC ? 1 : 0); ++t) {
                if (FUSE_FC && t >= 1) {   // y_{t-1}: [256 x 16] = h_{t-1} * W_fc^T into the fc accumulator columns
                    mbar_wait(fc_empty, (t & 1));   // (t-1)-th use: parity ((t-1) & 1) ^ 1
                    tcgen05_fence_after();
                    for (int k = 0; k < NP; ++k) {
                        const int kp = pair_of(k);
                        mbar_wait(&a_full[2 * kp], t & 1);
                        mbar_wait(&a_full[2 * kp + 1], t & 1);
                        mbar_wait(&w_full[stage], phase);
                        tcgen05_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int s2 = 0; s2 < 2; ++s2) {
                                const uint64_t da = umma_desc_k128(a_base + (2 * kp + s2) * SLAB_BYTES);
                                const uint64_t db = umma_desc_k128(w_base + stage * S::WST_BYTES + s2 * FC_HALF);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    umma_f16_pair(tmem_base + S::FC_COL, da + 2 * kk, db + 2 * kk, idesc_fc, (k | s2 | kk) != 0);
                            }
                            umma_commit_pair(&w_empty[stage], 3);
                            if (k == NP - 1) umma_commit_pair(fc_full, 3);
                        }
                        __syncwarp();
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                }
                if (t >= Tp) break;
                for (int j = 0; j < NCHUNK; ++j) {
                    const uint32_t as = j & 1, use = (uint32_t)t * (NCHUNK / 2) + (j >> 1);
                    mbar_wait(&acc_empty[as], (use & 1) ^ 1);
                    TRACE(0);
                    tcgen05_fence_after();
                    for (int k = 0; k < NK; ++k) {
                        const bool is_x = FUSE_X && k == 0;
                        const int kp = pair_of(FUSE_X ? k - 1 : k);
                        if (j == 0) {
                            if (is_x) mbar_wait(x_full, t & 1);
                            else { mbar_wait(&a_full[2 * kp], t & 1); mbar_wait(&a_full[2 * kp + 1], t & 1); }
                        }
                        mbar_wait(&w_full[stage], phase);
                        if (k == 0) TRACE(1);
                        if (k == NK - 1) TRACE(3);
                        tcgen05_fence_after();
                        if (elect_one()) {
                            if (!NPPC_DBG(dbg, 4)) {
                                if (is_x) {
                                    const uint64_t da = umma_desc_k128(x_base);
                                    const uint64_t db = umma_desc_k128(w_base + stage * S::WST_BYTES);
                                    for (int kk = 0; kk < xkk; ++kk)   // only the K = 16 steps that hold real input features
                                        umma_f16_pair(tmem_base + as * 128, da + 2 * kk, db + 2 * kk, idesc, kk != 0);
                                } else {
#pragma unroll
                                    for (int s2 = 0; s2 < 2; ++s2) {
                                        const uint64_t da = umma_desc_k128(a_base + (2 * kp + s2) * SLAB_BYTES);
                                        const uint64_t db = umma_desc_k128(w_base + stage * S::WST_BYTES + s2 * HALF_SLAB);
#pragma unroll
                                        for (int kk = 0; kk < 4; ++kk)
                                            umma_f16_pair(tmem_base + as * 128, da + 2 * kk, db + 2 * kk, idesc, (k | s2 | kk) != 0);
                                    }
                                }
                            }
                            umma_commit_pair(&w_empty[stage], 3);
                            if (k == NK - 1) umma_commit_pair(&acc_full[as], 3);
                            if (is_x && j == NCHUNK - 1) umma_commit_pair(x_free, 3);   // x_t consumed by every chunk of this step
                            if (k == NK - 1 && j == NCHUNK - 1) umma_commit_pair(a_free, 3);
                        }
                        __syncwarp();
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                    TRACE(4);
                }
            }
        }
    } else if (warp >= 4) {
        // ---- epilogue (both CTAs): thread = (row = 32*(warp%4) + lane, half = (warp-4)/4), software-pipelined over positions:
        //      the four 16-column TMEM loads of position p+1 are issued BEFORE h of position p is staged / stored, there is
        //      ONE tcgen05.wait::ld per position (after which the accumulator stage is released at once) and the Zx registers
        //      are refilled in place for the next position as soon as a block has consumed them.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const int ew = warp & 3, half = (warp - 4) >> 2;
        const int rloc = ew * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(ew * 32) << 16);
        unsigned char* hst = smem + S::HST_OFF;
        const uint32_t ae0 = mapa_u32(smem_u32(acc_empty), 0);
        const uint32_t fe0 = mapa_u32(smem_u32(fc_empty), 0);
        // this thread's 32 bytes of a late-slab row, position parity 0 / 1 (SW128: 16-byte chunk index ^ (row & 7))
        unsigned char* late_row = smem + S::A_OFF + late_slab * SLAB_BYTES + rloc * 128;
        const int lsw = rloc & 7;
        {
            uint32_t z[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = 0u;
            for (int j = 0; j < NCHUNK; ++j) tmem_st8(t_lane + S::C_COL + j * 16 + half * 8, z);   // c_0 = 0
            tmem_wait_st();
        }
        auto zx_ptr = [&](int t, int jc) -> const uint4* {
            return zx + (((((size_t)t * tiles + tile_c) * NCHUNK + jc) * 4 + ew) * 2 + half) * 256 + lane;
        };
        const bool valid = live && (row0 + rloc) < R;
        // h_t of position j: this thread's 16 units -> the CTA's staging tile [128 rows][32 units]; the store warp does the rest
        auto stage_out = [&](const uint32_t (&hv)[8], int t, int j) {
            const uint32_t b = j & 1, use = (uint32_t)t * (NCHUNK / 2) + (j >> 1);
            mbar_wait(&stage_free[b], (use & 1) ^ 1);
            uint4* dst = reinterpret_cast<uint4*>(hst + b * HST_BYTES + rloc * (CH * 2) + half * 32);
            dst[0] = make_uint4(hv[0], hv[1], hv[2], hv[3]);
            dst[1] = make_uint4(hv[4], hv[5], hv[6], hv[7]);
            __syncwarp();
            if (lane == 0) mbar_arrive(&staged[b]);
        };
        auto fc_out = [&](int t) {   // fc mini-chunk of step t-1: y[row][o][t-1]; the half-0 warps own the rows
            mbar_wait(fc_full, (t - 1) & 1);
            tcgen05_fence_after();
            uint32_t yv[16];
            tmem_ld16(t_lane + S::FC_COL, yv);
            tmem_wait_ld16(yv);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(fe0);
            if (valid) {
                float* dst = y + (size_t)(row0 + rloc) * O * Tp + (t - 1);
#pragma unroll
                for (int o = 0; o < 16; ++o)
                    if (o < O) dst[(size_t)o * Tp] = __uint_as_float(yv[o]) + __ldg(fc_b + o);
            }
        };
        uint32_t g0[16], g1[16], g2[16], g3[16], cp[8], hp[8], hkeep[8];
        uint4 zq[8];
        // issue the TMEM loads of position (t, j): 4 blocks of accumulators + the packed c of the chunk
        auto issue_loads = [&](int t, int j) {
            const uint32_t as = j & 1, use = (uint32_t)t * (NCHUNK / 2) + (j >> 1);
            mbar_wait(&acc_full[as], use & 1);
            if (threadIdx.x == 128) TRACE(5);
            if (threadIdx.x == 256) TRACE(10);
            tcgen05_fence_after();
            const uint32_t t_acc = t_lane + as * 128 + half * 64;
            tmem_ld16(t_acc, g0);
            tmem_ld16(t_acc + 16, g1);
            tmem_ld16(t_acc + 32, g2);
            tmem_ld16(t_acc + 48, g3);
            tmem_ld8(t_lane + S::C_COL + chunk_of(j) * 16 + half * 8, cp);
        };
        if (!FUSE_X) {
            const uint4* zp = zx_ptr(0, chunk_of(0));
#pragma unroll
            for (int q = 0; q < 8; ++q) zq[q] = __ldg(zp + q * 32);
        }
        issue_loads(0, 0);
        for (int t = 0; t < Tp; ++t) {
#pragma unroll 1
            for (int j = 0; j < NCHUNK; ++j) {
                const uint32_t as = j & 1;
                const int jc = chunk_of(j);   // the chunk (32 hidden units) this pair processes at position j
                // single wait for the 5 loads of this position; naming every destination keeps their uses below it
                asm volatile("tcgen05.wait::ld.sync.aligned;"
                             : "+r"(g0[0]), "+r"(g0[1]), "+r"(g0[2]), "+r"(g0[3]), "+r"(g0[4]), "+r"(g0[5]), "+r"(g0[6]), "+r"(g0[7]),
                               "+r"(g0[8]), "+r"(g0[9]), "+r"(g0[10]), "+r"(g0[11]), "+r"(g0[12]), "+r"(g0[13]), "+r"(g0[14]), "+r"(g0[15])
                             :: "memory");
                tmem_wait_ld16(g1);
                tmem_wait_ld16(g2);
                tmem_wait_ld16(g3);
                asm volatile("" : "+r"(cp[0]), "+r"(cp[1]), "+r"(cp[2]), "+r"(cp[3]), "+r"(cp[4]), "+r"(cp[5]), "+r"(cp[6]), "+r"(cp[7]));
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(ae0 + as * 8);   // the stage is in registers: its next MMAs may start
                if (threadIdx.x == 128) TRACE(6);
                if (threadIdx.x == 256) TRACE(11);
                int jn = j + 1, tn = t;
                if (jn == NCHUNK) { jn = 0; tn = t + 1; }
                const uint4* zn = zx_ptr(tn < Tp ? tn : t, chunk_of(jn));   // Zx of the next position (clamped at the very end)
                const float* bj = bias + jc * 128 + half * 64;              // fused: warp-uniform (broadcast) bias loads
                // one 4-unit block: g[gate*4 + uu] accumulators, pre-activations / bias, c pairs cp[2b], cp[2b+1] -> hp[2b], hp[2b+1]
                auto block = [&](const uint32_t (&g)[16], int b) {
                    float pi[4], pf[4], pg[4], po[4];
                    if (FUSE_X) {
                        const float4 vi = __ldg(reinterpret_cast<const float4*>(bj + b * 16));
                        const float4 vf = __ldg(reinterpret_cast<const float4*>(bj + b * 16 + 4));
                        const float4 vg = __ldg(reinterpret_cast<const float4*>(bj + b * 16 + 8));
                        const float4 vo = __ldg(reinterpret_cast<const float4*>(bj + b * 16 + 12));
                        pi[0] = vi.x; pi[1] = vi.y; pi[2] = vi.z; pi[3] = vi.w;
                        pf[0] = vf.x; pf[1] = vf.y; pf[2] = vf.z; pf[3] = vf.w;
                        pg[0] = vg.x; pg[1] = vg.y; pg[2] = vg.z; pg[3] = vg.w;
                        po[0] = vo.x; po[1] = vo.y; po[2] = vo.z; po[3] = vo.w;
                    } else {
                        const uint4 za = zq[2 * b], zb = zq[2 * b + 1];
                        const float2 i01 = __half22float2(*reinterpret_cast<const __half2*>(&za.x));
                        const float2 i23 = __half22float2(*reinterpret_cast<const __half2*>(&za.y));
                        const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&za.z));
                        const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&za.w));
                        const float2 g01 = __half22float2(*reinterpret_cast<const __half2*>(&zb.x));
                        const float2 g23 = __half22float2(*reinterpret_cast<const __half2*>(&zb.y));
                        const float2 o01 = __half22float2(*reinterpret_cast<const __half2*>(&zb.z));
                        const float2 o23 = __half22float2(*reinterpret_cast<const __half2*>(&zb.w));
                        pi[0] = i01.x; pi[1] = i01.y; pi[2] = i23.x; pi[3] = i23.y;
                        pf[0] = f01.x; pf[1] = f01.y; pf[2] = f23.x; pf[3] = f23.y;
                        pg[0] = g01.x; pg[1] = g01.y; pg[2] = g23.x; pg[3] = g23.y;
                        po[0] = o01.x; po[1] = o01.y; po[2] = o23.x; po[3] = o23.y;
                        if (tn < Tp) {   // refill in place: consumed one chunk period from now
                            zq[2 * b] = __ldg(zn + (2 * b) * 32);
                            zq[2 * b + 1] = __ldg(zn + (2 * b + 1) * 32);
                        }
                    }
                    float cn[4], hn[4];
                    const float2 c01 = __half22float2(*reinterpret_cast<const __half2*>(&cp[2 * b]));
                    const float2 c23 = __half22float2(*reinterpret_cast<const __half2*>(&cp[2 * b + 1]));
                    const float cprev[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
                    for (int uu = 0; uu < 4; ++uu) {
                        const float zi = __uint_as_float(g[uu]) + pi[uu];
                        const float zf = __uint_as_float(g[4 + uu]) + pf[uu];
                        const float zg = __uint_as_float(g[8 + uu]) + pg[uu];
                        const float zo = __uint_as_float(g[12 + uu]) + po[uu];
                        cn[uu] = sigmoid_half_arg(zf) * cprev[uu] + sigmoid_half_arg(zi) * tanh_fast(zg);
                        hn[uu] = sigmoid_half_arg(zo) * tanh_fast(cn[uu]);
                    }
                    __half2 q;
                    q = __floats2half2_rn(cn[0], cn[1]); cp[2 * b] = *reinterpret_cast<uint32_t*>(&q);
                    q = __floats2half2_rn(cn[2], cn[3]); cp[2 * b + 1] = *reinterpret_cast<uint32_t*>(&q);
                    q = __floats2half2_rn(hn[0], hn[1]); hp[2 * b] = *reinterpret_cast<uint32_t*>(&q);
                    q = __floats2half2_rn(hn[2], hn[3]); hp[2 * b + 1] = *reinterpret_cast<uint32_t*>(&q);
                };
                block(g0, 0);
                block(g1, 1);
                block(g2, 2);
                block(g3, 3);
                tmem_st8(t_lane + S::C_COL + jc * 16 + half * 8, cp);
                if (threadIdx.x == 128) TRACE(7);
                if (threadIdx.x == 256) TRACE(12);
                if (j == NCHUNK - 1) {
                    // acc_full of position 11 was seen: no MMA of this step reads the A operand any more -> positions 10, 11
                    // (the late slab) go straight into it (the store warp fences and publishes them)
                    const int c0 = half * 2, c1 = 4 + half * 2;   // 16-byte chunk of positions 10 / 11 inside the slab row
                    *reinterpret_cast<uint4*>(late_row + (((c0 + 0) ^ lsw) << 4)) = make_uint4(hkeep[0], hkeep[1], hkeep[2], hkeep[3]);
                    *reinterpret_cast<uint4*>(late_row + (((c0 + 1) ^ lsw) << 4)) = make_uint4(hkeep[4], hkeep[5], hkeep[6], hkeep[7]);
                    *reinterpret_cast<uint4*>(late_row + (((c1 + 0) ^ lsw) << 4)) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                    *reinterpret_cast<uint4*>(late_row + (((c1 + 1) ^ lsw) << 4)) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(late_written);
                }
                stage_out(hp, t, j);
                if (threadIdx.x == 128) TRACE(15);
                if (j == NCHUNK - 1) {
                    tmem_wait_st();   // c of this step is in TMEM before any load of the next step
                    if (FUSE_FC && half == 0) fc_out(t + 1);
                    if (t + 1 < Tp) issue_loads(t + 1, 0);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) hkeep[i] = hp[i];
                    issue_loads(t, j + 1);
                }
                if (threadIdx.x == 128) TRACE(8);
                if (threadIdx.x == 256) TRACE(13);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
#ifdef NPPC_REC_TRACE
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < 4 * 12 * 16; i += NTHREADS) g_trace[i] = trace_s[i];
#endif
    cluster_sync_all();  // neither CTA may exit while the pair's MMAs / multicast commits / remote arrives can still target it
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc_pair<512>(tmem_base);
    }
}

// fp32 [4H][K] (nn.LSTM row order) -> fp16 [4H][KP] with permuted rows, zero-padded K
// packed gate column p = chunk*128 + half*64 + blk*16 + gate*4 + uu  ->  nn.LSTM row gate*H + chunk*32 + half*16 + blk*4 + uu
__device__ __forceinline__ int packed_to_lstm_row(int p) {
    const int chunk = p >> 7, half = (p >> 6) & 1, blk = (p >> 4) & 3, gate = (p >> 2) & 3, uu = p & 3;
    return gate * H + chunk * CH + half * 16 + blk * 4 + uu;
}
// i, f, o rows carry the 1/2 of sigmoid(z) = 0.5 tanh(z/2) + 0.5 (gate order i, f, g, o)
__device__ __forceinline__ float packed_gate_scale(int p) { return ((p >> 2) & 3) == 2 ? 1.0f : 0.5f; }
__global__ void pack_w_kernel(const float* __restrict__ w, int K, int KP, __half* __restrict__ out) {
    long long n = (long long)H4 * KP;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int p = (int)(i / KP), k = (int)(i - (long long)p * KP);
        int src = packed_to_lstm_row(p);
        out[i] = __float2half_rn(k < K ? fminf(fmaxf(packed_gate_scale(p) * w[(size_t)src * K + k], -65504.f), 65504.f) : 0.f);
    }
}
__global__ void pack_fc_kernel(const float* __restrict__ w, int O, __half* __restrict__ out) {  // [O][H] f32 -> [16][H] fp16
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 16 * H) out[i] = __float2half_rn(i / H < O ? fminf(fmaxf(w[i], -65504.f), 65504.f) : 0.f);
}
__global__ void pack_b_kernel(const float* __restrict__ b, float* __restrict__ out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < H4) out[p] = packed_gate_scale(p) * b[packed_to_lstm_row(p)];
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

template <bool FUSE_X, bool FUSE_FC>
int launch_rec(const CUtensorMap& tw, const CUtensorMap& th, const CUtensorMap& thst, const CUtensorMap& tx,
               const CUtensorMap& twx, const void* zx, const float* bias, int RS, int Tp, const FcArgs& fc, cudaStream_t s,
               int xkk = 4) {
    auto kern = lstm_rec_kernel<FUSE_X, FUSE_FC>;
    NPPC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, RecSmem<FUSE_X>::TOTAL));
    const int tiles = RS / ROWS;
#ifdef NPPC_LSTM_ABLATE   // compile-time only (build.py adds -DNPPC_LSTM_ABLATE when NPPC_LSTM_ABLATE is set at BUILD time)
    static const int dbg = getenv("NPPC_LSTM_DBG") ? atoi(getenv("NPPC_LSTM_DBG")) : 0;   // timing ablations: wrong results
#else
    constexpr int dbg = 0;   // the shipped library has no run-time switch that changes results
#endif
    kern<<<nppc::cdiv(tiles, 2) * 2, NTHREADS, RecSmem<FUSE_X>::TOTAL, s>>>(tw, th, thst, tx, twx, *fc.tfc, (const uint4*)zx, bias,
                                                                            RS, Tp, fc.fc_b, fc.O, fc.R, fc.y, dbg, xkk);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
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
    // weight boxes are this CTA's HALF of a tile (64 of a chunk's 128 gate columns, 8 of the 16 fc rows) x two K slabs
    CUtensorMap tw[2], th, thst, tx, twx, tfc;
    int rc;
    for (int l = 0; l < 2; ++l) {
        rc = tc::make_tmap_f16_kslabs(&tw[l], p->wp_hh[l], H4, H, 64, 2);
        if (rc) return rc;
    }
    rc = tc::make_tmap_bf16_2d(&th, hseq, (uint64_t)M, H, H * 2, ROWS, 64);
    if (rc) return rc;
    rc = tc::make_tmap_bf16_2d(&thst, hseq, (uint64_t)M, H, H * 2, ROWS, CH, 0);
    if (rc) return rc;
    rc = tc::make_tmap_bf16_2d(&tx, xs, (uint64_t)M, (uint64_t)KP, (uint64_t)KP * 2, ROWS, 64);
    if (rc) return rc;
    rc = tc::make_tmap_bf16_2d(&twx, p->wp_ih[0], H4, (uint64_t)KP, (uint64_t)KP * 2, 64, 64);
    if (rc) return rc;
    rc = tc::make_tmap_f16_kslabs(&tfc, p->wp_fc, 16, H, 8, 2);
    if (rc) return rc;
    // layer 0: input projection (K = 64) fused into the recurrent kernel unless NPPC_LSTM_FUSE_X=0
    static const bool fuse_x = !(getenv("NPPC_LSTM_FUSE_X") && atoi(getenv("NPPC_LSTM_FUSE_X")) == 0);
    static const bool fuse_fc = !(getenv("NPPC_LSTM_FUSE_FC") && atoi(getenv("NPPC_LSTM_FUSE_FC")) == 0);
    const bool fc_in_rec = fuse_fc && p->O <= 16;   // fc fused into the last layer's recurrent kernel (mini-chunk per step)
    const FcArgs fc{&tfc, p->fc_b, p->O, R, y};
    if (fuse_x && KP == 64) {
        rc = launch_rec<true, false>(tw[0], th, thst, tx, twx, nullptr, p->bias_p[0], RS, Tp, fc, s, (p->I + 15) / 16);
    } else {
        rc = gemm_16bit_tn(xs, p->wp_ih[0], p->bias_p[0], zx, M, H4, KP, 2, s);
        if (rc) return rc;
        rc = launch_rec<false, false>(tw[0], th, thst, tx, twx, zx, nullptr, RS, Tp, fc, s);
    }
    if (rc) return rc;
    // layer 1 (+ fc)
    static const bool zx_pair = !(getenv("NPPC_GEMM_ZX") && atoi(getenv("NPPC_GEMM_ZX")) == 0);   // 0 = single-CTA GEMM
    rc = zx_pair ? gemm_zx_pair(hseq, p->wp_ih[1], p->bias_p[1], zx, M, s)
                 : gemm_16bit_tn(hseq, p->wp_ih[1], p->bias_p[1], zx, M, H4, H, 2, s);
    if (rc) return rc;
    if (fc_in_rec) return launch_rec<false, true>(tw[1], th, thst, tx, twx, zx, nullptr, RS, Tp, fc, s);
    rc = launch_rec<false, false>(tw[1], th, thst, tx, twx, zx, nullptr, RS, Tp, fc, s);
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
