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
#include "lstm_plan.cuh"
#include "tc_common.cuh"

namespace nppc {
int gemm_bf16_tn(const void* A, const void* W, const float* bias, void* C, long long M, int N, int K, cudaStream_t s);
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

#ifdef NPPC_REC_TRACE
__device__ unsigned long long g_trace[8 * 12 * 16];
#define TRACE(t, j, slot) do { if (blockIdx.x == 0 && (t) >= 2 && (t) < 10) g_trace[(((t) - 2) * 12 + (j)) * 16 + (slot)] = clock64(); } while (0)
#else
#define TRACE(t, j, slot) do { } while (0)
#endif

template <int OP>
struct RecSmemT {
    static constexpr int NST = OP > 16 ? 5 : 6;                          // W ring stages (227 KB budget)
    static constexpr int A_OFF = 0;
    static constexpr int W_OFF = NSLAB * SLAB_BYTES;
    static constexpr int FC_OFF = W_OFF + NST * SLAB_BYTES;              // float [OP][H]
    static constexpr int XCH_OFF = FC_OFF + OP * H * 4;                  // float [128][OP]
    static constexpr int BAR_OFF = XCH_OFF + ROWS * OP * 4;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
    static_assert(TOTAL <= 227 * 1024, "shared memory budget exceeded");
};

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

// CL = cluster size: the CL CTAs of a cluster share ONE W_hh stream — slab s is fetched from L2 by CTA (s % CL) and
// multicast into every CTA's ring, so L2->SM traffic for the weights drops by CL (the kernel is L2-bandwidth-bound).
template <int OP, int CL>
__global__ void __launch_bounds__(NTHREADS, 1)
lstm_rec_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
                const __nv_bfloat16* __restrict__ zx, __nv_bfloat16* __restrict__ hseq, int R, int Tp,
                const float* __restrict__ fc_w, const float* __restrict__ fc_b, int O, float* __restrict__ y, int dbg) {
    using RecSmem = RecSmemT<OP>;
    constexpr int NST = RecSmem::NST;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + RecSmem::BAR_OFF);
    uint64_t* w_empty = w_full + NST;
    uint64_t* a_full = w_empty + NST;      // [NSLAB]
    uint64_t* acc_full = a_full + NSLAB;
    uint64_t* acc_empty = acc_full + 1;
    uint64_t* h_ready = acc_empty + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(h_ready + 1);
    float* fc_s = reinterpret_cast<float*>(smem + RecSmem::FC_OFF);
    float* xch = reinterpret_cast<float*>(smem + RecSmem::XCH_OFF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * ROWS;

    // zero the A operand (h_{-1} = 0) and stage the fc weights
    for (int i = threadIdx.x; i < NSLAB * SLAB_BYTES / 16; i += NTHREADS)
        reinterpret_cast<uint4*>(smem + RecSmem::A_OFF)[i] = make_uint4(0, 0, 0, 0);
    if (OP > 0)
        for (int i = threadIdx.x; i < OP * H; i += NTHREADS) fc_s[i] = (i / H < O) ? fc_w[i] : 0.f;
    fence_proxy_async_smem();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_h);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], CL); }
        for (int i = 0; i < NSLAB; ++i) mbar_init(&a_full[i], 1);
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 256);
        mbar_init(h_ready, 256);
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
            for (int t = 0; t < Tp; ++t)
                for (int j = 0; j < NCHUNK; ++j)
                    for (int k = 0; k < NSLAB; ++k, ++s) {
                        if ((dbg & 1) && s >= (uint32_t)NST) continue;  // timing experiment: W ring filled once, never refilled
                        mbar_wait(&w_empty[stage], phase ^ 1);   // every CTA of the cluster has released this slot
                        mbar_arrive_expect_tx(&w_full[stage], SLAB_BYTES);
                        unsigned char* dst = smem + RecSmem::W_OFF + stage * SLAB_BYTES;
                        if (CL == 1) tma_load_2d(dst, &tmap_w, &w_full[stage], k * 64, j * 128);
                        else if (s % CL == crank) tma_load_2d_mcast(dst, &tmap_w, &w_full[stage], k * 64, j * 128, CMASK);
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 3) {
        // ---- A producer: reload h_t (written to global by the epilogue) as the A operand of step t+1 ----
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        if (lane == 0) {
            for (int k = 0; k < NSLAB; ++k) mbar_arrive(&a_full[k]);  // step 0: zeros already in place
            for (int t = 1; t < Tp; ++t) {
                mbar_wait(h_ready, (t - 1) & 1);
                TRACE(t, 0, 13);
                for (int k = 0; k < NSLAB; ++k) {
                    mbar_arrive_expect_tx(&a_full[k], SLAB_BYTES);
                    tma_load_2d(smem + RecSmem::A_OFF + k * SLAB_BYTES, &tmap_h, &a_full[k], k * 64, (t - 1) * R + row0);
                }
            }
        }
    } else if (warp == 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    } else if (warp == 1) {
        // ---- MMA issuer ----
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(ROWS, 128);
            int stage = 0; uint32_t phase = 0;
            uint32_t it = 0;  // chunk counter for the acc_empty parity
            const uint32_t a_base = smem_u32(smem + RecSmem::A_OFF);
            const uint32_t w_base = smem_u32(smem + RecSmem::W_OFF);
            for (int t = 0; t < Tp; ++t) {
                for (int j = 0; j < NCHUNK; ++j, ++it) {
                    mbar_wait(acc_empty, (it & 1) ^ 1);
                    TRACE(t, j, 0);
                    tcgen05_fence_after();
                    for (int k = 0; k < NSLAB; ++k) {
                        if (j == 0) mbar_wait(&a_full[k], t & 1);
                        if (j == 0 && k == 0) TRACE(t, j, 8);
                        if (!(dbg & 1) || it == 0) mbar_wait(&w_full[stage], phase);
                        TRACE(t, j, 1 + k);
                        tcgen05_fence_after();
                        uint64_t da = umma_desc_k128(a_base + k * SLAB_BYTES);
                        uint64_t db = umma_desc_k128(w_base + stage * SLAB_BYTES);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_base, da + 2 * kk, db + 2 * kk, idesc, (k | kk) != 0);
                        if (dbg & 1) { }
                        else if (CL == 1) umma_commit(&w_empty[stage]);
                        else umma_commit_mcast(&w_empty[stage], CMASK);
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(acc_full);
                    TRACE(t, j, 7);
                }
            }
        }
    } else if (warp >= 4) {
        // ---- epilogue: thread = (row = 32*(warp%4) + lane, half = (warp-4)/4) ----
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const int ew = warp & 3, half = (warp - 4) >> 2;
        const int rloc = ew * 32 + lane;
        const int row = row0 + rloc;
        const bool valid = row < R;
        const uint32_t t_lane = tmem_base + ((uint32_t)(ew * 32) << 16);
        const uint32_t t_acc = t_lane + half * 64;
        // c_0 = 0
        {
            uint32_t z[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = 0u;
            for (int j = 0; j < NCHUNK; ++j) tmem_st16(t_lane + 128 + j * CH + half * 16, z);
            tmem_wait_st();
        }
        uint32_t it = 0;
        // software-pipelined Zx stream: this thread's 64 pre-activations (i,f,g,o x 16 units = 128 contiguous bytes) of
        // chunk it+1 are requested while chunk it is being computed; chunk it+3 is pulled into L2 ahead of that.
        auto zx_ptr = [&](int t, int j) -> const uint4* {
            return reinterpret_cast<const uint4*>(zx + ((size_t)t * R + (valid ? row : 0)) * H4 + half * 64 + j * 128);
        };
        uint4 zraw[8];
        {
            const uint4* zp = zx_ptr(0, 0);
#pragma unroll
            for (int q = 0; q < 8; ++q) zraw[q] = valid ? __ldg(zp + q) : make_uint4(0, 0, 0, 0);
        }
        for (int t = 0; t < Tp; ++t) {
            float fcacc[OP > 0 ? OP : 1];
#pragma unroll
            for (int o = 0; o < (OP > 0 ? OP : 1); ++o) fcacc[o] = 0.f;
            __nv_bfloat16* hrow = hseq + ((size_t)t * R + (valid ? row : 0)) * H + half * 16;
#pragma unroll 1
            for (int j = 0; j < NCHUNK; ++j, ++it) {
                mbar_wait(acc_full, it & 1);
                if (threadIdx.x == 128) TRACE(t, j, 9);
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
                if (threadIdx.x == 128) TRACE(t, j, 10);
                uint4 zcur[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) zcur[q] = zraw[q];
                {
                    int jn = j + 1, tn = t;
                    if (jn == NCHUNK) { jn = 0; tn = t + 1; }
                    if (tn < Tp && valid && !(dbg & 4)) {
                        const uint4* zp = zx_ptr(tn, jn);
#pragma unroll
                        for (int q = 0; q < 8; ++q) zraw[q] = __ldg(zp + q);
                        int jp = j + 3, tp = t;
                        if (jp >= NCHUNK) { jp -= NCHUNK; tp = t + 1; }
                        if (tp < Tp) asm volatile("prefetch.global.L2 [%0];" ::"l"(zx_ptr(tp, jp)));
                    }
                }
                const __nv_bfloat16* zb = reinterpret_cast<const __nv_bfloat16*>(zcur);
                uint32_t hp[8];
                float hv[16];
                if (dbg & 2) {
#pragma unroll
                    for (int u = 0; u < 16; ++u) hv[u] = __uint_as_float(gi[u]) + __bfloat162float(zb[u]);
                } else
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    float zi = __uint_as_float(gi[u]) + __bfloat162float(zb[u]);
                    float zf = __uint_as_float(gf[u]) + __bfloat162float(zb[16 + u]);
                    float zg = __uint_as_float(gg[u]) + __bfloat162float(zb[32 + u]);
                    float zo = __uint_as_float(go[u]) + __bfloat162float(zb[48 + u]);
                    float c = sigmoid_fast(zf) * __uint_as_float(cc[u]) + sigmoid_fast(zi) * tanh_fast(zg);
                    cc[u] = __float_as_uint(c);
                    hv[u] = sigmoid_fast(zo) * tanh_fast(c);
                }
                tmem_st16(t_lane + 128 + j * CH + half * 16, cc);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    __nv_bfloat162 p = __floats2bfloat162_rn(hv[2 * u], hv[2 * u + 1]);
                    hp[u] = *reinterpret_cast<uint32_t*>(&p);
                }
                if (valid) {
                    uint4* dst = reinterpret_cast<uint4*>(hrow + j * CH);
                    dst[0] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                    dst[1] = make_uint4(hp[4], hp[5], hp[6], hp[7]);
                }
                if (threadIdx.x == 128) TRACE(t, j, 11);
                if (OP > 0) {
                    const float* wf = fc_s + j * CH + half * 16;
#pragma unroll
                    for (int o = 0; o < OP; ++o) {
                        const float4* w4 = reinterpret_cast<const float4*>(wf + o * H);
                        float a = fcacc[o];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float4 w = w4[q];
                            a = fmaf(hv[4 * q], w.x, a); a = fmaf(hv[4 * q + 1], w.y, a);
                            a = fmaf(hv[4 * q + 2], w.z, a); a = fmaf(hv[4 * q + 3], w.w, a);
                        }
                        fcacc[o] = a;
                    }
                }
            }
            // end of step: publish h_t to the async proxy (TMA reload), then the fc output of this step
            tmem_wait_st();
            __threadfence();
            fence_proxy_async_all();
            mbar_arrive(h_ready);
            if (threadIdx.x == 128) TRACE(t, 11, 12);
            if (OP > 0) {
                if (half == 1) {
#pragma unroll
                    for (int o = 0; o < OP; ++o) xch[rloc * (OP > 0 ? OP : 1) + o] = fcacc[o];
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (half == 0 && valid) {
                    for (int o = 0; o < O; ++o) {
                        float v = 0.f;
#pragma unroll
                        for (int oo = 0; oo < OP; ++oo) if (oo == o) v = fcacc[oo];
                        y[((size_t)row * O + o) * Tp + t] = v + xch[rloc * (OP > 0 ? OP : 1) + o] + fc_b[o];
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // no CTA may exit while peers can still multicast into it / arrive on its barriers
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// fp32 [4H][K] (nn.LSTM row order) -> bf16 [4H][KP] with permuted rows, zero-padded K
__global__ void pack_w_kernel(const float* __restrict__ w, int K, int KP, __nv_bfloat16* __restrict__ out) {
    long long n = (long long)H4 * KP;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int p = (int)(i / KP), k = (int)(i - (long long)p * KP);
        int chunk = p >> 7, half = (p >> 6) & 1, gate = (p >> 4) & 3, u = p & 15;
        int src = gate * H + chunk * CH + half * 16 + u;
        out[i] = __float2bfloat16(k < K ? w[(size_t)src * K + k] : 0.f);
    }
}
__global__ void pack_b_kernel(const float* __restrict__ b, float* __restrict__ out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < H4) {
        int chunk = p >> 7, half = (p >> 6) & 1, gate = (p >> 4) & 3, u = p & 15;
        out[p] = b[gate * H + chunk * CH + half * 16 + u];
    }
}

template <int OP, int CL>
int launch_rec_cl(const CUtensorMap& tw, const CUtensorMap& th, const __nv_bfloat16* zx, __nv_bfloat16* hseq, int R, int Tp,
                  const float* fc_w, const float* fc_b, int O, float* y, cudaStream_t s) {
    using RecSmem = RecSmemT<OP>;
    auto kern = lstm_rec_kernel<OP, CL>;
    NPPC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, RecSmem::TOTAL));
    int tiles = nppc::cdiv(R, ROWS);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nppc::cdiv(tiles, CL) * CL));  // padded CTAs own no valid rows but keep the cluster protocol
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
    static int dbg = getenv("NPPC_REC_DBG") ? atoi(getenv("NPPC_REC_DBG")) : 0;
    NPPC_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tw, th, zx, hseq, R, Tp, fc_w, fc_b, O, y, dbg));
    NPPC_COUNT_LAUNCH(1);
    return NPPC_OK;
}

int rec_cluster_size() {
    static int cl = -1;
    if (cl < 0) {
        const char* e = getenv("NPPC_LSTM_CLUSTER");
        cl = e ? atoi(e) : 4;
        if (cl != 1 && cl != 2 && cl != 4) cl = 4;
    }
    return cl;
}

template <int OP>
int launch_rec(const CUtensorMap& tw, const CUtensorMap& th, const __nv_bfloat16* zx, __nv_bfloat16* hseq, int R, int Tp,
               const float* fc_w, const float* fc_b, int O, float* y, cudaStream_t s) {
    switch (rec_cluster_size()) {
        case 1: return launch_rec_cl<OP, 1>(tw, th, zx, hseq, R, Tp, fc_w, fc_b, O, y, s);
        case 2: return launch_rec_cl<OP, 2>(tw, th, zx, hseq, R, Tp, fc_w, fc_b, O, y, s);
        default: return launch_rec_cl<OP, 4>(tw, th, zx, hseq, R, Tp, fc_w, fc_b, O, y, s);
    }
}

}  // namespace

namespace nppc {

size_t lstm_workspace_tc(const nppc_lstm_plan* p, int R, int Tp) {
    (void)p;
    size_t rows = (size_t)Tp * R + ROWS;  // + one tile of slack for the row tail of the last step
    return rows * H4 * 2 + rows * H * 2 + 512;
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
        pack_w_kernel<<<256, 256, 0, s>>>(wih[l], Kin, KP, p->wp_ih[l]);
        pack_w_kernel<<<256, 256, 0, s>>>(whh[l], H, H, p->wp_hh[l]);
        pack_b_kernel<<<cdiv(H4, 256), 256, 0, s>>>(p->bias[l], p->bias_p[l]);
    }
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

void lstm_plan_free_tc(nppc_lstm_plan* p) {
    for (int l = 0; l < 2; ++l) {
        cudaFree(p->wp_ih[l]); cudaFree(p->wp_hh[l]); cudaFree(p->bias_p[l]);
    }
}

int lstm_forward_tc(const nppc_lstm_plan* p, const void* xs, int R, int Tp, int KP, void* ws, size_t ws_bytes, float* y,
                    cudaStream_t s) {
    NPPC_CHECK_ARG(p->H == H && p->wp_hh[0], "nppc_lstm_forward(impl 1): the tcgen05 path is built for H=384 (got %d)", p->H);
    NPPC_CHECK_ARG(KP == p->KP0, "nppc_lstm_forward(impl 1): KP must be %d (got %d)", p->KP0, KP);
    NPPC_CHECK_ARG(p->O <= OPMAX, "nppc_lstm_forward(impl 1): output size %d > %d", p->O, OPMAX);
    NPPC_CHECK_ARG(ws_bytes >= lstm_workspace_tc(p, R, Tp), "nppc_lstm_forward: workspace too small");
    NPPC_CHECK_ARG((long long)Tp * R + ROWS < (1LL << 31), "nppc_lstm_forward(impl 1): T'*R too large");
    const size_t rows = (size_t)Tp * R + ROWS;
    uintptr_t base = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
    __nv_bfloat16* zx = (__nv_bfloat16*)base;
    __nv_bfloat16* hseq = zx + rows * H4;
    const long long M = (long long)Tp * R;
    CUtensorMap tw[2], th;
    for (int l = 0; l < 2; ++l) {
        int rc = tc::make_tmap_bf16_2d(&tw[l], p->wp_hh[l], H4, H, H * 2, 128, 64);
        if (rc) return rc;
    }
    int rc = tc::make_tmap_bf16_2d(&th, hseq, (uint64_t)M, H, H * 2, ROWS, 64);
    if (rc) return rc;
    // layer 0
    rc = gemm_bf16_tn(xs, p->wp_ih[0], p->bias_p[0], zx, M, H4, KP, s);
    if (rc) return rc;
    rc = launch_rec<0>(tw[0], th, zx, hseq, R, Tp, nullptr, nullptr, 0, nullptr, s);
    if (rc) return rc;
    // layer 1 (+ fc)
    rc = gemm_bf16_tn(hseq, p->wp_ih[1], p->bias_p[1], zx, M, H4, H, s);
    if (rc) return rc;
    if (p->O <= 8) return launch_rec<8>(tw[1], th, zx, hseq, R, Tp, p->fc_w, p->fc_b, p->O, y, s);
    if (p->O <= 16) return launch_rec<16>(tw[1], th, zx, hseq, R, Tp, p->fc_w, p->fc_b, p->O, y, s);
    return launch_rec<24>(tw[1], th, zx, hseq, R, Tp, p->fc_w, p->fc_b, p->O, y, s);
}

}  // namespace nppc

#ifdef NPPC_REC_TRACE
extern "C" int nppc_debug_rec_trace(unsigned long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, g_trace, sizeof(g_trace)) == cudaSuccess ? 0 : -2;
}
#endif
