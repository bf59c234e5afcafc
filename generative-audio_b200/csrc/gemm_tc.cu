// bf16 tensor-core GEMM on tcgen05 + TMA + TMEM:  C[M,N] (bf16) = A[M,K] * W[N,K]^T + bias[N]
// Used for the LSTM input projections (M = T'*R up to ~4.2 M rows, N = 4H = 1536, K = 64 / 384).
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warp 2 = TMEM
// allocator, warps 4-7 = epilogue (TMEM -> registers -> +bias -> bf16 -> swizzled smem -> TMA store).  128 x BN output tiles,
// 64-wide K slabs in a 4-stage shared-memory ring (128-byte swizzle), two TMEM accumulator stages so the epilogue of
// tile i overlaps the MMAs of tile i+1.  Consecutive CTAs take the n-tiles of the same m-block so A is read from
// HBM once and re-used out of L2.
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace {
using namespace nppc::tc;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 3;
constexpr int NTHREADS = 256;

template <int BN>
struct GemmSmem {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int C_OFFSET = STAGES * STAGE_BYTES;       // epilogue staging: BN/64 boxes of [128 rows][64 cols] bf16
    static constexpr int C_BYTES = BM * BN * 2;
    static constexpr int BAR_OFFSET = C_OFFSET + C_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // barriers + alignment slack
    static_assert(TOTAL <= 227 * 1024, "shared memory budget exceeded");
};

// a_wrap: the A operand has only a_wrap K-blocks; K-block kb of the product reads A block kb % a_wrap (split-precision
//         products [hi | lo | hi] x [Whi | Whi | Wlo] without storing hi twice); a_wrap = K / 64 for a plain GEMM.
// OUT32 : C is fp32 [M][N] row-major, written straight from the accumulator registers (no 16-bit rounding of the result).
template <int BN, bool F16, bool ZX, bool OUT32>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const float* __restrict__ bias, long long M, int N, int K,
                    int a_wrap, float* __restrict__ c32) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    using S = GemmSmem<BN>;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (int)((M + BM - 1) / BM), n_tiles = N / BN, kb_count = K / BK;
    const long long num_tiles = (long long)m_tiles * n_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        tma_prefetch_desc(&tmap_c);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<2 * BN>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int m_blk = (int)(tile / n_tiles), n_blk = (int)(tile % n_tiles);
                for (int kb = 0; kb < kb_count; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* sa = smem + stage * S::STAGE_BYTES;
                    unsigned char* sb = sa + S::A_BYTES;
                    mbar_arrive_expect_tx(&full[stage], S::STAGE_BYTES);
                    tma_load_2d(sa, &tmap_a, &full[stage], (kb % a_wrap) * BK, m_blk * BM);
                    tma_load_2d(sb, &tmap_b, &full[stage], kb * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = F16 ? umma_idesc_f16(BM, BN) : umma_idesc_bf16(BM, BN);
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[as], aphase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < kb_count; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tcgen05_fence_after();
                    uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
                    uint64_t da = umma_desc_k128(sa), db = umma_desc_k128(sa + S::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    umma_commit(&empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        int as = 0; uint32_t aphase = 0;
        for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int m_blk = (int)(tile / n_tiles), n_blk = (int)(tile % n_tiles);
            mbar_wait(&tfull[as], aphase);
            tcgen05_fence_after();
            const int rloc = ew * 32 + lane;
            const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + as * BN;
            unsigned char* cs = smem + S::C_OFFSET;
            if (OUT32) {
                const long long grow = (long long)m_blk * BM + rloc;
                float* crow = c32 + (size_t)grow * N + (size_t)n_blk * BN;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(t_row + c * 32, v);
                    tmem_wait_ld();
                    if (grow < M) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                   __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                            if (bias) {
                                const float4 bq = __ldg(reinterpret_cast<const float4*>(bias + n_blk * BN + c * 32) + q);
                                o.x += bq.x; o.y += bq.y; o.z += bq.z; o.w += bq.w;
                            }
                            reinterpret_cast<float4*>(crow + c * 32)[q] = o;
                        }
                    }
                }
                tcgen05_fence_before();
                mbar_arrive(&tempty[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
                continue;
            }
            // the previous tile's TMA stores must have finished reading the staging buffer
            if (threadIdx.x == 128) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(t_row + c * 32, v);
                tmem_wait_ld();
                uint32_t packed[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float a = __uint_as_float(v[2 * i]), b = __uint_as_float(v[2 * i + 1]);
                    if (bias) {
                        a += __ldg(bias + n_blk * BN + c * 32 + 2 * i);
                        b += __ldg(bias + n_blk * BN + c * 32 + 2 * i + 1);
                    }
                    if (F16) {  // saturating: fp16 has the mantissa the LSTM pre-activations need, not the range
                        __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
                        packed[i] = *reinterpret_cast<uint32_t*>(&h);
                    } else {
                        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                        packed[i] = *reinterpret_cast<uint32_t*>(&h);
                    }
                }
                if (ZX) {
                    // LSTM pre-activation layout (see lstm_tc.cu): [m_blk][chunk][warp-quad][half][piece q][lane][8 x fp16],
                    // staged in exactly that order (conflict-free 16-byte stores) so that each 128-column chunk of the tile
                    // is one contiguous 32 KB block in global memory; the recurrent kernel's epilogue thread (same
                    // row <-> lane mapping) reads its 64 values back with fully coalesced 16-byte loads.
                    const int col = c * 32;
                    const int chunk_l = col >> 7, half = (col >> 6) & 1, q0 = (col >> 3) & 7;
                    uint4* dst = reinterpret_cast<uint4*>(cs) + ((((chunk_l * 4 + ew) * 2 + half) * 8 + q0) * 32 + lane);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        dst[q * 32] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                } else {
                    // staging layout = TMA SWIZZLE_128B box [128 rows][64 cols]: 16-byte chunk q of row r lives at q ^ (r & 7)
                    unsigned char* box = cs + (c >> 1) * (BM * 128) + rloc * 128;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        int chunk = ((c & 1) * 4 + q) ^ (rloc & 7);
                        *reinterpret_cast<uint4*>(box + chunk * 16) =
                            make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                    }
                }
            }
            // accumulator stage is free as soon as it is in shared memory
            tcgen05_fence_before();
            mbar_arrive(&tempty[as]);
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (threadIdx.x == 128) {
                if (ZX) {
                    // tmap_c views the interleaved buffer as [blocks*256 rows][64 x 16 bit]: one 32 KB block = 256 rows
#pragma unroll
                    for (int cl = 0; cl < BN / 128; ++cl)
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmap_c),
                                     "r"(smem_u32(cs + cl * 32768)), "r"(0),
                                     "r"((m_blk * (N >> 7) + n_blk * (BN / 128) + cl) * 256)
                                     : "memory");
                } else {
#pragma unroll
                    for (int cb = 0; cb < BN / 64; ++cb)
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmap_c),
                                     "r"(smem_u32(cs + cb * (BM * 128))), "r"(n_blk * BN + cb * 64), "r"(m_blk * BM)
                                     : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        if (threadIdx.x == 128) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<2 * BN>(tmem_base);
    }
}

template <int BN, bool F16, bool ZX, bool OUT32 = false>
int launch_gemm(const void* A, const void* W, const float* bias, void* C, long long M, int N, int K, cudaStream_t s, int KA = 0) {
    if (KA <= 0) KA = K;
    CUtensorMap ta, tb;
    int rc = make_tmap_bf16_2d(&ta, A, (uint64_t)M, (uint64_t)KA, (uint64_t)KA * 2, BM, BK);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tb, W, (uint64_t)N, (uint64_t)K, (uint64_t)K * 2, BN, BK);
    if (rc) return rc;
    CUtensorMap tcm;
    if (OUT32) tcm = tb;   // unused by the fp32 epilogue
    else if (ZX) rc = make_tmap_bf16_2d(&tcm, C, (uint64_t)(M / 128) * (N / 128) * 256, 64, 128, 256, 64, 0);
    else rc = make_tmap_bf16_2d(&tcm, C, (uint64_t)M, (uint64_t)N, (uint64_t)N * 2, BM, 64);
    if (rc) return rc;
    using S = GemmSmem<BN>;
    NPPC_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, F16, ZX, OUT32>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    long long tiles = ((M + BM - 1) / BM) * (N / BN);
    int grid = (int)(tiles < nppc::sm_count() ? tiles : nppc::sm_count());
    gemm_bf16_tn_kernel<BN, F16, ZX, OUT32><<<grid, NTHREADS, S::TOTAL, s>>>(ta, tb, tcm, bias, M, N, K, KA / BK, OUT32 ? (float*)C : nullptr);
    NPPC_COUNT_LAUNCH(1);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

}  // namespace

namespace nppc {
// f16 != 0: A, W and C are IEEE fp16 instead of bf16 (same kernel, other tcgen05 operand format)
int gemm_16bit_tn(const void* A, const void* W, const float* bias, void* C, long long M, int N, int K, int f16,
                  cudaStream_t s) {
    NPPC_CHECK_ARG(A && W && C, "nppc_gemm_bf16_tn: null pointer");
    NPPC_CHECK_ARG(M > 0 && N > 0 && K > 0 && K % BK == 0 && N % 128 == 0,
                   "nppc_gemm_bf16_tn: need K %% 64 == 0 and N %% 128 == 0 (M=%lld N=%d K=%d)", M, N, K);
    NPPC_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && ((uintptr_t)C % 16 == 0),
                   "nppc_gemm_bf16_tn: pointers must be 16-byte aligned");
    if (f16 == 2) {  // fp16 operands, output in the LSTM pre-activation layout
        NPPC_CHECK_ARG(M % 128 == 0 && N % 256 == 0 && (M / 128) * (N / 128) * 256 < (1LL << 31),
                       "gemm (LSTM layout): need M %% 128 == 0, N %% 256 == 0");
        return launch_gemm<256, true, true>(A, W, bias, C, M, N, K, s);
    }
    if (f16) return N % 256 == 0 ? launch_gemm<256, true, false>(A, W, bias, C, M, N, K, s) : launch_gemm<128, true, false>(A, W, bias, C, M, N, K, s);
    return N % 256 == 0 ? launch_gemm<256, false, false>(A, W, bias, C, M, N, K, s) : launch_gemm<128, false, false>(A, W, bias, C, M, N, K, s);
}
}  // namespace nppc

extern "C" int nppc_gemm_bf16_tn(const void* A, const void* W, const float* bias, void* C, long long M, int N, int K,
                                 void* stream) {
    return nppc::gemm_16bit_tn(A, W, bias, C, M, N, K, 0, (cudaStream_t)stream);
}

// fp16 operands with a wrapped A operand (A is [M][KA], K-block kb reads A block kb % (KA/64)) and optional fp32 output:
// the split-precision products of the TCN path (tcn_cl.cu): A = [hi | lo], W = [Whi | Whi | Wlo], K = 3 Kp, KA = 2 Kp.
extern "C" int nppc_gemm_f16_tn_ex(const void* A, const void* W, const float* bias, void* C, long long M, int N, int K, int KA,
                                   int out_f32, void* stream) {
    NPPC_CHECK_ARG(A && W && C, "nppc_gemm_f16_tn_ex: null pointer");
    NPPC_CHECK_ARG(M > 0 && N > 0 && K > 0 && K % BK == 0 && N % 128 == 0 && KA > 0 && KA % BK == 0 && KA <= K,
                   "nppc_gemm_f16_tn_ex: need K %% 64 == 0, KA %% 64 == 0, KA <= K, N %% 128 == 0 (M=%lld N=%d K=%d KA=%d)", M, N, K, KA);
    NPPC_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && ((uintptr_t)C % 16 == 0),
                   "nppc_gemm_f16_tn_ex: pointers must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    if (out_f32) return N % 256 == 0 ? launch_gemm<256, true, false, true>(A, W, bias, C, M, N, K, s, KA)
                                     : launch_gemm<128, true, false, true>(A, W, bias, C, M, N, K, s, KA);
    return N % 256 == 0 ? launch_gemm<256, true, false>(A, W, bias, C, M, N, K, s, KA) : launch_gemm<128, true, false>(A, W, bias, C, M, N, K, s, KA);
}

// same with IEEE fp16 operands / output (the TCN 1x1 convolutions, tcn_cl.cu)
extern "C" int nppc_gemm_f16_tn(const void* A, const void* W, const float* bias, void* C, long long M, int N, int K,
                                void* stream) {
    return nppc::gemm_16bit_tn(A, W, bias, C, M, N, K, 1, (cudaStream_t)stream);
}
