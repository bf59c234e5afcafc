// a7, implementation 0: fp32 SIMT 2-layer LSTM + fc.  Reference-precision path (parity 1e-4) and the on-GPU
// yardstick for the tensor-core path.  One CTA owns RT sequences for all T' steps (h in shared memory, c in
// registers); thread j owns hidden unit j and accumulates its 4 gates for the RT rows; weights stream from L2
// (all CTAs walk the same transposed weight matrix in the same order).
#include <string.h>
#include "lstm_plan.cuh"

namespace {

constexpr int RT = 16;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// xs [Tp][R][KP] fp32 (first Kin columns used); hseq_out [Tp][R][H] or null; y [R][O][Tp] or null (last layer)
__global__ void __launch_bounds__(384) lstm_layer_f32_kernel(const float* __restrict__ xs, int R, int RS, int Tp, int KP,
                                                             int Kin, int H, const float* __restrict__ w_ihT,
                                                             const float* __restrict__ w_hhT,
                                                             const float* __restrict__ bias, float* __restrict__ hseq_out,
                                                             const float* __restrict__ fc_w, const float* __restrict__ fc_b,
                                                             int O, float* __restrict__ y) {
    extern __shared__ __align__(16) float sm[];
    float* hs = sm;              // [H][RT]
    float* xsm = sm + H * RT;    // [Kin][RT]
    const int j = threadIdx.x;   // hidden unit
    const int row0 = blockIdx.x * RT;
    const int H4 = 4 * H;
    float c[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) c[r] = 0.f;
    for (int i = threadIdx.x; i < H * RT; i += blockDim.x) hs[i] = 0.f;
    const float b_i = bias[j], b_f = bias[H + j], b_g = bias[2 * H + j], b_o = bias[3 * H + j];
    for (int t = 0; t < Tp; ++t) {
        // stage the input tile transposed: xsm[k][r]
        for (int idx = threadIdx.x; idx < RT * Kin; idx += blockDim.x) {
            int r = idx / Kin, k = idx - r * Kin;
            int row = row0 + r;
            xsm[k * RT + r] = (row < R) ? xs[((size_t)t * RS + row) * KP + k] : 0.f;
        }
        __syncthreads();
        float acc[4][RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) { acc[0][r] = b_i; acc[1][r] = b_f; acc[2][r] = b_g; acc[3][r] = b_o; }
        for (int k = 0; k < Kin; ++k) {
            const float* w = w_ihT + (size_t)k * H4 + j;
            float w0 = w[0], w1 = w[H], w2 = w[2 * H], w3 = w[3 * H];
            const float4* xv = reinterpret_cast<const float4*>(xsm + k * RT);
#pragma unroll
            for (int q = 0; q < RT / 4; ++q) {
                float4 v = xv[q];
                float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc[0][4 * q + e] += w0 * vv[e];
                    acc[1][4 * q + e] += w1 * vv[e];
                    acc[2][4 * q + e] += w2 * vv[e];
                    acc[3][4 * q + e] += w3 * vv[e];
                }
            }
        }
        for (int k = 0; k < H; ++k) {
            const float* w = w_hhT + (size_t)k * H4 + j;
            float w0 = w[0], w1 = w[H], w2 = w[2 * H], w3 = w[3 * H];
            const float4* hv = reinterpret_cast<const float4*>(hs + k * RT);
#pragma unroll
            for (int q = 0; q < RT / 4; ++q) {
                float4 v = hv[q];
                float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc[0][4 * q + e] += w0 * vv[e];
                    acc[1][4 * q + e] += w1 * vv[e];
                    acc[2][4 * q + e] += w2 * vv[e];
                    acc[3][4 * q + e] += w3 * vv[e];
                }
            }
        }
        __syncthreads();  // everyone is done reading hs / xsm
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            float ig = sigmoidf_(acc[0][r]), fg = sigmoidf_(acc[1][r]), gg = tanhf(acc[2][r]), og = sigmoidf_(acc[3][r]);
            c[r] = fg * c[r] + ig * gg;
            float h = og * tanhf(c[r]);
            hs[j * RT + r] = h;
            int row = row0 + r;
            if (hseq_out && row < R) hseq_out[((size_t)t * R + row) * H + j] = h;
        }
        __syncthreads();
        if (y && threadIdx.x < O * RT) {
            int o = threadIdx.x / RT, r = threadIdx.x - o * RT;
            int row = row0 + r;
            float a = fc_b[o];
            for (int k = 0; k < H; ++k) a += fc_w[(size_t)o * H + k] * hs[k * RT + r];
            if (row < R) y[((size_t)row * O + o) * Tp + t] = a;
        }
        // (next iteration's first __syncthreads orders these hs reads before the next hs writes)
    }
}

__global__ void transpose_kernel(const float* __restrict__ w, int rows, int cols, float* __restrict__ wT) {
    // w [rows][cols] -> wT [cols][rows]
    long long n = (long long)rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
        wT[(size_t)c * rows + r] = w[i];
    }
}
__global__ void add_kernel(const float* a, const float* b, int n, float* o) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i] + b[i];
}

}  // namespace

namespace nppc {

size_t lstm_workspace_f32(const nppc_lstm_plan* p, int R, int Tp) { return sizeof(float) * (size_t)Tp * R * p->H; }

int lstm_forward_f32(const nppc_lstm_plan* p, const float* xs, int R, int RS, int Tp, int KP, void* ws, size_t ws_bytes,
                     float* y, cudaStream_t s) {
    NPPC_CHECK_ARG(ws_bytes >= lstm_workspace_f32(p, R, Tp), "nppc_lstm_forward: workspace too small (%zu < %zu)", ws_bytes,
                   lstm_workspace_f32(p, R, Tp));
    NPPC_CHECK_ARG(p->H % 32 == 0 && p->H <= 384, "nppc_lstm_forward(impl 0): H must be a multiple of 32, <= 384 (got %d)", p->H);
    NPPC_CHECK_ARG(p->O * RT <= p->H, "nppc_lstm_forward(impl 0): O*%d must be <= H", RT);
    NPPC_CHECK_ARG(KP >= p->I, "nppc_lstm_forward: KP (%d) < input size (%d)", KP, p->I);
    float* hseq = (float*)ws;
    int grid = cdiv(R, RT);
    size_t smem0 = sizeof(float) * RT * (p->H + p->I), smem1 = sizeof(float) * RT * (2 * p->H);
    NPPC_CUDA_OK(cudaFuncSetAttribute(lstm_layer_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    lstm_layer_f32_kernel<<<grid, p->H, smem0, s>>>(xs, R, RS, Tp, KP, p->I, p->H, p->w_ihT[0], p->w_hhT[0], p->bias[0], hseq,
                                                   nullptr, nullptr, 0, nullptr);
    lstm_layer_f32_kernel<<<grid, p->H, smem1, s>>>(hseq, R, R, Tp, p->H, p->H, p->H, p->w_ihT[1], p->w_hhT[1], p->bias[1],
                                                   nullptr, p->fc_w, p->fc_b, p->O, y);
    NPPC_COUNT_LAUNCH(2);
    NPPC_LAUNCH_OK();
    return NPPC_OK;
}

}  // namespace nppc

extern "C" int nppc_lstm_plan_create(nppc_lstm_plan** plan, int I, int H, int O, const float* w_ih0, const float* w_hh0,
                                     const float* b_ih0, const float* b_hh0, const float* w_ih1, const float* w_hh1,
                                     const float* b_ih1, const float* b_hh1, const float* fc_w, const float* fc_b,
                                     void* stream) {
    NPPC_CHECK_ARG(plan && w_ih0 && w_hh0 && b_ih0 && b_hh0 && w_ih1 && w_hh1 && b_ih1 && b_hh1 && fc_w && fc_b,
                   "nppc_lstm_plan_create: null pointer");
    NPPC_CHECK_ARG(I > 0 && H > 0 && O > 0, "nppc_lstm_plan_create: bad sizes I=%d H=%d O=%d", I, H, O);
    cudaStream_t s = (cudaStream_t)stream;
    nppc_lstm_plan* p = new nppc_lstm_plan();
    memset(p, 0, sizeof(*p));
    p->I = I; p->H = H; p->O = O;
    const int H4 = 4 * H;
    const float* wih[2] = {w_ih0, w_ih1};
    const float* whh[2] = {w_hh0, w_hh1};
    const float* bih[2] = {b_ih0, b_ih1};
    const float* bhh[2] = {b_hh0, b_hh1};
    for (int l = 0; l < 2; ++l) {
        int Kin = l == 0 ? I : H;
        NPPC_CUDA_OK(cudaMalloc(&p->w_ihT[l], sizeof(float) * (size_t)Kin * H4));
        NPPC_CUDA_OK(cudaMalloc(&p->w_hhT[l], sizeof(float) * (size_t)H * H4));
        NPPC_CUDA_OK(cudaMalloc(&p->bias[l], sizeof(float) * H4));
        transpose_kernel<<<256, 256, 0, s>>>(wih[l], H4, Kin, p->w_ihT[l]);
        transpose_kernel<<<256, 256, 0, s>>>(whh[l], H4, H, p->w_hhT[l]);
        add_kernel<<<nppc::cdiv(H4, 256), 256, 0, s>>>(bih[l], bhh[l], H4, p->bias[l]);
    }
    NPPC_CUDA_OK(cudaMalloc(&p->fc_w, sizeof(float) * (size_t)O * H));
    NPPC_CUDA_OK(cudaMalloc(&p->fc_b, sizeof(float) * O));
    NPPC_CUDA_OK(cudaMemcpyAsync(p->fc_w, fc_w, sizeof(float) * (size_t)O * H, cudaMemcpyDeviceToDevice, s));
    NPPC_CUDA_OK(cudaMemcpyAsync(p->fc_b, fc_b, sizeof(float) * O, cudaMemcpyDeviceToDevice, s));
    NPPC_LAUNCH_OK();
    int rc = nppc::lstm_plan_pack_tc(p, w_ih0, w_hh0, w_ih1, w_hh1, s);
    if (rc) return rc;
    *plan = p;
    return NPPC_OK;
}

extern "C" void nppc_lstm_plan_destroy(nppc_lstm_plan* p) {
    if (!p) return;
    for (int l = 0; l < 2; ++l) {
        cudaFree(p->w_ihT[l]); cudaFree(p->w_hhT[l]); cudaFree(p->bias[l]);
    }
    cudaFree(p->fc_w); cudaFree(p->fc_b);
    nppc::lstm_plan_free_tc(p);
    delete p;
}

extern "C" size_t nppc_lstm_workspace_bytes(const nppc_lstm_plan* plan, int R, int Tp, int impl) {
    if (!plan || R <= 0 || Tp <= 0) return 0;
    return impl == 0 ? nppc::lstm_workspace_f32(plan, R, Tp) : nppc::lstm_workspace_tc(plan, R, Tp);
}

extern "C" int nppc_lstm_forward(const nppc_lstm_plan* plan, const void* xs, int R, int R_stride, int Tp, int KP, int impl,
                                 void* workspace, size_t workspace_bytes, float* y, void* stream) {
    NPPC_CHECK_ARG(plan && xs && y && workspace, "nppc_lstm_forward: null pointer");
    NPPC_CHECK_ARG(R > 0 && Tp > 0 && R_stride >= R, "nppc_lstm_forward: bad sizes R=%d R_stride=%d Tp=%d", R, R_stride, Tp);
    if (impl == 0) return nppc::lstm_forward_f32(plan, (const float*)xs, R, R_stride, Tp, KP, workspace, workspace_bytes, y, (cudaStream_t)stream);
    if (impl == 1) return nppc::lstm_forward_tc(plan, xs, R, R_stride, Tp, KP, workspace, workspace_bytes, y, (cudaStream_t)stream);
    nppc::set_error("nppc_lstm_forward: unknown impl %d", impl);
    return NPPC_ERR_INVALID_ARGUMENT;
}
