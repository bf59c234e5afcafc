// LSTM plan: derived weight caches for the two implementations of a7 (never serialised; rebuilt from the fp32
// master weights by nppc_lstm_plan_create).
#pragma once
#include "common.cuh"

struct nppc_lstm_plan {
    int I, H, O;
    // impl 0 (fp32 SIMT): transposed fp32 weights, W^T[k][4H], summed biases
    float* w_ihT[2];   // [Kin][4H]
    float* w_hhT[2];   // [H][4H]
    float* bias[2];    // [4H] = b_ih + b_hh
    float* fc_w;       // [O][H]
    float* fc_b;       // [O]
    // impl 1 (fp16 tcgen05): gate-interleaved, K-padded fp16 weights (see lstm_tc.cu; 16-bit buffers are typed __nv_bfloat16* for size only)
    int KP0;                   // padded input width of layer 0 (multiple of 64)
    __nv_bfloat16* wp_ih[2];   // [4H][KP] permuted rows
    __nv_bfloat16* wp_hh[2];   // [4H][H]  permuted rows
    float* bias_p[2];          // [4H] permuted
    __nv_bfloat16* wp_fc;      // [16][H] fp16 fc weights, rows >= O zero (fused fc of the last layer, O <= 16)
};

namespace nppc {
int lstm_forward_f32(const nppc_lstm_plan* p, const float* xs, int R, int RS, int Tp, int KP, void* ws, size_t ws_bytes,
                     float* y, cudaStream_t s);
size_t lstm_workspace_f32(const nppc_lstm_plan* p, int R, int Tp);
int lstm_forward_tc(const nppc_lstm_plan* p, const void* xs, int R, int RS, int Tp, int KP, void* ws, size_t ws_bytes,
                    float* y, cudaStream_t s);
size_t lstm_workspace_tc(const nppc_lstm_plan* p, int R, int Tp);
int lstm_plan_pack_tc(nppc_lstm_plan* p, const float* w_ih0, const float* w_hh0, const float* w_ih1,
                      const float* w_hh1, cudaStream_t s);
void lstm_plan_free_tc(nppc_lstm_plan* p);
}  // namespace nppc
