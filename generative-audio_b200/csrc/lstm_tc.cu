// a7, implementation 1: bf16 tcgen05 tensor-core LSTM (placeholder until the tcgen05 kernels land).
#include "lstm_plan.cuh"
namespace nppc {
size_t lstm_workspace_tc(const nppc_lstm_plan*, int, int) { return 0; }
int lstm_plan_pack_tc(nppc_lstm_plan*, const float*, const float*, const float*, const float*, cudaStream_t) { return NPPC_OK; }
void lstm_plan_free_tc(nppc_lstm_plan*) {}
int lstm_forward_tc(const nppc_lstm_plan*, const void*, int, int, int, void*, size_t, float*, cudaStream_t) {
    set_error("nppc_lstm_forward: impl 1 (tcgen05) not built yet");
    return NPPC_ERR_UNSUPPORTED;
}
}  // namespace nppc
extern "C" int nppc_gemm_bf16_tn(const void*, const void*, const float*, void*, long long, int, int, void*) {
    nppc::set_error("nppc_gemm_bf16_tn: not built yet");
    return NPPC_ERR_UNSUPPORTED;
}
