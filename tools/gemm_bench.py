"""Times nppc_gemm_bf16_tn (row-major TMA-store epilogue) at the LSTM input-projection shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import generative_audio_b200 as g
M = 253 * 16512
for K in (64, 384):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(1536, K, device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(1536, device="cuda")
    for _ in range(2):
        c = g.ops.gemm_bf16_tn(a, w, b)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c = g.ops.gemm_bf16_tn(a, w, b); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    print(f"gemm M={M} N=1536 K={K}: {ms:.2f} ms  {2*M*1536*K/ms/1e9:.0f} TFLOP/s  write {M*1536*2/ms/1e6:.0f} GB/s")
    del a, c
