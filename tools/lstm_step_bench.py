"""Stepwise tcgen05 LSTM at the training shape (R = 4096 rows, I = 34, H = 384, O = 10): forward(train) + backward timing per
time step, and the target of the ncu capture of lstm_step_fwd_kernel / lstm_step_bwd_kernel.
    python tools/lstm_step_bench.py [R] [Tp]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import generative_audio_b200 as g
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
Tp = int(sys.argv[2]) if len(sys.argv) > 2 else 253
I, H, O = 34, 384, 10
gen = torch.Generator().manual_seed(0)
b = 1.0 / H ** 0.5
u = lambda *s: ((torch.rand(*s, generator=gen) * 2 - 1) * b).cuda()
params = [u(4 * H, I), u(4 * H, H), u(4 * H), u(4 * H), u(4 * H, H), u(4 * H, H), u(4 * H), u(4 * H), u(O, H), u(O)]
RS = -(-R // 128) * 128
xs = torch.zeros(Tp, RS, 64, device="cuda", dtype=torch.float16)
xs[:, :R, :I] = torch.randn(Tp, R, I, generator=gen).cuda().half()
dy = torch.randn(R, O, Tp, generator=gen).cuda() * 1e-4
def run():
    y, ws = g.ops.lstm_step_forward(params, xs, R, train=True)
    return g.ops.lstm_step_backward(params, xs, R, ws, dy)
for _ in range(2):
    run()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); y, ws = g.ops.lstm_step_forward(params, xs, R, train=True); e[1].record(); g.ops.lstm_step_backward(params, xs, R, ws, dy); e[2].record()
torch.cuda.synchronize()
print(f"R={R} Tp={Tp}: forward {e[0].elapsed_time(e[1]):.2f} ms ({1e3 * e[0].elapsed_time(e[1]) / (2 * Tp):.1f} us per layer-step), "
      f"backward {e[1].elapsed_time(e[2]):.2f} ms ({1e3 * e[1].elapsed_time(e[2]) / (2 * Tp):.1f} us per layer-step incl. dW)")
