"""Times the sub-band LSTM op alone at the bench shape (R = B*257 sequences, T' = 253) with CUDA events.
usage: python tools/lstm_bench.py [B] [O] [impl]   (env NPPC_LSTM_CLUSTER=1|2|4)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

import generative_audio_b200 as g
import weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
O = int(sys.argv[2]) if len(sys.argv) > 2 else 10
impl = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pre = "audio_pc_wrapper.net." if O == 10 else "pretrained_restoration_model."
p = weights.synth_state_dict(5, 0, pre)
lp = "sb_model.sequence_model."
plan = g.ops.LstmPlan(*[p[lp + f"{k}_l{l}"].cuda() for l in (0, 1) for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")],
                      p["sb_model.fc_output_layer.weight"].cuda(), p["sb_model.fc_output_layer.bias"].cuda())
R, Tp = B * 257, 253
torch.manual_seed(0)
dt = torch.float16 if impl == 1 else torch.float32
RS = g.ops.padded_rows(R, dt)
xs = torch.randn(Tp, RS, 64, device="cuda").to(dt)
xs[:, :, 34:] = 0
xs[:, R:] = 0
for _ in range(2):
    y = plan.forward(xs, impl, R)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = plan.forward(xs, impl, R)
    e1.record()
    e1.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
flops = (2 * 1536 * (34 + 384) + 2 * 1536 * 768) * R * Tp
print(f"cluster={os.environ.get('NPPC_LSTM_CLUSTER', 'default')} B={B} O={O} impl={impl}: {ms:.2f} ms/call, "
      f"{flops / ms / 1e9:.1f} TFLOP/s, checksum {y.float().abs().mean().item():.6f}")
