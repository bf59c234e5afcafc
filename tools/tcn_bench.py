"""Times one full-band TCN stack (channel-last tcgen05 path) at the bench shape: python tools/tcn_bench.py [B] [C]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import generative_audio_b200 as g

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
C = int(sys.argv[2]) if len(sys.argv) > 2 else 257
torch.manual_seed(0)
m = g.modules.SequenceModel(input_size=C, output_size=257, hidden_size=512, num_layers=2, bidirectional=False,
                            sequence_model="TCN", output_activate_function="ReLU").cuda().eval()
m.use_tc_convs = True
x = torch.rand(B, C, 253, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y = m(x)
    e1.record()
    e1.synchronize()
print(f"TCN stack B={B} C={C}: {e0.elapsed_time(e1) / 20:.3f} ms per stack, checksum {float(y.float().abs().mean()):.6f}")
