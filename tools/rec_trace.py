"""Debug: per-chunk timeline of the recurrent LSTM kernel (library built with NPPC_REC_TRACE=1)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import generative_audio_b200 as g
import weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
p = weights.synth_state_dict(5, 0, "pretrained_restoration_model.")
lp = "sb_model.sequence_model."
plan = g.ops.LstmPlan(*[p[lp + f"{k}_l{l}"].cuda() for l in (0, 1) for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")],
                      p["sb_model.fc_output_layer.weight"].cuda(), p["sb_model.fc_output_layer.bias"].cuda())
R, Tp = B * 257, 40
RS = g.ops.padded_rows(R, torch.float16)
torch.manual_seed(0)
xs = torch.randn(Tp, RS, 64, device="cuda").to(torch.float16); xs[:, :, 34:] = 0
y = plan.forward(xs, 1, R)
torch.cuda.synchronize()
lib = g._lib.load()
buf = (C.c_longlong * (4 * 12 * 16))()
lib.nppc_debug_rec_trace.argtypes = [C.c_void_p]
assert lib.nppc_debug_rec_trace(buf) == 0
a = np.array(buf, dtype=np.int64).reshape(4, 12, 16)
t0 = a[0, 0, 0]
print("MMA: 0 acc_empty ok | 1 w_full[0] ok | 3 w_full[last] ok | 4 commit issued || epi warp 4 (half 0): 5 acc_full seen (loads issued next) | 6 wait::ld done, acc_empty arrived | 7 math+STTM done | 15 h staged | 8 next loads issued || W producer: 9 w_empty ok (k=0) | 14 (k=last) || epi warp 8 (half 1, same scheduler): 10 acc_full seen | 11 wait::ld done | 12 math done | 13 next loads issued")
for t in range(3):
    for j in range(12):
        r = a[t, j] - t0
        print(f"t={t+5} j={j:2d} " + " ".join(f"{int(v):7d}" if a[t, j, i] != 0 else "      -" for i, v in enumerate(r[:16])))
