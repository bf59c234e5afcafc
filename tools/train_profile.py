"""Where the training step (config 3) spends its time: torch profiler, top CUDA kernels. usage: python tools/train_profile.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import generative_audio_b200 as g
from helpers import build_model, wave
m2, _ = build_model(5, 2, "tc")
stepper = g.NPPCAudioStep(m2, 500, 1.0)
opt = torch.optim.Adam(m2.audio_pc_wrapper.parameters(), lr=1e-4)
clean = wave(32, 64000, 3, 0.03).cuda()
noisy = clean + 0.3 * wave(32, 64000, 4, 1.0).cuda()
for _ in range(2):
    stepper.train_step((noisy, clean), opt)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    stepper.train_step((noisy, clean), opt)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=80))
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
stepper.train_step((noisy, clean), opt)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0):.1f} ms, total {1e3 * (t2 - t0):.1f} ms")
