#!/usr/bin/env python
"""BASELINE config 1 shape on the GPU: ONE 4 s utterance through the public serving call (NPPCModel.forward_host: pinned host
waveform in, pinned host w_mat out), latency per call — where the ~330 Python->C launches and the plan-key checks show.
    python tools/latency_b1.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

from helpers import build_model, wave

out = {}
for impl in ("tc", "tcp", "f32"):
    m, _ = build_model(5, 1, impl)
    x = wave(1, 64000, 7).pin_memory()
    o = torch.empty(1, 5, 2, 257, 251).pin_memory()
    for _ in range(3):
        m.forward_host(x, o)
        m.host_copy_done()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        m.forward_host(x, o)
        m.host_copy_done()
        ts.append((time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter()
    m.forward_host(x, o)
    enq = (time.perf_counter() - t0) * 1e3
    m.host_copy_done()
    ts.sort()
    out[impl] = {"latency_ms_median": ts[len(ts) // 2], "latency_ms_min": ts[0], "host_enqueue_ms": enq,
                 "audio_s_per_s": 4.0 / (ts[len(ts) // 2] * 1e-3)}
print(json.dumps({"what": "B = 1 x 4 s, NPPCModel.forward_host wall-clock latency (host in -> host out)", "results": out}, indent=1))
