"""Per-kernel roofline numbers for the HBM-bound kernels at the bench shape (B utterances x 4 s):
achieved GB/s = ALGORITHMIC bytes (SURVEY.md §8d / DESIGN.md) / CUDA-event time, vs MEASURED_PEAKS.json hbm_gbs.
Inputs are larger than L2 at B=64 for the big kernels; a 256 MiB L2 flush runs between timed launches.
usage: python tools/kernel_bench.py [B] > profiles/r01_kernel_rooflines.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

import generative_audio_b200 as g

ops = g.ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L, F, T, Tp, N = 64000, 257, 251, 253, 5
peak = 6500.3
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
src = "fallback"
if os.path.exists(pk):
    peak, src = json.load(open(pk))["hbm_gbs"], "measured"
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.manual_seed(0)


ONCE = os.environ.get("NPPC_KB_ONCE") == "1"   # one launch per op: the run that ncu --set full captures


def timeit(fn, reps=7):
    if ONCE:
        reps = 1
    for _ in range(0 if ONCE else 3):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


wave = torch.randn(B, L, device=dev) * 0.05
mag, re, im = ops.stft_mri(wave)
re3, im3 = re[:, 0].contiguous(), im[:, 0].contiguous()
crm = torch.randn(B, 2, F, T, device=dev)
x4 = torch.rand(B, 1, F, Tp, device=dev) + 0.1
planes = [torch.rand(B, F, Tp, device=dev) for _ in range(4)]
sb = torch.rand(B, F, 34, Tp, device=dev)
head = torch.randn(B, N, 2, F, T, device=dev)
gt, pred = torch.randn(B, 2, F, T, device=dev), torch.randn(B, 2, F, T, device=dev)
FT = F * T
rows = []


def add(name, ref, nbytes, fn):
    ms = timeit(fn)
    gbs = nbytes / ms / 1e6
    rows.append({"kernel": name, "reference": ref, "algorithmic_bytes": nbytes, "ms": ms, "achieved_gbs": gbs,
                 "peak_gbs": peak, "frac": gbs / peak})


add("stft_mri (a1)", "utils.py:107-147", B * (4 * L + 3 * 4 * FT), lambda: ops.stft_mri(wave))
add("istft (a15)", "utils.py:60-70", B * (2 * 4 * FT + 4 * L), lambda: ops.istft(re3, im3, L))
add("crm_decompress_apply (a9+a10)", "mask.py:57-60, utils.py:241-249", B * 7 * 4 * FT,
    lambda: ops.crm_decompress_apply(crm, re3, im3, True))
add("pad_offline_laplace_norm [1,257,253] (a2)", "base_model.py:210-224", B * (4 * F * T + 4 * F * Tp),
    lambda: ops.pad_offline_laplace_norm(x4[..., :T].contiguous(), 2))
add("offline_laplace_norm [257,34,253] (a2)", "base_model.py:210-224", B * 2 * 4 * F * 34 * Tp, lambda: ops.offline_laplace_norm(sb))
add("cumulative_laplace_norm [257,34,253] (a2)", "base_model.py:227-257", B * 2 * 4 * F * 34 * Tp,
    lambda: ops.cumulative_laplace_norm(sb))
add("unfold N=15 (a5)", "base_model.py:15-46", B * (4 * F * Tp + 4 * F * 31 * Tp), lambda: ops.unfold(x4, 15))
add("subband_pack fused unfold+cat+norm -> fp16 [T',R,64] (a5+a2)", "fullsubnet_plus.py:203-223",
    B * (4 * 4 * F * Tp + 2 * F * 64 * Tp), lambda: ops.subband_pack(*planes, 15, 1, 64, torch.float16))
add("gram_schmidt_complex n=5 (a12)", "pc_wrapper.py:8-44", B * 2 * N * 2 * FT * 4, lambda: ops.gram_schmidt_complex(head))
add("gs_loss_fused n=5 (a12+a14)", "pc_wrapper.py:8-44 + trainer.py:259-298", B * (2 * N * 2 * FT * 4 + 2 * 2 * FT * 4),
    lambda: ops.gs_loss_fused(head, gt, pred))
add("projection_loss n=5 (a14)", "trainer.py:259-298", B * (N * 2 * FT * 4 + 2 * 2 * FT * 4), lambda: ops.projection_loss(head, gt, pred))
print(json.dumps({"batch": B, "peak_source": src, "note": "CUDA events, median of 7, L2 flushed between launches",
                  "kernels": rows}, indent=1))
