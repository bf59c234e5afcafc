"""Throughput of the other BASELINE.json configurations on one B200 (synthetic data, deterministic synthetic weights):
  #2 enhance-only (STFT -> FullSubNet+ -> cRM -> iSTFT), batch 64 x 4 s
  #3 NPPC-audio training step (frozen backbone on the kernels, PC head through autograd, Adam), batch 32, groups 2
  #4 inpainting NPPC forward, batch 128 x [128 x 500] log-mag spectrograms, n_dirs = 10 (library convolutions)
     + one training iteration of the inpainting NPPC trainer at the same batch size
  N1 validator consumer (pc_variations: 5 directions x 6 alphas -> audio), batch 8 x 4 s
CUDA events, 3 warm-up + 5 timed iterations, 256 MiB L2 flush between iterations.  usage: python tools/config_bench.py"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

import generative_audio_b200 as g
import weights
from helpers import build_model, wave

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5, warm=3):
    for _ in range(warm):
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


out = {}
with torch.no_grad():
    m, _ = build_model(5, 1, "tc")
    x = wave(64, 64000, 1).cuda()
    ms = timeit(lambda: m.enhance(x))
    out["config2_enhance_only_b64"] = {"ms_per_step": ms, "audio_s_per_s": 64 * 4.0 / (ms * 1e-3)}
    x8 = wave(8, 64000, 2).cuda()
    ms = timeit(lambda: m.pc_variations(x8))
    out["n1_pc_variations_b8"] = {"ms_per_step": ms, "waveforms_per_step": 8 * 5 * 6 + 8, "audio_s_per_s": 8 * 4.0 / (ms * 1e-3)}

m2, _ = build_model(5, 2, "tc")
stepper = g.NPPCAudioStep(m2, 500, 1.0)
opt = torch.optim.Adam(m2.audio_pc_wrapper.parameters(), lr=1e-4)
clean = wave(32, 64000, 3, 0.03)
noisy = clean + 0.3 * wave(32, 64000, 4, 1.0)
ms = timeit(lambda: stepper.train_step((noisy.cuda(), clean.cuda()), opt), reps=3, warm=2)
out["config3_train_step_b32_g2"] = {"ms_per_step": ms, "audio_s_per_s": 32 * 4.0 / (ms * 1e-3),
                                     "note": "frozen half on the inference kernels; PC head fwd + bwd hand-written (stepwise tcgen05 LSTM + BPTT, "
                                             "tcgen05 GEMMs for all 1x1 convs fwd/dX/dW, coefficient-space GS + loss backward), Adam step"}
del stepper, opt, m2
torch.cuda.empty_cache()

with torch.no_grad():
    I = g.inpainting
    N_DIRS = 10
    rest = I.UNet(I.UNetConfig(in_channels=1, out_channels=1))
    shp = lambda mod: [(k, tuple(v.shape)) for k, v in mod.state_dict().items()]
    rest.load_state_dict(weights.synth_unet_state_dict(shp(rest), 0, "rest."))
    ck = os.path.join(tempfile.mkdtemp(), "rest.pt")
    torch.save({"model_state_dict": rest.state_dict()}, ck)
    cfg = I.NPPCModelConfig(pretrained_restoration_model_configuration=I.UNetConfig(in_channels=1, out_channels=1),
                            pretrained_restoration_model_path=ck,
                            audio_pc_wrapper_configuration=I.AudioInpaintingPCWrapperConfig(
                                model_configuration=I.UNetConfig(in_channels=2, out_channels=N_DIRS), n_dirs=N_DIRS))
    mi = I.NPPCModel(cfg)
    mi.pc_wrapper.net.load_state_dict(weights.synth_unet_state_dict(shp(mi.pc_wrapper.net), 0, "head."))
    B = 128
    spec = torch.randn(B, 2, 128, 500, device=dev)
    mask = torch.ones(B, 500, device=dev)
    mask[:, 200:218] = 0
    clean_n, m4, masked_n = I.preprocess_data(spec, spec * mask[:, None, None, :], mask)
    ms = timeit(lambda: mi(masked_n, m4), reps=3, warm=2)
    out["config4_inpainting_b128_ndirs10"] = {"ms_per_step": ms, "audio_s_per_s": B * 4.0 / (ms * 1e-3),
                                              "note": "UNet convolutions on cuDNN (row N4), glue + real Gram-Schmidt on the kernels"}
    I.set_compute_dtype(mi, torch.float16)
    w32 = None
    ms = timeit(lambda: mi(masked_n, m4), reps=3, warm=2)
    out["config4_inpainting_b128_ndirs10_fp16convs"] = {"ms_per_step": ms, "audio_s_per_s": B * 4.0 / (ms * 1e-3),
                                                        "note": "same with the UNet convolutions in fp16 channels-last (cuDNN tensor cores)"}
    w16 = mi(masked_n[:8], m4[:8])
    I.set_compute_dtype(mi, None)
    w32 = mi(masked_n[:8], m4[:8])
    out["config4_inpainting_b128_ndirs10_fp16convs"]["w_mat_rel_err_vs_fp32"] = ((w16 - w32).abs().max() / w32.abs().max()).item()
    I.set_compute_dtype(mi, "tc")
    ms = timeit(lambda: mi(masked_n, m4), reps=3, warm=2)
    wtc = mi(masked_n[:8], m4[:8])
    out["config4_inpainting_b128_ndirs10_tcgen05"] = {"ms_per_step": ms, "audio_s_per_s": B * 4.0 / (ms * 1e-3),
                                                      "note": "every 3x3 convolution on the in-house tcgen05 implicit-GEMM kernel (NHWC fp16, TMA halo, "
                                                              "two-tensor K loop instead of the decoder's cat); no cuDNN convolution in the step",
                                                      "w_mat_rel_err_vs_fp32": ((wtc - w32).abs().max() / w32.abs().max()).item()}
    I.set_compute_dtype(mi, None)

# config 4, training: one iteration of the inpainting NPPC trainer (nppc_trainer.py:146-154) at the shipped batch size — frozen
# restoration UNet on the tcgen05 convolutions, PC head in train mode through autograd (library convolutions), masking + real
# Gram-Schmidt + objective forward / backward on the kernels, clipped Adam.  Reported, never fatal (first added without GPU access).
try:
    I.set_compute_dtype(mi.pretrained_restoration_model, "tc")
    st = I.InpaintingNPPCStep(mi, 1.0, 500, max_grad_norm=1.0)
    st.step = 600
    opt_i = torch.optim.Adam(mi.pc_wrapper.parameters(), lr=1e-4, betas=(0.5, 0.999))
    batch = (spec * mask[:, None, None, :], mask, spec)
    ms = timeit(lambda: st.train_step(batch, opt_i), reps=3, warm=2)
    out["config4_inpainting_train_step_b128_ndirs10"] = {"ms_per_step": ms, "audio_s_per_s": B * 4.0 / (ms * 1e-3),
                                                         "note": "head UNet fwd/bwd on the library (fp32 autograd), everything else on the kernels"}
except Exception as e:   # noqa: BLE001
    out["config4_inpainting_train_step_b128_ndirs10"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
finally:
    I.set_compute_dtype(mi, None)
    mi.eval()
print(json.dumps(out, indent=1))
