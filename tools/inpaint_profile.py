"""Top CUDA kernels of BASELINE config 4 (inpainting, B = 128, n_dirs = 10) with the UNet convolutions on tcgen05."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import generative_audio_b200 as g
import weights
torch.set_grad_enabled(False)
I = g.inpainting
N_DIRS, B = 10, int(sys.argv[1]) if len(sys.argv) > 1 else 128
rest = I.UNet(I.UNetConfig(in_channels=1, out_channels=1))
shp = lambda mod: [(k, tuple(v.shape)) for k, v in mod.state_dict().items()]
rest.load_state_dict(weights.synth_unet_state_dict(shp(rest), 0, "rest."))
ck = os.path.join(tempfile.mkdtemp(), "rest.pt")
torch.save({"model_state_dict": rest.state_dict()}, ck)
cfg = I.NPPCModelConfig(pretrained_restoration_model_configuration=I.UNetConfig(in_channels=1, out_channels=1), pretrained_restoration_model_path=ck,
                        audio_pc_wrapper_configuration=I.AudioInpaintingPCWrapperConfig(model_configuration=I.UNetConfig(in_channels=2, out_channels=N_DIRS), n_dirs=N_DIRS))
mi = I.NPPCModel(cfg)
mi.pc_wrapper.net.load_state_dict(weights.synth_unet_state_dict(shp(mi.pc_wrapper.net), 0, "head."))
spec = torch.randn(B, 2, 128, 500, device="cuda")
mask = torch.ones(B, 500, device="cuda"); mask[:, 200:218] = 0
clean_n, m4, masked_n = I.preprocess_data(spec, spec * mask[:, None, None, :], mask)
I.set_compute_dtype(mi, "tc")
for _ in range(2):
    mi(masked_n, m4)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    mi(masked_n, m4)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=90))
