"""Adds tests/golden/inpaint_step_b2_grads.npz: ONE training step of the UNMODIFIED reference inpainting NPPC trainer
(nppc_audio/inpainting/trainer/nppc_trainer.py:338-385 base_step, :146-154 zero_grad / backward / clip_grad_norm_ / step) on
CPU, for the inputs / weights of inpaint_model_b2.npz.  The PC head's UNet is in TRAIN mode (BatchNorm batch statistics, as in
the reference: only the restoration UNet is put into eval mode, inpainting/nppc/nppc_model.py:112), the restoration UNet is
frozen under no_grad.  Stored per schedule point (step 0 and 600): objective, reconst_err, second_moment_mse, gradients of a few
head parameters (strided samples of the large ones), the global gradient norm before clipping, one BatchNorm running mean
after the forward, and the same parameters after the clipped Adam step.  Dev container only:
    python oracle/make_golden_inpainting_grads.py
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLD = os.path.join(HERE, "..", "tests", "golden")

import ref_loader  # noqa: E402
import weights  # noqa: E402

ref_loader.install_shims(trainer=True)
if "wandb" not in sys.modules:
    try:
        import wandb  # noqa: F401
    except Exception:
        sys.modules["wandb"] = types.ModuleType("wandb")
from nppc_audio.inpainting.networks.unet import UNet, UNetConfig  # noqa: E402
from nppc_audio.inpainting.nppc.nppc_model import NPPCModel, NPPCModelConfig  # noqa: E402
from nppc_audio.inpainting.nppc.pc_wrapper import AudioInpaintingPCWrapperConfig  # noqa: E402
from nppc_audio.inpainting.trainer.nppc_trainer import NPPCAudioInpaintingTrainer  # noqa: E402

N_DIRS = 3
MAX_GRAD_NORM = 1.0                       # NPPCAudioInpaintingTrainerConfig.max_grad_norm
ADAM = dict(lr=1e-4, betas=(0.5, 0.999))  # scripts/train/config/config_nppc.yaml:56-60
# parameter name -> flat stride of the stored sample (1 = the whole tensor)
PICKS = {"outc.conv.weight": 1, "outc.conv.bias": 1, "inc.conv.conv.0.weight": 1, "inc.conv.conv.1.weight": 1,
         "inc.conv.conv.4.bias": 1, "down4.mpconv.1.conv.3.weight": 4099, "down4.mpconv.1.conv.4.weight": 1,
         "up1.conv.conv.0.weight": 8191, "up4.conv.conv.4.bias": 1}


def shapes(m):
    return [(k, tuple(v.shape)) for k, v in m.state_dict().items()]


def build():
    rest = UNet(UNetConfig(in_channels=1, out_channels=1))
    rest.load_state_dict(weights.synth_unet_state_dict(shapes(rest), 0, "rest."))
    ck = os.path.join(tempfile.mkdtemp(), "rest.pt")
    torch.save({"model_state_dict": rest.state_dict()}, ck)
    cfg = NPPCModelConfig(pretrained_restoration_model_configuration=UNetConfig(in_channels=1, out_channels=1),
                          pretrained_restoration_model_path=ck,
                          audio_pc_wrapper_configuration=AudioInpaintingPCWrapperConfig(
                              model_configuration=UNetConfig(in_channels=2, out_channels=N_DIRS), n_dirs=N_DIRS),
                          device="cpu")
    model = NPPCModel(cfg)
    model.pc_wrapper.net.load_state_dict(weights.synth_unet_state_dict(shapes(model.pc_wrapper.net), 0, "head."))
    assert model.pc_wrapper.training and not model.pretrained_restoration_model.training     # the reference's modes, untouched
    return model


def main():
    g = np.load(os.path.join(GOLD, "inpaint_model_b2.npz"))
    masked_spec, mask, clean_spec = (torch.from_numpy(g[k]) for k in ("masked_spec", "mask", "clean_spec"))
    out = {}
    for step in (0, 600):
        model = build()
        opt = torch.optim.Adam(model.parameters(), **ADAM)
        stub = types.SimpleNamespace(nppc_model=model, step=step,
                                     config=types.SimpleNamespace(second_moment_loss_grace=500, second_moment_loss_lambda=1.0))
        stub._calculate_final_objective = types.MethodType(NPPCAudioInpaintingTrainer._calculate_final_objective, stub)
        with torch.enable_grad():
            reconst, objective, log = NPPCAudioInpaintingTrainer.base_step(stub, (masked_spec, mask, clean_spec))
            opt.zero_grad()
            objective.backward()
        net = model.pc_wrapper.net
        params = dict(net.named_parameters())
        for k, s in PICKS.items():
            out[f"s{step}_grad_{k}"] = params[k].grad.detach().flatten()[::s].numpy().copy()
        norm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=MAX_GRAD_NORM)
        out[f"s{step}_grad_norm"] = np.array(norm.item())
        opt.step()
        for k, s in PICKS.items():
            out[f"s{step}_after_{k}"] = params[k].detach().flatten()[::s].numpy().copy()
        out[f"s{step}_objective"] = objective.detach().reshape(1).numpy()
        out[f"s{step}_reconst_err"] = reconst.detach().numpy()
        out[f"s{step}_second_moment_mse"] = log["second_moment_mse"].numpy()
        out[f"s{step}_w_mat_absmax"] = np.array(log["w_mat"].abs().max().item())
        out[f"s{step}_bn_running_mean"] = dict(net.named_buffers())["inc.conv.conv.1.running_mean"].numpy().copy()
        out[f"s{step}_rest_has_grad"] = np.array(any(p.grad is not None for p in model.pretrained_restoration_model.parameters()))
    np.savez_compressed(os.path.join(GOLD, "inpaint_step_b2_grads.npz"), **out)
    print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})


if __name__ == "__main__":
    main()
