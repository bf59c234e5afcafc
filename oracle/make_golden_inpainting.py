"""Generates tests/golden/inpaint_* by running the UNMODIFIED reference inpainting NPPC model
(/root/reference/nppc_audio/inpainting) on CPU with deterministic synthetic weights.  Dev container only:
    python oracle/make_golden_inpainting.py
"""
import json
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
import utils as ref_utils  # noqa: E402
from nppc_audio.inpainting.networks.unet import UNet, UNetConfig  # noqa: E402
from nppc_audio.inpainting.nppc.nppc_model import NPPCModel, NPPCModelConfig  # noqa: E402
from nppc_audio.inpainting.nppc.pc_wrapper import AudioInpaintingPCWrapperConfig  # noqa: E402
from nppc_audio.inpainting.trainer.nppc_trainer import NPPCAudioInpaintingTrainer  # noqa: E402

torch.set_grad_enabled(False)
N_DIRS, B, Fq, T, GAP = 3, 2, 40, 50, 9


def shapes(m):
    return [(k, tuple(v.shape)) for k, v in m.state_dict().items()]


def main():
    rest = UNet(UNetConfig(in_channels=1, out_channels=1))
    with open(os.path.join(GOLD, "unet_manifest.json"), "w") as f:
        json.dump({"source": "reference UNet(in=1,out=1).state_dict() keys/shapes (nppc_audio/inpainting/networks/unet.py:247-262)",
                   "entries": [[k, list(s)] for k, s in shapes(rest)]}, f)
    rest.load_state_dict(weights.synth_unet_state_dict(shapes(rest), 0, "rest."))
    tmp = tempfile.mkdtemp()
    ck = os.path.join(tmp, "rest.pt")
    torch.save({"model_state_dict": rest.state_dict()}, ck)
    cfg = NPPCModelConfig(pretrained_restoration_model_configuration=UNetConfig(in_channels=1, out_channels=1),
                          pretrained_restoration_model_path=ck,
                          audio_pc_wrapper_configuration=AudioInpaintingPCWrapperConfig(
                              model_configuration=UNetConfig(in_channels=2, out_channels=N_DIRS), n_dirs=N_DIRS),
                          device="cpu")
    model = NPPCModel(cfg)
    model.pc_wrapper.net.load_state_dict(weights.synth_unet_state_dict(shapes(model.pc_wrapper.net), 0, "head."))
    model.eval()

    rng = np.random.Generator(np.random.PCG64(123))
    clean_spec = torch.from_numpy(rng.standard_normal((B, 2, Fq, T)).astype(np.float32))
    mask = torch.ones(B, T)
    for b, off in enumerate((11, 30)):
        mask[b, off:off + GAP] = 0
    masked_spec = clean_spec * mask[:, None, None, :]
    clean_n, m4, masked_n = ref_utils.preprocess_data(clean_spec, masked_spec, mask)
    pred = model.get_pred_spec_mag_norm(masked_n, m4)
    w_mat = model(masked_n, m4)
    out = dict(clean_spec=clean_spec, masked_spec=masked_spec, mask=mask, clean_n=clean_n, masked_n=masked_n, pred=pred,
               w_mat=w_mat)
    # base_step of the reference trainer, unbound, at three points of the lambda schedule
    for step in (0, 300, 600):
        stub = types.SimpleNamespace(nppc_model=model, step=step,
                                     config=types.SimpleNamespace(second_moment_loss_grace=500, second_moment_loss_lambda=1.0))
        stub._calculate_final_objective = types.MethodType(NPPCAudioInpaintingTrainer._calculate_final_objective, stub)
        _, objective, log = NPPCAudioInpaintingTrainer.base_step(stub, (masked_spec, mask, clean_spec))
        for k in ("err_norm", "err_proj", "w_norms", "reconst_err", "second_moment_mse"):
            out[f"s{step}_{k}"] = log[k]
        out[f"s{step}_objective"] = objective.reshape(1)
    np.savez_compressed(os.path.join(GOLD, "inpaint_model_b2.npz"), **{k: v.numpy() for k, v in out.items()})
    print({k: tuple(v.shape) for k, v in out.items()})
    print("w_mat norms", w_mat.flatten(2).norm(dim=2))


if __name__ == "__main__":
    main()
