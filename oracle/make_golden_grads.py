"""Adds tests/golden/model_step_g2_b4_grads.npz: gradients of the reference NPPCAudioTrainer.base_step objective w.r.t. a few
PC-head parameters (unmodified reference, CPU autograd) for the same inputs/weights as model_step_g2_b4.  Dev container only."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

ns = ref_loader.load_reference(trainer=True)
import make_golden as MG  # noqa: E402  (re-uses build_model; its main() is not run)

g = np.load(os.path.join(HERE, "..", "tests", "golden", "model_step_g2_b4.npz"))
noisy, clean = torch.from_numpy(g["noisy"]), torch.from_numpy(g["clean"])
torch.set_grad_enabled(True)
m2, cfg2 = MG.build_model(5, 2)
m2.train(False)
out = {}
for step in (0, 600):
    stub = types.SimpleNamespace(nppc_model=m2, device="cpu", step=step,
                                 config=types.SimpleNamespace(nppc_model_configuration=cfg2, second_moment_loss_grace=500,
                                                              second_moment_loss_lambda=1.0))
    T = ns.NPPCAudioTrainer
    stub._get_true_and_pred_crm = types.MethodType(T._get_true_and_pred_crm, stub)
    stub._calculate_final_objective = types.MethodType(T._calculate_final_objective, stub)
    m2.zero_grad()
    reconst, obj, log = T.base_step(stub, (noisy, clean))
    obj.backward()
    net = m2.audio_pc_wrapper.net
    picks = {"sb_fc_w": net.sb_model.fc_output_layer.weight, "sb_fc_b": net.sb_model.fc_output_layer.bias,
             "lstm_b_hh_l1": net.sb_model.sequence_model.bias_hh_l1, "lstm_w_ih_l0": net.sb_model.sequence_model.weight_ih_l0,
             "tsse_fcat_w": net.channel_attention.feature_concate_fc.weight,
             "tcn0_prelu1": net.fb_model.sequence_model[0].prelu1.weight,
             "tcn7_norm2_w": net.fb_model_imag.sequence_model[7].norm2.weight,
             "fb_fc_b": net.fb_model_real.fc_output_layer.bias}
    for k, p in picks.items():
        out[f"s{step}_{k}"] = p.grad.detach().numpy().copy()
    out[f"s{step}_objective"] = obj.detach().numpy()
    # global grad norm over the head
    out[f"s{step}_head_grad_norm"] = np.array(torch.sqrt(sum((p.grad ** 2).sum() for p in net.parameters() if p.grad is not None)).item())
    out[f"s{step}_backbone_has_grad"] = np.array(any(p.grad is not None for p in m2.pretrained_restoration_model.parameters()))
np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "model_step_g2_b4_grads.npz"), **out)
print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})
