"""NPPCModel.differentiable_forward (opt-in): the reference trainer's own pattern — `w_mat = nppc_model(noisy)`, the loss
written in torch, `.backward()` (nppc_audio/trainer.py:255-298,100-106) — on the product model: head forward + backward on the
hand-written training Functions, Gram-Schmidt through training.GramSchmidtFn (backward for an arbitrary upstream gradient),
gradients against the unmodified reference's CPU autograd (tests/golden/model_step_g2_b4_grads.npz).

The coefficient-space mathematics and the Function's plumbing are pinned on CPU (tests/test_training_math_cpu.py::
test_gram_schmidt_backward_for_an_arbitrary_upstream_gradient, ::test_differentiable_gram_schmidt_function_glue_on_cpu); this
file adds the kernel side (scratch decode, Gram pass over the stacked [x; g], linear combination).  Written after the round's
GPU budget was spent — not yet run on hardware, hence xfail(strict=False) and its place at the end of the suite."""
import pytest
import torch

import nppc_oracle as O
from conftest import load_golden, rel_err
from helpers import build_model

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="authored without GPU access (round-2 budget spent); kernels involved are verified individually")]


@pytest.mark.parametrize("step", [0, 600])
def test_reference_trainer_pattern_through_differentiable_forward(step):
    import generative_audio_b200 as G
    ops = G.ops
    g, gg = load_golden("model_step_g2_b4"), load_golden("model_step_g2_b4_grads")
    m, _ = build_model(5, 2, "f32")
    noisy, clean = g["noisy"].cuda(), g["clean"].cuda()
    c = m.config.stft_configuration
    with torch.no_grad():
        w0 = m(noisy)                                                   # the inference path (no graph)
        _, nr, ni = ops.stft_mri(noisy, c.nfft, c.hop_length, c.win_length)
        _, cr, ci = ops.stft_mri(clean, c.nfft, c.hop_length, c.win_length)
        gt = ops.drop_band(ops.build_cirm(nr[:, 0], ni[:, 0], cr[:, 0], ci[:, 0]), 2)
        pred = ops.drop_band(m.get_pred_crm(noisy), 2)
    m.differentiable_forward = True
    try:
        m.zero_grad(set_to_none=True)
        with torch.enable_grad():
            w = m(noisy)
            assert w.requires_grad and w.shape == w0.shape
            out = O.nppc_loss(w, gt, pred, step=step, grace=500, lambda0=1.0)      # trainer.py:259-298 in plain torch ops
            out["objective"].backward()
        with torch.no_grad():
            assert not m(noisy).requires_grad                                       # grad mode off: the inference path again
    finally:
        m.differentiable_forward = False
    assert rel_err(w.detach().cpu(), w0.cpu()) < 5e-2                                # sanity only: fp16-operand training head vs the f32 path
    assert abs(out["objective"].item() - gg[f"s{step}_objective"].item()) < 2e-3
    net = m.audio_pc_wrapper.net
    assert all(p.grad is None for p in m.pretrained_restoration_model.parameters())
    picks = {"sb_fc_w": net.sb_model.fc_output_layer.weight, "sb_fc_b": net.sb_model.fc_output_layer.bias,
             "lstm_b_hh_l1": net.sb_model.sequence_model.bias_hh_l1, "lstm_w_ih_l0": net.sb_model.sequence_model.weight_ih_l0,
             "tsse_fcat_w": net.channel_attention.feature_concate_fc.weight,
             "tcn0_prelu1": net.fb_model.sequence_model[0].prelu1.weight,
             "tcn7_norm2_w": net.fb_model_imag.sequence_model[7].norm2.weight,
             "fb_fc_b": net.fb_model_real.fc_output_layer.bias}
    for k, p in picks.items():
        assert rel_err(p.grad.cpu(), gg[f"s{step}_{k}"]) < 2e-2, k
    m.zero_grad(set_to_none=True)
