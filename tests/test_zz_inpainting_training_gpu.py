"""a16 grad path on the GPU: InpaintingNPPCStep.base_step(requires_grad=True) / train_step (PC-head UNet in train mode through
torch autograd over the library convolutions; masking, real Gram-Schmidt and the objective — forward AND backward — on the
kernels) against ONE training step of the unmodified reference (tests/golden/inpaint_step_b2_grads.npz).

The host-side mathematics of this step is pinned on CPU (tests/test_inpainting.py::
test_inpainting_training_step_host_math_vs_reference_gradients); this file adds the kernel glue (scratch decode of the Gram /
coefficient matrices, the real linear combination riding on nppc_complex_lincomb).  It was written after the round's GPU
budget was spent and has not run on hardware yet — hence xfail(strict=False) and its place at the end of the suite: a
pass shows up as XPASS, a failure cannot mask the verified tests."""
import pytest
import torch

from conftest import load_golden, rel_err
from test_inpainting import N_DIRS, _check_step_against_reference, _product_model

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="authored without GPU access (round-2 budget spent); kernels involved are verified individually")]


def _batch(gd):
    return gd["masked_spec"], gd["mask"], gd["clean_spec"]


@pytest.mark.parametrize("step", [0, 600])
def test_inpainting_training_step_vs_reference_gradients(step):
    import generative_audio_b200 as g
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gd, gg = load_golden("inpaint_model_b2"), load_golden("inpaint_step_b2_grads")
    # (1) gradients of base_step(requires_grad=True), before any clipping
    m = _product_model()
    m.pc_wrapper.train()
    stepper = g.inpainting.InpaintingNPPCStep(m, 1.0, 500)
    stepper.step = step
    with torch.enable_grad():
        reconst, objective, log = stepper.base_step(_batch(gd), requires_grad=True)
        objective.backward()
    assert log["w_mat"].shape == (2, N_DIRS, 40, 50) and not log["w_mat"].requires_grad
    assert rel_err(reconst.detach().cpu(), gg[f"s{step}_reconst_err"]) < 2e-3
    assert rel_err(log["second_moment_mse"].cpu(), gg[f"s{step}_second_moment_mse"]) < 5e-3
    assert all(p.grad is None for p in m.pretrained_restoration_model.parameters())          # frozen restoration UNet
    before = {k: p.grad.detach().clone() for k, p in m.pc_wrapper.net.named_parameters()}
    # (2) the whole iteration on a fresh model: zero_grad / backward / clip_grad_norm_ / Adam (nppc_trainer.py:146-154)
    m2 = _product_model()
    stepper2 = g.inpainting.InpaintingNPPCStep(m2, 1.0, 500, max_grad_norm=1.0)
    stepper2.step = step
    opt = torch.optim.Adam(m2.parameters(), lr=1e-4, betas=(0.5, 0.999))
    with torch.enable_grad():
        obj2, log2 = stepper2.train_step(_batch(gd), opt)
    assert stepper2.step == step + 1 and m2.pc_wrapper.training and not m2.pretrained_restoration_model.training
    _check_step_against_reference(m2.pc_wrapper.net, gg, step, obj2.item(), log2["grad_norm"].item(), before, 2e-2)
    # (3) the no-grad statistics path agrees with the autograd path on the same (train-mode) head
    with torch.no_grad():
        _, obj3, _ = stepper.base_step(_batch(gd))
    assert abs(obj3.item() - objective.item()) < 1e-4 * abs(objective.item())


def test_reference_trainer_pattern_through_differentiable_forward():
    """`model.differentiable_forward = True`: w_mat = model(x, mask) keeps the graph (UNet autograd, MaskOutFn, GramSchmidtRealFn);
    the reference trainer's loss written in torch (oracle restatement, on the device) and backward reproduce its gradients."""
    import generative_audio_b200 as g
    import nppc_oracle as O
    from test_inpainting import PICKS
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gd, gg = load_golden("inpaint_model_b2"), load_golden("inpaint_step_b2_grads")
    m = _product_model()
    m.pc_wrapper.train()
    clean_n, m4, masked_n = g.inpainting.preprocess_data(gd["clean_spec"].cuda(), gd["masked_spec"].cuda(), gd["mask"].cuda())
    with torch.enable_grad():
        assert not m(masked_n, m4).requires_grad
        m.differentiable_forward = True
        w = m(masked_n, m4)
        assert w.requires_grad
        st = O.inpaint_loss(w, clean_n, m.get_pred_spec_mag_norm(masked_n, m4), step=600, grace=500, lambda0=1.0)
        st["objective"].backward()
    assert abs(st["objective"].item() - gg["s600_objective"].item()) < 2e-3 * abs(gg["s600_objective"].item())
    params = dict(m.pc_wrapper.net.named_parameters())
    for k, s in PICKS.items():
        assert rel_err(params[k].grad.flatten()[::s].cpu(), gg[f"s600_grad_{k}"]) < 2e-2, k
