"""Training step of the inpainting NPPC variant (a16): NPPCAudioInpaintingTrainer.base_step
(nppc_audio/inpainting/trainer/nppc_trainer.py:338-385) with gradients, and the loop body of train() (:146-154: zero_grad,
backward, clip_grad_norm_(max_grad_norm), optimizer.step).

What runs where:
  * log-magnitude normalisation, the frozen restoration UNet (eval mode, folded BatchNorm; tcgen05 convolutions when
    set_compute_dtype(model.pretrained_restoration_model, "tc")) and the mask blend: the inference kernels, under no_grad —
    ONE restoration pass shared by the head's input and the error (the reference runs it twice);
  * the PC head's UNet in TRAIN mode (BatchNorm batch statistics + running-stat update, as the reference leaves it:
    only the restoration UNet is put into eval mode, inpainting/nppc/nppc_model.py:112): torch autograd over the library's
    convolution / batch-norm forward and backward (SURVEY.md §2 row 10 allows the library for the UNet convolutions; the
    tcgen05 implicit-GEMM kernel of conv_tc.cu has no dgrad / wgrad yet);
  * `* (1 - mask)`, real Gram-Schmidt (detached normalisers, pc_wrapper.py:43-59), the projection / second-moment objective
    and ALL of their backward: hand-written — `nppc_mask_blend`, `nppc_gs_loss_fused_real` (two HBM passes), the
    coefficient-space solve of gs_backward.py on the Gram matrix the forward left in its scratch, and one streaming
    `nppc_complex_lincomb` pass (ops.real_lincomb)."""
import torch

from . import ops
from .gs_backward import gs_coeffs_from_gram, gs_grad_coeffs, gs_loss_grad_coeffs


class MaskOutFn(torch.autograd.Function):
    """x [B,C,F,T] * (1 - mask [B,1,F,T])  (AudioInpaintingPCWrapper.forward, pc_wrapper.py:78-82); the backward is the same
    kernel applied to the incoming gradient."""

    @staticmethod
    def forward(ctx, x, mask):
        ctx.save_for_backward(mask)
        return ops.mask_blend(None, x, mask)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return ops.mask_blend(None, g.contiguous(), mask), None


class GsLossRealFn(torch.autograd.Function):
    """head [B, n, F, T] (already zero outside the gap) -> objective, with w_mat and the statistics of the reference's log dict
    as non-differentiable outputs.  gt = clean log-magnitude, pred = the frozen restoration's output: neither carries a
    gradient (get_pred_spec_mag_norm runs under no_grad, nppc_model.py:157-158)."""

    @staticmethod
    def forward(ctx, head, gt, pred, lam):
        w, st, G, _ = ops.gs_loss_fused_real_with_gram(head, gt, pred)
        lam = torch.as_tensor(lam, dtype=torch.float64, device=head.device)
        objective = st["reconst_err"].mean() + (lam * st["second_moment_mse"].double().mean()).float()
        ctx.save_for_backward(head, gt, pred, G, lam)
        outs = (w, st["err_norm"], st["err_proj"], st["w_norms"], st["reconst_err"], st["second_moment_mse"])
        ctx.mark_non_differentiable(*outs)
        return (objective, *outs)

    @staticmethod
    def backward(ctx, g_obj, *unused):
        head, gt, pred, G, lam = ctx.saved_tensors
        A = gs_coeffs_from_gram(G, head.shape[1])        # fp64 replay from the fp64 Gram matrix (the scratch holds A in fp32 only)
        coef = gs_loss_grad_coeffs(G, A, lam, real=True) * g_obj.double()
        return ops.real_lincomb(head, gt, pred, coef), None, None, None


class GramSchmidtRealFn(torch.autograd.Function):
    """gram_schmidt_to_spec_mag (inpainting/nppc/pc_wrapper.py:43-59) with a backward for an ARBITRARY upstream gradient (a loss
    written in torch by the caller, as the reference trainer does): x [B, n, F, T] -> w_mat, n <= 6.  Same construction as
    training.GramSchmidtFn on real numbers: one Gram pass over the stacked (x; g), gs_grad_coeffs, one linear combination."""

    @staticmethod
    def forward(ctx, x):
        if x.shape[1] > 6:
            raise NotImplementedError("differentiable Gram-Schmidt: n_dirs <= 6 (the backward stacks 2 n vectors; kernels take 12)")
        ctx.save_for_backward(x)
        return ops.gram_schmidt_real(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        B, n = x.shape[:2]
        S = torch.cat([x, g.to(x.dtype)], dim=1).contiguous()
        G2 = ops.gram_matrix_real(S)
        coef = gs_grad_coeffs(G2, gs_coeffs_from_gram(G2, n))
        full = torch.zeros(B, 2 * n, 2 * n + 1, dtype=coef.dtype, device=coef.device)
        full[:, :n, :2 * n] = coef
        zero = torch.zeros(B, *x.shape[2:], device=x.device, dtype=x.dtype)
        return ops.real_lincomb(S, zero, zero, full)[:, :n].contiguous()


def head_forward_train(pc_wrapper, mag_spec: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """AudioInpaintingPCWrapper.forward up to (not including) Gram-Schmidt, with an autograd graph: UNet(2 -> n_dirs) in its
    current mode, then the masking kernel."""
    return MaskOutFn.apply(pc_wrapper.net(mag_spec), mask)


def clip_grad_norm_(parameters, max_norm: float, eps: float = 1e-6) -> torch.Tensor:
    """torch.nn.utils.clip_grad_norm_ semantics (L2, clip coefficient max_norm / (norm + 1e-6) clamped to 1) without a host
    sync: every gradient is scaled by a device scalar, so the step can be captured / enqueued ahead."""
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.zeros(())
    norm = torch.linalg.vector_norm(torch.stack(torch._foreach_norm(grads)))
    coef = torch.clamp(max_norm / (norm + eps), max=1.0)
    torch._foreach_mul_(grads, coef)
    return norm
