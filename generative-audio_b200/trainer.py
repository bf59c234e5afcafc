"""NPPC-audio training step (a14): NPPCAudioTrainer.base_step (nppc_audio/trainer.py:234-317), _get_true_and_pred_crm
(:344-371) and _calculate_final_objective (:337-342).

`base_step(batch)` (no grad): everything on the B200 kernels; the frozen backbone and the noisy STFT run ONCE (the reference
runs the backbone twice and the noisy STFT three times); Gram-Schmidt + loss are two HBM passes in total.
`base_step(batch, requires_grad=True)` / `train_step`: same statistics with an autograd graph through the PC head — the
frozen half still runs on the hand-written kernels, the head's forward/backward on torch autograd (see training.py)."""
import torch

from . import ops, training
from .nppc_model import NPPCModel


def second_moment_lambda(step: int, grace: float, lambda0: float) -> float:
    return max(min(-1 + 2 * step / grace, 1), 1e-6) * lambda0


class NPPCAudioStep:
    """Holds what base_step reads from the reference trainer: model, step counter, loss hyper-parameters."""

    def __init__(self, nppc_model: NPPCModel, second_moment_loss_grace: float = 500, second_moment_loss_lambda: float = 1.0,
                 amp_dtype=None):
        self.nppc_model = nppc_model
        self.amp_dtype = amp_dtype   # e.g. torch.bfloat16: autocast for the head's GEMM-shaped ops (BASELINE config 3)
        self.step = 0
        self.second_moment_loss_grace = second_moment_loss_grace
        self.second_moment_loss_lambda = second_moment_loss_lambda

    def base_step(self, batch, requires_grad: bool = False):
        """batch = (noisy [B,L], clean [B,L]) -> (reconst_err [B], objective [], log dict with the reference's keys)."""
        if requires_grad:
            return self._base_step_autograd(batch)
        with torch.no_grad():
            return self._base_step_kernels(batch)

    def _frozen_half(self, noisy, clean):
        """Kernels only: STFTs, frozen backbone, enhanced spectrum (conj quirk), gt / pred cIRM after drop_band."""
        model = self.nppc_model
        G = model.audio_pc_wrapper.net.num_groups_in_drop_band
        c = model.config.stft_configuration
        with torch.no_grad():
            mag, real, imag = ops.stft_mri(noisy.to(model.device), c.nfft, c.hop_length, c.win_length)
            pred_crm = model.pretrained_restoration_model(mag, real, imag)
            emag, ereal, eimag = ops.crm_decompress_apply(pred_crm, real, imag, conj=True)
            _, cr, ci = ops.stft_mri(clean.to(model.device), c.nfft, c.hop_length, c.win_length)
            gt = ops.drop_band(ops.build_cirm(real[:, 0], imag[:, 0], cr[:, 0], ci[:, 0]), G)
            pred = ops.drop_band(pred_crm, G)
        return (mag, real, imag, emag[:, None], ereal[:, None], eimag[:, None]), gt, pred

    def _base_step_autograd(self, batch):
        model = self.nppc_model
        noisy, clean = batch
        feats, gt, pred = self._frozen_half(noisy, clean)
        lam = second_moment_lambda(self.step, self.second_moment_loss_grace, self.second_moment_loss_lambda)
        with torch.enable_grad():
            head = training.head_forward_autograd(model.audio_pc_wrapper.net, *feats, amp_dtype=self.amp_dtype)
            w_mat = training.gram_schmidt_autograd(head)
            st = training.nppc_loss_autograd(w_mat, gt, pred, lam)
        log = {"noisy_complex": noisy, "clean_complex": clean, "pred_crm": pred, "w_mat": w_mat.detach(),
               "err_norm": st["err_norm"].detach(), "err_proj": st["err_proj"].detach(),
               "err_proj_mag": st["err_proj_mag"].detach(), "w_norms": st["w_norms"].detach(),
               "reconst_err": st["reconst_err"].detach(), "second_moment_mse": st["second_moment_mse"].detach(),
               "objective": st["objective"].detach()}
        return st["reconst_err"], st["objective"], log

    def train_step(self, batch, optimizer):
        """zero_grad / backward / DP gradient all-reduce / optimizer.step (trainer.py:100-106) -> (objective, log)."""
        optimizer.zero_grad(set_to_none=True)
        _, objective, log = self.base_step(batch, requires_grad=True)
        objective.backward()
        training.allreduce_gradients(self.nppc_model.audio_pc_wrapper.parameters())
        optimizer.step()
        self.step += 1
        return objective.detach(), log

    def _base_step_kernels(self, batch):
        model = self.nppc_model
        noisy, clean = batch
        G = model.audio_pc_wrapper.net.num_groups_in_drop_band
        head, pred_crm = model.forward_stages(noisy)
        c = model.config.stft_configuration
        _, nr, ni = ops.stft_mri(noisy.to(model.device), c.nfft, c.hop_length, c.win_length)
        _, cr, ci = ops.stft_mri(clean.to(model.device), c.nfft, c.hop_length, c.win_length)
        gt = ops.drop_band(ops.build_cirm(nr[:, 0], ni[:, 0], cr[:, 0], ci[:, 0]), G)
        pred = ops.drop_band(pred_crm, G)
        w_mat, st = ops.gs_loss_fused(head, gt, pred)
        lam = second_moment_lambda(self.step, self.second_moment_loss_grace, self.second_moment_loss_lambda)
        objective = st["reconst_err"].mean() + lam * st["second_moment_mse"].mean()
        log = {
            "noisy_complex": noisy, "clean_complex": clean, "pred_crm": pred, "w_mat": w_mat,
            "err_norm": st["err_norm"], "err_proj": st["err_proj"], "err_proj_mag": st["err_proj"].abs(),
            "w_norms": st["w_norms"], "reconst_err": st["reconst_err"], "second_moment_mse": st["second_moment_mse"],
            "objective": objective,
        }
        return st["reconst_err"], objective, log
