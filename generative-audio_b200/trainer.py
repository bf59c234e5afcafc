"""NPPC-audio training step (a14): NPPCAudioTrainer.base_step (nppc_audio/trainer.py:234-317), _get_true_and_pred_crm
(:344-371) and _calculate_final_objective (:337-342).

`base_step(batch)` (no grad): everything on the B200 kernels; the frozen backbone and the noisy STFT run ONCE (the reference
runs the backbone twice and the noisy STFT three times); Gram-Schmidt + loss are two HBM passes in total.
`base_step(batch, requires_grad=True)` / `train_step`: same statistics with an autograd graph through the PC head whose heavy
nodes are hand-written (training.py): stepwise tcgen05 LSTM forward + BPTT, tcgen05 GEMMs for every 1x1 convolution (forward,
dX and dW), a coefficient-space Gram-Schmidt + objective backward; no nn.LSTM / F.conv1d launches.  Data parallel: gradients
are all-reduced bucket by bucket while the backward is still running (training.GradBucketReducer)."""
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
        self.amp_dtype = amp_dtype   # kept for API compatibility: the GEMM-shaped ops always run fp16-operand / fp32-accumulate
        self.step = 0
        self._reducer = None
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
            head = training.head_forward_train(model.audio_pc_wrapper.net, *feats)
            objective, w_mat, err_norm, err_proj, w_norms, reconst_err, second_moment = training.GsLossFn.apply(head, gt, pred, lam)
        err_proj = torch.view_as_complex(err_proj)
        log = {"noisy_complex": noisy, "clean_complex": clean, "pred_crm": pred, "w_mat": w_mat, "err_norm": err_norm,
               "err_proj": err_proj, "err_proj_mag": err_proj.abs(), "w_norms": w_norms, "reconst_err": reconst_err,
               "second_moment_mse": second_moment, "objective": objective.detach()}
        return reconst_err, objective, log

    def train_step(self, batch, optimizer):
        """zero_grad / backward (+ overlapped DP gradient all-reduce) / optimizer.step (trainer.py:100-106) -> (objective, log)."""
        if self._reducer is None:
            self._reducer = training.GradBucketReducer(self.nppc_model.audio_pc_wrapper.parameters())
        optimizer.zero_grad(set_to_none=True)
        self._reducer.reset()
        _, objective, log = self.base_step(batch, requires_grad=True)
        objective.backward()
        self._reducer.finish()
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
