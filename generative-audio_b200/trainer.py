"""NPPC-audio training-step statistics (a14): forward of NPPCAudioTrainer.base_step (nppc_audio/trainer.py:234-317),
_get_true_and_pred_crm (:344-371) and _calculate_final_objective (:337-342) on the B200 kernels.

Compared with the reference step the frozen backbone and the noisy STFT run ONCE (the reference runs the backbone
twice and the noisy STFT three times) and Gram-Schmidt + loss are two HBM passes in total.
Round-1 scope: forward statistics (reconst_err, objective, log dict); the hand-written backward of the PC head is
not built yet, so `objective` carries no autograd graph."""
import torch

from . import ops
from .nppc_model import NPPCModel


def second_moment_lambda(step: int, grace: float, lambda0: float) -> float:
    return max(min(-1 + 2 * step / grace, 1), 1e-6) * lambda0


class NPPCAudioStep:
    """Holds what base_step reads from the reference trainer: model, step counter, loss hyper-parameters."""

    def __init__(self, nppc_model: NPPCModel, second_moment_loss_grace: float = 500, second_moment_loss_lambda: float = 1.0):
        self.nppc_model = nppc_model
        self.step = 0
        self.second_moment_loss_grace = second_moment_loss_grace
        self.second_moment_loss_lambda = second_moment_loss_lambda

    @torch.no_grad()
    def base_step(self, batch):
        """batch = (noisy [B,L], clean [B,L]) -> (reconst_err [B], objective [], log dict with the reference's keys)."""
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
