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
from .checkpoint import CheckpointMixin
from .nppc_model import NPPCModel


def second_moment_lambda(step: int, grace: float, lambda0: float) -> float:
    return max(min(-1 + 2 * step / grace, 1), 1e-6) * lambda0


class NPPCAudioStep(CheckpointMixin):
    """Holds what base_step reads from the reference trainer: model, step counter, loss hyper-parameters."""

    def __init__(self, nppc_model: NPPCModel, second_moment_loss_grace: float = 500, second_moment_loss_lambda: float = 1.0,
                 amp_dtype=None):
        self.nppc_model = nppc_model
        self.amp_dtype = amp_dtype   # kept for API compatibility: the GEMM-shaped ops always run fp16-operand / fp32-accumulate
        self.step = 0
        self._reducer = None
        self._graph = None          # train_step_graphed: (CUDAGraph, static inputs, static outputs)
        self._lam_t = None          # loss weight as a DEVICE scalar, so that a captured step follows the schedule
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
        lam = self._lambda_tensor()
        with torch.enable_grad():
            head = training.head_forward_train(model.audio_pc_wrapper.net, *feats)
            objective, w_mat, err_norm, err_proj, w_norms, reconst_err, second_moment = training.GsLossFn.apply(head, gt, pred, lam)
        err_proj = torch.view_as_complex(err_proj)
        log = {"noisy_complex": noisy, "clean_complex": clean, "pred_crm": pred, "w_mat": w_mat, "err_norm": err_norm,
               "err_proj": err_proj, "err_proj_mag": err_proj.abs(), "w_norms": w_norms, "reconst_err": reconst_err,
               "second_moment_mse": second_moment, "objective": objective.detach()}
        return reconst_err, objective, log

    def _lambda_tensor(self):
        if self._lam_t is None:
            self._lam_t = torch.zeros((), device=self.nppc_model.device, dtype=torch.float64)
        if not torch.cuda.is_current_stream_capturing():
            self._lam_t.fill_(second_moment_lambda(self.step, self.second_moment_loss_grace, self.second_moment_loss_lambda))
        return self._lam_t

    def _step_body(self, batch, optimizer):
        optimizer.zero_grad(set_to_none=True)
        if self._reducer is not None:
            self._reducer.reset()
        _, objective, log = self.base_step(batch, requires_grad=True)
        objective.backward()
        if self._reducer is not None:
            self._reducer.finish()
        optimizer.step()
        return objective.detach(), log

    def train_step_graphed(self, batch, optimizer):
        """train_step with the WHOLE step (frozen half, head forward, hand-written backward, DP all-reduce, optimizer) captured
        in one CUDA graph: ~9000 kernel launches become one replay, which removes the host-side launch cost that otherwise
        bounds the step (measured 123 ms wall for 105 ms of kernels at B = 32).  The first call warms up and captures; later
        calls copy the batch into the static input buffers and replay.  Needs a capturable optimizer
        (torch.optim.Adam(..., capturable=True)); the loss weight lambda is a device scalar refreshed before every replay.
        No autograd graph of an earlier EAGER backward may still be alive when the capture starts (drop the old loss tensor):
        its AccumulateGrad nodes belong to the legacy stream and CUDA refuses the cross-stream dependency during capture."""
        import torch.distributed as dist
        noisy, clean = batch
        dev = self.nppc_model.device
        if self._reducer is None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self._reducer = training.GradBucketReducer(self.nppc_model.audio_pc_wrapper.parameters())
        if self._graph is None or self._graph[1][0].shape != noisy.shape:
            static_in = (noisy.to(dev).clone(), clean.to(dev).clone())
            self._lambda_tensor()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):       # warm-up: plans, workspaces, optimizer state, allocator pools
                    self._step_body(static_in, optimizer)
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._step_body(static_in, optimizer)
            self._graph = (graph, static_in, static_out)
        graph, static_in, static_out = self._graph
        static_in[0].copy_(noisy, non_blocking=True)
        static_in[1].copy_(clean, non_blocking=True)
        self._lambda_tensor()
        graph.replay()
        self.step += 1
        return static_out

    def train_step(self, batch, optimizer):
        """zero_grad / backward (+ overlapped DP gradient all-reduce) / optimizer.step (trainer.py:100-106) -> (objective, log)."""
        if self._reducer is None:
            self._reducer = training.GradBucketReducer(self.nppc_model.audio_pc_wrapper.parameters())
        out = self._step_body(batch, optimizer)
        self.step += 1
        return out

    def _base_step_kernels(self, batch):
        model = self.nppc_model
        noisy, clean = batch
        G = model.audio_pc_wrapper.net.num_groups_in_drop_band
        taps = {}
        head, pred_crm = model.forward_stages(noisy, taps)
        c = model.config.stft_configuration
        nr, ni = taps["real"], taps["imag"]          # the noisy STFT forward_stages already computed (one STFT, not two)
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
