"""a16 — the NPPC audio *inpainting* variant (masked-gap restoration) behind the reference's call surface:

    UNetConfig / UNet                    nppc_audio/inpainting/networks/unet.py:117-290 (+ tmp_utils.py:8-101)
    RestorationWrapper                   nppc_audio/inpainting/networks/unet.py:293-313
    AudioInpaintingPCWrapper(Config)     nppc_audio/inpainting/nppc/pc_wrapper.py:61-88
    NPPCModel(Config) (inpainting)       nppc_audio/inpainting/nppc/nppc_model.py:22-159
    preprocess_data                      utils.py:294-306
    InpaintingNPPCStep.base_step         nppc_audio/inpainting/trainer/nppc_trainer.py:338-385 (forward statistics)

Same module tree / state_dict keys as the reference (`inc.conv.conv.0.weight`, `down1.mpconv.1.conv.4.running_var`, ...),
so reference checkpoints (`{"model_state_dict": ...}`) load with strict=True.  CUDA only.  Eval mode (inference, the frozen
restoration UNet): BatchNorm folded into the convolution weights, dropout off, 3x3 convolutions on the library (fp32 parity
path) or on the in-house tcgen05 implicit GEMM (set_compute_dtype(model, "tc"), row N4); everything around them is
hand-written: log-magnitude normalisation, mask blending, the real Gram-Schmidt and the projection / second-moment loss (two
HBM passes, csrc/gram_schmidt.cu).  Train mode (the PC head during InpaintingNPPCStep.train_step): the module tree as written
(BatchNorm batch statistics, dropout) through torch autograd, with the masking, Gram-Schmidt and objective — forward and
backward — on the kernels (inpainting_training.py)."""
from pathlib import Path
from typing import Literal, Optional

import pydantic
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .checkpoint import CheckpointMixin
from .modules import _NoDerivedState


class UNetConfig(pydantic.BaseModel):
    in_channels: int = 1
    out_channels: int = 1
    dropout: float = 0.0


def _double_conv(in_ch, out_ch, dropout=0.0):
    """(conv3x3 -> BN -> LeakyReLU(0.2)) x 2 [-> Dropout]; index positions match tmp_utils.double_conv's Sequential."""
    layers = [nn.Conv2d(in_ch, out_ch, 3, padding=1), nn.BatchNorm2d(out_ch), nn.LeakyReLU(0.2),
              nn.Conv2d(out_ch, out_ch, 3, padding=1), nn.BatchNorm2d(out_ch), nn.LeakyReLU(0.2)]
    if dropout:
        layers.append(nn.Dropout(dropout))
    return nn.Sequential(*layers)


class _DoubleConv(_NoDerivedState, nn.Module):
    def __init__(self, in_ch, out_ch, dropout=0.0):
        super().__init__()
        self.conv = _double_conv(in_ch, out_ch, dropout)
        self._fold = None
        self._fold_key = None

    def _folded(self):
        """Eval-mode BatchNorm folded into the preceding conv: w' = w * g / sqrt(var + eps), b' = (b - mean) * g / sqrt(..) + beta.
        Derived cache keyed on the parameter versions; never serialised."""
        ps = [p for m in (self.conv[0], self.conv[1], self.conv[3], self.conv[4]) for p in list(m.parameters()) + list(m.buffers())]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._fold is None or key != self._fold_key:
            out = []
            with torch.no_grad():
                for ci, bi in ((0, 1), (3, 4)):
                    conv, bn = self.conv[ci], self.conv[bi]
                    sc = (bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps))
                    w = (conv.weight.double() * sc[:, None, None, None]).float().contiguous()
                    b = ((conv.bias.double() - bn.running_mean.double()) * sc + bn.bias.double()).float().contiguous()
                    out.append((w, b))
            self._fold, self._fold_key = out, key
        return self._fold

    mc_dropout = False   # MC-dropout inference (utils.enable_dropout, utils.py:334-338): Dropout layers stay active in eval
    compute_dtype = None  # None = fp32 (parity path, cuDNN); torch.float16 / bfloat16 = cuDNN tensor-core convolutions;
    #                       "tc" = the in-house tcgen05 implicit-GEMM convolutions on NHWC fp16 (row N4, set_compute_dtype("tc"))

    def forward_tc(self, x0, x1=None):
        """x0 [B,H,W,C0p] (and x1: the decoder's cat([skip, up]) is two K-loop segments, never materialised) -> [B,H,W,Cout] fp16."""
        (w0, b0), (w1, b1) = self._folded()
        C_in = w0.shape[1]
        C1 = 0 if x1 is None else x1.shape[3]
        C0 = C_in - C1
        key = (self._fold_key, C0, C1)
        if getattr(self, "_tc_key", None) != key:
            self._tc = (ops.conv3x3_pack_weights(w0, C0, C1), b0, ops.conv3x3_pack_weights(w1, w1.shape[1], 0), b1)
            self._tc_key = key
        p0, b0, p1, b1 = self._tc
        y = ops.conv3x3_tc(x0, x1, p0, b0, 0.2)
        y = ops.conv3x3_tc(y, None, p1, b1, 0.2)
        if self.mc_dropout and len(self.conv) > 6:
            y = F.dropout(y, self.conv[6].p, training=True)
        return y

    def _tc(self):
        return self.compute_dtype == "tc" and not self.training

    def forward(self, x):
        if self.training:   # BatchNorm batch statistics (+ running-stat update) and dropout as written: autograd over the library
            return self.conv(x)
        if self.compute_dtype == "tc":
            return self.forward_tc(*x) if isinstance(x, tuple) else self.forward_tc(x)
        (w0, b0), (w1, b1) = self._folded()
        if self.compute_dtype is not None:
            dt = self.compute_dtype
            if getattr(self, "_fold16_key", None) != (self._fold_key, dt):
                self._fold16 = [(w.to(dt).contiguous(memory_format=torch.channels_last), b.to(dt)) for w, b in self._fold]
                self._fold16_key = (self._fold_key, dt)
            (w0, b0), (w1, b1) = self._fold16
            x = x.to(dt).contiguous(memory_format=torch.channels_last)
        x = F.leaky_relu(F.conv2d(x, w0, b0, padding=1), 0.2, inplace=True)
        x = F.leaky_relu(F.conv2d(x, w1, b1, padding=1), 0.2, inplace=True)
        if self.mc_dropout and len(self.conv) > 6:
            x = F.dropout(x, self.conv[6].p, training=True)
        return x


class _InConv(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv = _DoubleConv(in_ch, out_ch)

    def forward(self, x):
        return self.conv(x)


class _Down(nn.Module):
    def __init__(self, in_ch, out_ch, dropout=0.0):
        super().__init__()
        self.mpconv = nn.Sequential(nn.MaxPool2d(2), _DoubleConv(in_ch, out_ch, dropout))

    def forward(self, x):
        if self.mpconv[1]._tc():   # NHWC fp16 in and out
            return self.mpconv[1](ops.maxpool2x2_nhwc(x))
        return self.mpconv(x)


class _Up(nn.Module):
    """bilinear x2 (align_corners=True) -> zero-pad to the skip's size -> cat([skip, up]) -> double conv (tmp_utils.py:59-88)."""

    def __init__(self, in_ch, out_ch, dropout=0.0):
        super().__init__()
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = _DoubleConv(in_ch, out_ch, dropout)

    def forward(self, x1, x2):
        if self.conv._tc():   # NHWC fp16 in and out; cat([skip, up]) becomes the conv kernel's two K segments
            return self.conv((x2, ops.upsample2x_pad_nhwc(x1, x2.size(1), x2.size(2))))   # upsample + pad in one kernel
        x1 = self.up(x1)
        dy, dx = x2.size(2) - x1.size(2), x2.size(3) - x1.size(3)
        x1 = F.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
        return self.conv(torch.cat([x2, x1], dim=1))


class _OutConv(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv = nn.Conv2d(in_ch, out_ch, 1)

    def forward(self, x):
        return self.conv(x)


class UNet(nn.Module):
    """unet.py:247-290: 64-128-256-512-512 encoder, bilinear decoder, 1x1 output conv."""

    def __init__(self, config: UNetConfig):
        super().__init__()
        self.config = config
        d = config.dropout
        self.inc = _InConv(config.in_channels, 64)
        self.down1 = _Down(64, 128)
        self.down2 = _Down(128, 256)
        self.down3 = _Down(256, 512, d)
        self.down4 = _Down(512, 512, d)
        self.up1 = _Up(1024, 256, d)
        self.up2 = _Up(512, 128, d)
        self.up3 = _Up(256, 64)
        self.up4 = _Up(128, 64)
        self.outc = _OutConv(64, config.out_channels)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("generative_audio_b200.inpainting.UNet: CUDA tensors only (no CPU fallback)")
        return self._forward(x)

    def _forward(self, x):
        tc = self.inc.conv._tc()   # train mode always takes the fp32 NCHW autograd path
        if tc:
            x = ops.nchw_to_nhwc_f16(x, 64)
        x1 = self.inc(x)
        x2 = self.down1(x1)
        x3 = self.down2(x2)
        x4 = self.down3(x3)
        x5 = self.down4(x4)
        x = self.up1(x5, x4)
        x = self.up2(x, x3)
        x = self.up3(x, x2)
        x = self.up4(x, x1)
        if tc:
            return ops.conv1x1_out(x, self.outc.conv.weight, self.outc.conv.bias)
        return self.outc(x.float().contiguous())


class RestorationWrapper(nn.Module):
    """unet.py:293-313: keep the known bins, take the network's output inside the gap."""

    def __init__(self, base_net: UNet):
        super().__init__()
        self.net = base_net

    def forward(self, x_in: torch.Tensor, mask: torch.Tensor):
        return ops.mask_blend(x_in, self.net(x_in), mask)


def gram_schmidt_to_spec_mag(x: torch.Tensor) -> torch.Tensor:
    """inpainting/nppc/pc_wrapper.py:43-59: real MGS over [B, n_dirs, F*T]; un-normalised directions are returned.
    Differentiable like the reference's when x carries a graph (inpainting_training.GramSchmidtRealFn, n_dirs <= 6)."""
    if torch.is_grad_enabled() and x.requires_grad:
        from .inpainting_training import GramSchmidtRealFn
        return GramSchmidtRealFn.apply(x)
    return ops.gram_schmidt_real(x)


class AudioInpaintingPCWrapperConfig(pydantic.BaseModel):
    model_configuration: UNetConfig
    n_dirs: int


class AudioInpaintingPCWrapper(nn.Module):
    """pc_wrapper.py:61-88: UNet(2 -> n_dirs) -> * (1 - mask) -> real Gram-Schmidt.  (The reference's two debug
    `.cpu().numpy()` host syncs at :83,86 are not reproduced.)"""

    def __init__(self, pc_wrapper_config: AudioInpaintingPCWrapperConfig):
        super().__init__()
        self.config = pc_wrapper_config
        self.net = UNet(self.config.model_configuration)

    def head(self, mag_spec: torch.Tensor, mask: torch.Tensor):
        y = self.net(mag_spec)
        if self.training and torch.is_grad_enabled() and y.requires_grad:   # train-mode head: keep the graph (masking kernel as an
            # autograd Function); an eval-mode head stays the graph-less inference path whatever the caller's grad mode is
            from .inpainting_training import MaskOutFn
            return MaskOutFn.apply(y, mask)
        return ops.mask_blend(None, y, mask)

    def forward(self, mag_spec: torch.Tensor, mask: torch.Tensor):
        return gram_schmidt_to_spec_mag(self.head(mag_spec, mask))


class NPPCModelConfig(pydantic.BaseModel):
    pretrained_restoration_model_configuration: UNetConfig
    pretrained_restoration_model_path: Optional[str] = None
    audio_pc_wrapper_configuration: AudioInpaintingPCWrapperConfig
    device: Literal["cpu", "cuda"] = "cuda"


class NPPCModel(nn.Module):
    """nppc_model.py:31-159 (local-checkpoint path; the wandb artifact loader is out of scope)."""

    def __init__(self, config: NPPCModelConfig):
        super().__init__()
        if config.device != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("generative_audio_b200.inpainting.NPPCModel runs on CUDA (sm_100a) only: no CPU fallback")
        self.config = config
        self.device = torch.device("cuda")
        if not config.pretrained_restoration_model_path:
            raise ValueError("pretrained_restoration_model_path must be provided")
        checkpoint = torch.load(Path(config.pretrained_restoration_model_path).absolute(), map_location=self.device)
        base_net = UNet(config.pretrained_restoration_model_configuration)
        base_net.load_state_dict(checkpoint["model_state_dict"])
        base_net.to(self.device)
        self.pretrained_restoration_model = RestorationWrapper(base_net)
        self.pretrained_restoration_model.eval()
        self.pc_wrapper = AudioInpaintingPCWrapper(config.audio_pc_wrapper_configuration)
        self.pc_wrapper.to(self.device)
        self.eval()

    def get_pred_spec_mag_norm(self, masked_spec_mag_log, mask):
        with torch.no_grad():
            return self.pretrained_restoration_model(masked_spec_mag_log, mask)

    # Opt-in, as for the audio model: with `model.differentiable_forward = True` and grad mode on, forward() keeps the autograd
    # graph through the PC head (UNet through torch autograd, masking and real Gram-Schmidt as autograd Functions over the
    # kernels; n_dirs <= 6), so the reference trainer's own `w_mat = self.nppc_model(x, mask)`; loss; `.backward()`
    # (nppc_trainer.py:347-373) works unchanged.  InpaintingNPPCStep.train_step is the faster way (fused Gram-Schmidt + loss).
    differentiable_forward = False

    def forward(self, masked_spec_mag_norm: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """masked_spec_mag_norm, mask [B,1,F,T] -> w_mat [B,n_dirs,F,T] (nppc_model.py:119-145)."""
        pred = self.get_pred_spec_mag_norm(masked_spec_mag_norm, mask)
        x = torch.cat((masked_spec_mag_norm, pred), dim=1)
        if self.differentiable_forward and torch.is_grad_enabled():
            return self.pc_wrapper(x, mask)
        with torch.no_grad():
            return self.pc_wrapper(x, mask)


def preprocess_data(clean_spec: torch.Tensor, masked_spec: torch.Tensor, mask: torch.Tensor):
    """utils.py:294-306: spec [B,2,F,T] (re, im), mask [B,T] -> (clean_log_norm, mask [B,1,F,T], masked_log_norm)."""
    clean, masked, _, _ = ops.logmag_normalize(clean_spec, masked_spec)
    m = mask.unsqueeze(1).unsqueeze(2).expand(-1, 1, clean_spec.shape[2], -1).contiguous()
    return clean, m, masked


def second_moment_lambda(step: int, grace: float, lambda0: float) -> float:
    return max(min(-1.0 + 2.0 * step / grace, 1.0), 1e-6) * lambda0


class InpaintingNPPCStep(CheckpointMixin):
    """The inpainting NPPC trainer's base_step (nppc_trainer.py:338-385) and the body of its train() loop (:146-154): one
    restoration pass is shared by the PC head and the error (the reference runs it twice), Gram-Schmidt and the loss statistics
    are fused.  base_step(batch) is the no-grad statistics path (every op a kernel; the head in whatever mode it is in);
    base_step(batch, requires_grad=True) / train_step build the autograd graph of inpainting_training.py."""

    def __init__(self, nppc_model: NPPCModel, second_moment_loss_lambda: float = 1.0, second_moment_loss_grace: float = 500.0,
                 max_grad_norm: float = 1.0):
        self.nppc_model = nppc_model
        self.lambda0, self.grace = second_moment_loss_lambda, second_moment_loss_grace
        self.max_grad_norm = max_grad_norm      # NPPCAudioInpaintingTrainerConfig.max_grad_norm (nppc_trainer.py:39)
        self.step = 0
        self._reducer = None                    # data parallel: bucketed gradient all-reduce overlapping the backward

    def base_step(self, batch, requires_grad: bool = False):
        """batch = (masked_spec [B,2,F,T], mask [B,T], clean_spec [B,2,F,T]) -> (reconst_err [B], objective [], log dict)."""
        if requires_grad:
            return self._base_step_autograd(batch)
        with torch.no_grad():
            return self._base_step_kernels(batch)

    def _base_step_autograd(self, batch):
        from . import inpainting_training as IT
        masked_spec, mask, clean_spec = batch
        model = self.nppc_model
        with torch.no_grad():
            clean, m, masked = preprocess_data(clean_spec.cuda(), masked_spec.cuda(), mask.cuda())
            pred = model.get_pred_spec_mag_norm(masked, m)
            x = torch.cat((masked, pred), dim=1)
        lam = second_moment_lambda(self.step, self.grace, self.lambda0)
        with torch.enable_grad():
            head = IT.head_forward_train(model.pc_wrapper, x, m)
            objective, w_mat, err_norm, err_proj, w_norms, reconst_err, second_moment = IT.GsLossRealFn.apply(head, clean, pred, lam)
        log = dict(w_mat=w_mat, err_norm=err_norm, err_proj=err_proj, w_norms=w_norms, reconst_err=reconst_err,
                   second_moment_mse=second_moment, objective=objective.detach())
        return reconst_err, objective, log

    def train_step(self, batch, optimizer):
        """One iteration of NPPCAudioInpaintingTrainer.train (nppc_trainer.py:146-154, :184): the PC head in train mode, the
        restoration UNet frozen in eval mode; zero_grad -> backward -> clip_grad_norm_(max_grad_norm) -> optimizer.step.
        Returns (objective, log); log["grad_norm"] is the pre-clip global norm (a device scalar: no host sync in the step).
        Under torch.distributed (one process per GPU) the head's gradients are averaged over the ranks bucket by bucket while
        the backward is still running (training.GradBucketReducer), BatchNorm statistics stay per rank (plain DDP semantics)."""
        import torch.distributed as dist

        from . import inpainting_training as IT
        model = self.nppc_model
        model.pc_wrapper.train()
        model.pretrained_restoration_model.eval()
        if self._reducer is None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from .training import GradBucketReducer     # one process per GPU: mean of the per-rank gradients,
            self._reducer = GradBucketReducer(model.pc_wrapper.parameters())   # each bucket exchanged as soon as it is final
        _, objective, log = self.base_step(batch, requires_grad=True)
        optimizer.zero_grad(set_to_none=True)
        if self._reducer is not None:
            self._reducer.reset()
        objective.backward()
        if self._reducer is not None:
            self._reducer.finish()                       # clip AFTER averaging: the norm of the global-batch gradient, as DDP would
        log["grad_norm"] = IT.clip_grad_norm_(list(model.parameters()), self.max_grad_norm)
        optimizer.step()
        self.step += 1
        return objective.detach(), log

    @torch.no_grad()
    def validate(self, batches):
        """NPPCAudioInpaintingTrainer.validate (nppc_trainer.py:689-706): the whole model in eval mode, mean objective and mean
        reconstruction error over an iterable of batches -> (avg_val_loss, avg_val_reconst_err) as device scalars (one host
        read at the end instead of two per batch); the PC head goes back to train mode afterwards.
        Deviation, on purpose: the reference ends with `self.nppc_model.train()`, which also flips the FROZEN restoration UNet
        into train mode (dropout active, BatchNorm batch statistics) for every later training step; here it stays in eval."""
        model = self.nppc_model
        was_training = model.pc_wrapper.training
        model.eval()
        try:
            losses, errs = [], []
            for batch in batches:
                reconst_err, objective, _ = self._base_step_kernels(batch)
                losses.append(objective)
                errs.append(reconst_err.mean())
            if not losses:
                raise ValueError("validate: no batches")
            return torch.stack(losses).mean(), torch.stack(errs).mean()
        finally:
            model.pc_wrapper.train(was_training)

    def _base_step_kernels(self, batch):
        masked_spec, mask, clean_spec = batch
        clean, m, masked = preprocess_data(clean_spec.cuda(), masked_spec.cuda(), mask.cuda())
        pred = self.nppc_model.get_pred_spec_mag_norm(masked, m)
        head = self.nppc_model.pc_wrapper.head(torch.cat((masked, pred), dim=1), m)
        w_mat, st = ops.gs_loss_fused_real(head, clean, pred)
        lam = second_moment_lambda(self.step, self.grace, self.lambda0)
        objective = st["reconst_err"].mean() + lam * st["second_moment_mse"].mean()
        log = dict(w_mat=w_mat, err_norm=st["err_norm"], err_proj=st["err_proj"], w_norms=st["w_norms"],
                   reconst_err=st["reconst_err"], second_moment_mse=st["second_moment_mse"], objective=objective)
        return st["reconst_err"], objective, log


# ---- row N4 (second half): the MC-dropout + PCA baseline of the inpainting evaluation, batched on the GPU ----------------
def set_compute_dtype(model: nn.Module, dtype=None):
    """Throughput option: run the UNet's 3x3 convolutions in fp16 / bf16 (channels-last, tensor cores, fp32 accumulate inside
    the library kernels); None restores the fp32 parity path.  The 1x1 output convolution and everything outside the UNet
    stay fp32."""
    for m in model.modules():
        if isinstance(m, _DoubleConv):
            m.compute_dtype = dtype


def enable_dropout(model: nn.Module, on: bool = True):
    """utils.enable_dropout (utils.py:334-338): keep the Dropout layers of the UNet active at inference time."""
    for m in model.modules():
        if isinstance(m, _DoubleConv):
            m.mc_dropout = on


def pca_batch(outputs: torch.Tensor, n_components: int = 5):
    """utils.compute_pca_sklearn_batch (utils.py:392-470) without leaving the GPU: outputs [K,B,D] -> unit principal
    components [B,n,D], PCs scaled by their singular values [B,n,D], importance weights S / sum(S) [B,n], mean [B,D],
    singular values [B,n].  One batched thin SVD of the centred [B,K,D] stack instead of a per-item sklearn PCA on the CPU
    (exact SVD; sklearn's auto solver is the randomised one for these shapes), with sklearn's sign convention
    (svd_flip, v-based: the largest-magnitude entry of every component is positive)."""
    K, B, D = outputs.shape
    n = min(n_components, K)
    X = outputs.permute(1, 0, 2).float()
    mean = X.mean(dim=1)
    U, S, Vh = torch.linalg.svd(X - mean[:, None, :], full_matrices=False)
    Vh, S = Vh[:, :n], S[:, :n]
    idx = Vh.abs().argmax(dim=2, keepdim=True)
    Vh = Vh * torch.sign(torch.gather(Vh, 2, idx))
    return Vh, Vh * S[:, :, None], S / S.sum(dim=1, keepdim=True), mean, S


@torch.no_grad()
def calculate_unet_baseline(model: RestorationWrapper, masked_spec: torch.Tensor, mask: torch.Tensor, n_mc_samples: int = 50,
                            n_components: int = 5, chunk: int = 64):
    """utils.calculate_unet_baseline (utils.py:545-648): n_mc_samples stochastic passes of the restoration UNet with dropout
    active, PCA of the predictions inside the gap, results scattered back to full [B,*,F,T] spectrograms (zero outside the
    gap).  The K passes run as batches of `chunk` samples through one UNet call each; PCA is one batched SVD on the GPU.
    Every item must mask the same number of bins (as in the reference)."""
    B, _, Fq, T = masked_spec.shape
    enable_dropout(model, True)
    try:
        reps = masked_spec.repeat(n_mc_samples, 1, 1, 1)   # sample-major: [K*B,1,F,T]
        mreps = mask.repeat(n_mc_samples, 1, 1, 1)
        preds = torch.cat([model(reps[i:i + chunk], mreps[i:i + chunk]) for i in range(0, reps.shape[0], chunk)])
    finally:
        enable_dropout(model, False)
    gap = mask.reshape(B, -1) == 0
    n_gap = int(gap[0].sum())
    if not bool((gap.sum(dim=1) == n_gap).all()):
        raise ValueError("every batch item must mask the same number of bins")
    flat = preds.reshape(n_mc_samples, B, Fq * T)
    vals = flat[:, gap].reshape(n_mc_samples, B, n_gap)
    pcs, scaled, weights, mean, svals = pca_batch(vals, n_components)

    def scatter(v):   # [B,n,n_gap] or [B,n_gap] -> full spectrograms, zero in the known region
        lead = v.shape[1:-1]
        full = torch.zeros(B, *lead, Fq * T, device=v.device, dtype=v.dtype)
        full[gap[:, None, :].expand(B, *lead, Fq * T) if lead else gap] = v.reshape(-1)
        return full.reshape(B, *lead, Fq, T)

    return {"mean_prediction": scatter(mean).unsqueeze(1), "principal_components": scatter(pcs),
            "scaled_principal_components": scatter(scaled), "importance_weights": weights, "singular_vals": svals}
