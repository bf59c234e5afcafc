"""NPPCModel (a13): drop-in for nppc_audio/nppc_model.py:25-132 — same constructor config, forward /
get_pred_crm signatures, sub-module attribute names and state_dict keys; CUDA only (no CPU fallback)."""
import os
from pathlib import Path

import torch
import torch.nn as nn

from . import ops
from .config import NPPCModelConfig
from .fullsubnet_plus import FullSubNet_Plus
from .pc_wrapper import AudioPCWrapper


def load_pretrained_model(model_path, model_config, lstm_impl="tc") -> FullSubNet_Plus:
    """utils.load_pretrained_model / preload_model (utils.py:82-104): `{"model": state_dict}` .tar, strict=False."""
    model = FullSubNet_Plus(model_config, lstm_impl=lstm_impl)
    p = Path(model_path).expanduser().absolute()
    assert p.exists(), f"The file {p.as_posix()} is not exist. please check path."
    ck = torch.load(p.as_posix(), map_location="cpu")
    model.load_state_dict(ck["model"], strict=False)
    return model


class NPPCModel(nn.Module):
    def __init__(self, config: NPPCModelConfig):
        super().__init__()
        self.config = config
        if config.device != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("generative-audio_b200.NPPCModel runs on CUDA (sm_100a) only: there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device())
        impl0 = "tc" if config.lstm_impl == "auto" else config.lstm_impl
        self.pretrained_restoration_model = load_pretrained_model(
            config.pretrained_restoration_model_path, config.pretrained_restoration_model_configuration, impl0)
        self.pretrained_restoration_model.to(self.device).eval()
        self.audio_pc_wrapper = AudioPCWrapper(config.audio_pc_wrapper_configuration, lstm_impl=impl0)
        self.audio_pc_wrapper.to(self.device)
        # lstm_impl="auto": utterances whose enhanced real / imag normaliser means cancel deeper than this (x the depth of a
        # random-sign sum) are re-run through the split-precision path; expected fp16-path error ~ 7e-4 * depth * O(1..6)
        self.auto_depth_threshold = float(os.environ.get("NPPC_AUTO_DEPTH", "2.5"))
        self.last_auto = None

    def _set_impl(self, impl: str):
        """Switch both networks between the fp16 tensor-core path ("tc"), the split-precision one ("tcp") and fp32 SIMT."""
        tc_convs = impl == "tc" and os.environ.get("NPPC_TCN_TC", "1") != "0"
        for net in (self.pretrained_restoration_model, self.audio_pc_wrapper.net):
            net.lstm_impl = impl
            for m in (net.fb_model, net.fb_model_real, net.fb_model_imag):
                m.use_tc_convs = tc_convs

    def _stft(self, wave):
        c = self.config.stft_configuration
        return ops.stft_mri(wave.to(self.device, non_blocking=True), c.nfft, c.hop_length, c.win_length)

    @torch.no_grad()
    def forward_stages(self, noisy_waveform: torch.Tensor, taps: dict = None):
        """STFT -> frozen backbone -> decompress + conj(M)*N (utils.py:241-249 quirk) -> PC head (pre-Gram-Schmidt).
        Returns (head [B,n,2,F',T], pred_crm [B,2,F,T]); `taps` (optional dict) receives the noisy / enhanced spectra."""
        mag, real, imag = self._stft(noisy_waveform)
        pred_crm = self.pretrained_restoration_model(mag, real, imag)
        emag, ereal, eimag = ops.crm_decompress_apply(pred_crm, real, imag, conj=True)
        head = self.audio_pc_wrapper.head(mag, real, imag, emag[:, None], ereal[:, None], eimag[:, None])
        if taps is not None:
            taps.update(mag=mag, real=real, imag=imag, emag=emag, ereal=ereal, eimag=eimag)
        return head, pred_crm

    @torch.no_grad()
    def conditioning(self, noisy_waveform: torch.Tensor) -> torch.Tensor:
        """Per-utterance cancellation depth [B] of the PC head's enhanced real / imag normalisers (the amplifier of the fp16
        path's backbone error, DESIGN.md "Conditioning"), from one pass of the current path."""
        taps = {}
        self.forward_stages(noisy_waveform, taps)
        return self._depth(taps)

    def _depth(self, taps):
        la = self.audio_pc_wrapper.net.look_ahead
        Fq, T = taps["ereal"].shape[-2:]
        count = Fq * (T + la)     # the reference's mean runs over the zero-padded look-ahead frames too
        d = ops.cancel_depth(taps["ereal"], count)
        return ops.cancel_depth(taps["eimag"], count, d)

    # Opt-in: `model.differentiable_forward = True` makes forward() build an autograd graph through the PC head whenever grad
    # mode is on — the reference trainer's own pattern (`w_mat = self.nppc_model(noisy)`, a loss in torch, `.backward()`,
    # nppc_audio/trainer.py:255-298,100-106) then works unchanged.  Off by default: inference callers rarely wrap their calls
    # in no_grad, and the training forward keeps every LSTM gate for BPTT (18 GB at B = 32).  NPPCAudioStep (trainer.py)
    # is the faster way to train: it fuses Gram-Schmidt with the objective and shares the backbone pass with get_pred_crm.
    differentiable_forward = False

    def forward_train(self, noisy_waveform: torch.Tensor) -> torch.Tensor:
        """forward() with gradients w.r.t. the PC head's parameters: frozen half on the inference kernels (no_grad, as
        nppc_model.py:94), head + Gram-Schmidt through the hand-written autograd Functions of training.py."""
        from . import training
        with torch.no_grad():
            mag, real, imag = self._stft(noisy_waveform)
            pred_crm = self.pretrained_restoration_model(mag, real, imag)
            emag, ereal, eimag = ops.crm_decompress_apply(pred_crm, real, imag, conj=True)
        with torch.enable_grad():
            head = training.head_forward_train(self.audio_pc_wrapper.net, mag, real, imag, emag[:, None], ereal[:, None], eimag[:, None])
            return training.GramSchmidtFn.apply(head)

    def forward(self, noisy_waveform: torch.Tensor) -> torch.Tensor:
        """noisy_waveform [B, L] -> w_mat [B, n_dirs, 2, F', T]."""
        if self.differentiable_forward and torch.is_grad_enabled():
            return self.forward_train(noisy_waveform)
        with torch.no_grad():
            return self._forward_inference(noisy_waveform)

    def _forward_inference(self, noisy_waveform: torch.Tensor) -> torch.Tensor:
        if self.config.lstm_impl != "auto":
            head, _ = self.forward_stages(noisy_waveform)
            return ops.gram_schmidt_complex(head)
        # auto: fp16 tensor-core pass for everybody, split-precision re-run for the utterances whose normalisers cancel
        if self.audio_pc_wrapper.net.num_groups_in_drop_band > 1 and noisy_waveform.shape[0] > 1:
            raise NotImplementedError('lstm_impl="auto" needs num_groups_in_drop_band == 1 (drop_band interleaves the batch)')
        taps = {}
        self._set_impl("tc")
        head, _ = self.forward_stages(noisy_waveform, taps)
        w = ops.gram_schmidt_complex(head)
        depth = self._depth(taps)
        idx = torch.nonzero(depth > self.auto_depth_threshold).flatten()      # host sync: the price of the routing decision
        if idx.numel():
            self._set_impl("tcp")
            try:
                hp, _ = self.forward_stages(noisy_waveform.to(self.device)[idx].contiguous())
                w[idx] = ops.gram_schmidt_complex(hp)
            finally:
                self._set_impl("tc")
        self.last_auto = dict(depth=depth, rerouted=idx)
        return w

    @torch.no_grad()
    def forward_host(self, noisy_host: torch.Tensor, out_host: torch.Tensor = None) -> torch.Tensor:
        """Serving loop entry: PINNED host waveform [B, L] -> PINNED host w_mat [B, n_dirs, 2, F', T].
        The upload and the kernels run on the current stream; the download of w_mat (165 MB at B = 64) is queued on a side
        stream behind them, so it overlaps the NEXT call's compute instead of stalling it.  Returns out_host immediately;
        its contents are valid after `self.host_copy_done()` (or any device synchronise)."""
        dev = next(self.parameters()).device
        if not (noisy_host.is_pinned() and (out_host is None or out_host.is_pinned())):
            raise ValueError("forward_host needs pinned host tensors (torch.Tensor.pin_memory())")
        w = self.forward(noisy_host.to(dev, non_blocking=True))
        if out_host is None:
            out_host = torch.empty(w.shape, dtype=w.dtype).pin_memory()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        self._copy_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._copy_stream):
            out_host.copy_(w, non_blocking=True)
        w.record_stream(self._copy_stream)
        return out_host

    def host_copy_done(self):
        """Block the host until every download queued by forward_host has landed."""
        if getattr(self, "_copy_stream", None) is not None:
            self._copy_stream.synchronize()

    @torch.no_grad()
    def get_pred_crm(self, noisy_waveform: torch.Tensor) -> torch.Tensor:
        """compressed cRM of the frozen backbone, [B,2,F,T] (nppc_model.py:117-132)."""
        mag, real, imag = self._stft(noisy_waveform)
        return self.pretrained_restoration_model(mag, real, imag)

    @torch.no_grad()
    def enhance(self, noisy_waveform: torch.Tensor) -> torch.Tensor:
        """Enhance-only pipeline (BASELINE config #2; model_validator.py:84-133 == utils.py:37-72):
        STFT -> backbone -> decompress -> M*N -> iSTFT(length=L)."""
        c = self.config.stft_configuration
        mag, real, imag = self._stft(noisy_waveform)
        crm = self.pretrained_restoration_model(mag, real, imag)
        _, er, ei = ops.crm_decompress_apply(crm, real, imag, conj=False, want_mag=False)
        return ops.istft(er, ei, noisy_waveform.shape[-1], c.nfft, c.hop_length)

    @torch.no_grad()
    def pc_variations(self, noisy_waveform: torch.Tensor, alphas: torch.Tensor = None, normalize: bool = True):
        """The validator's consumer of the PCs (NPPCAudioValidator._crm_directions_to_spectograms +
        the alpha sweep and save_audio_files normalisation of visualize_pc_spectrograms, validator.py:55-143,246-290):
        every direction's decompressed cRM applied to the noisy STFT (M*N), `enhanced + alpha * pc` for every alpha, ONE
        batched iSTFT over all B*n*A variations and the per-waveform peak normalisation — one shared backbone pass and
        4 kernel launches instead of the reference's second backbone pass and 1 + n*A serial torch.istft calls.
        Returns dict(w_mat, pc_real, pc_imag [B,n,F,T], enhanced_real, enhanced_imag [B,F,T], enhanced [B,L],
        variations [B,n,A,L], alphas [A])."""
        c = self.config.stft_configuration
        L = noisy_waveform.shape[-1]
        if alphas is None:
            alphas = torch.linspace(-3, 3, 6)   # validator.py:246
        mag, real, imag = self._stft(noisy_waveform)
        pred_crm = self.pretrained_restoration_model(mag, real, imag)
        emag, ereal, eimag = ops.crm_decompress_apply(pred_crm, real, imag, conj=True)
        head = self.audio_pc_wrapper.head(mag, real, imag, emag[:, None], ereal[:, None], eimag[:, None])
        w_mat = ops.gram_schmidt_complex(head)
        if w_mat.shape[3] != real.shape[2]:
            raise RuntimeError("pc_variations needs num_groups_in_drop_band == 1 (full-band directions), as the validator does")
        _, er, ei = ops.crm_decompress_apply(pred_crm, real, imag, conj=False, want_mag=False)   # M*N: the audible estimate
        pc_re, pc_im, var_re, var_im = ops.pc_variations(w_mat, real, imag, er, ei, alphas)
        B, n, A = var_re.shape[:3]
        Fq, T = var_re.shape[3:]
        waves = ops.istft(var_re.reshape(B * n * A, Fq, T), var_im.reshape(B * n * A, Fq, T), L, c.nfft, c.hop_length)
        enhanced = ops.istft(er, ei, L, c.nfft, c.hop_length)
        if normalize:
            ops.peak_normalize_(waves)
            ops.peak_normalize_(enhanced)
        return dict(w_mat=w_mat, pc_real=pc_re, pc_imag=pc_im, enhanced_real=er, enhanced_imag=ei, enhanced=enhanced,
                    variations=waves.reshape(B, n, A, L), alphas=alphas)
