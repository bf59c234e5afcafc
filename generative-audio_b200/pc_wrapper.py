"""PC wrapper (a12): gram_schmidt_to_crm + AudioPCWrapper (nppc_audio/pc_wrapper.py:8-106), and the real-valued
variant of the inpainting wrapper (nppc_audio/inpainting/nppc/pc_wrapper.py:43-59)."""
import torch
import torch.nn as nn

from . import ops
from .config import AudioPCWrapperConfig
from .networks import MultiDirectionFullSubNet_Plus


def gram_schmidt_to_crm(x: torch.Tensor) -> torch.Tensor:
    """x [B, n_dirs, 2, F, T] -> orthogonalised (un-normalised) directions, same shape.  Differentiable like the reference's
    (pc_wrapper.py:8-44, normalisers detached at :37) when x carries a graph: the kernels then run under
    training.GramSchmidtFn (n_dirs <= 6) instead of silently dropping the gradient."""
    if torch.is_grad_enabled() and x.requires_grad:
        from .training import GramSchmidtFn
        return GramSchmidtFn.apply(x)
    return ops.gram_schmidt_complex(x)


def gram_schmidt_to_spec_mag(x: torch.Tensor) -> torch.Tensor:
    """x [B, n_dirs, F, T] real -> same shape (graph-preserving: see inpainting.gram_schmidt_to_spec_mag)."""
    from .inpainting import gram_schmidt_to_spec_mag as impl
    return impl(x)


class AudioPCWrapper(nn.Module):
    def __init__(self, audio_pc_wrapper_config: AudioPCWrapperConfig, lstm_impl: str = "tc"):
        super().__init__()
        self.net = MultiDirectionFullSubNet_Plus(audio_pc_wrapper_config.multi_direction_configuration, lstm_impl=lstm_impl)
        self.n_dirs = self.net.n_directions

    def head(self, noisy_mag, noisy_real, noisy_imag, enhanced_mag, enhanced_real, enhanced_imag):
        crm = self.net(noisy_mag, noisy_real, noisy_imag, enhanced_mag, enhanced_real, enhanced_imag)
        B, _, Fp, T = crm.shape
        return crm.reshape(B, self.n_dirs, 2, Fp, T)

    def forward(self, noisy_mag, noisy_real, noisy_imag, enhanced_mag=None, enhanced_real=None, enhanced_imag=None):
        """6 x [B,1,F,T] -> w_mat [B, n_dirs, 2, F', T]."""
        return gram_schmidt_to_crm(self.head(noisy_mag, noisy_real, noisy_imag, enhanced_mag, enhanced_real, enhanced_imag))
