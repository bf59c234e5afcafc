"""PC head (a11): MultiDirectionFullSubNet_Plus (nppc_audio/networks.py:17-163) — FullSubNet+ with doubled
full-band input (noisy ++ enhanced through shared attention weights), RAW padded noisy magnitude as the
sub-band neighbour source (networks.py:133) and 2*n_directions outputs."""
from typing import Optional

import torch

from . import ops
from .config import MultiDirectionConfig
from .fullsubnet_plus import FullSubNet_Plus


class MultiDirectionFullSubNet_Plus(FullSubNet_Plus):
    def __init__(self, config: Optional[MultiDirectionConfig] = None, lstm_impl: str = "tc"):
        if config is None:
            config = MultiDirectionConfig()
        config.output_size = 2 * config.n_directions  # same in-place mutation as networks.py:23
        super().__init__(config, lstm_impl=lstm_impl)
        self.n_directions = config.n_directions
        self._build_fullband(self.num_freqs * 2)

    @torch.no_grad()
    def forward(self, noisy_mag, noisy_real, noisy_imag, enhanced_mag=None, enhanced_real=None, enhanced_imag=None):
        """6 x [B,1,F,T] -> [B, 2*n_directions, F', T] (channel = dir*2 + {re,im})."""
        B, Cc, F, T = noisy_mag.shape

        def stream(noisy, enh, att, model):
            a = att(self._pad_norm(noisy))
            b = att(self._pad_norm(enh))
            return model(torch.cat([a, b], dim=1)).contiguous()

        fb_out = stream(noisy_mag, enhanced_mag, self.channel_attention, self.fb_model)
        fbr_out = stream(noisy_real, enhanced_real, self.channel_attention_real, self.fb_model_real)
        fbi_out = stream(noisy_imag, enhanced_imag, self.channel_attention_imag, self.fb_model_imag)
        raw = torch.nn.functional.pad(noisy_mag[:, 0], [0, self.look_ahead]).contiguous()
        y, Fp = self._subband(raw, fb_out, fbr_out, fbi_out)
        # [B*F', 2n, T'] -> [B, 2n, F', T]: same memory layout as reshape/permute/slice of networks.py:156-161
        return ops.assemble_mask(y, B, Fp, self.look_ahead)
