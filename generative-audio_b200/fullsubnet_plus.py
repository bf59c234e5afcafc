"""FullSubNet+ backbone (a8) on B200: mirrors FullSubNet_Plus (fullsubnet_plus.py:45-230).

Data flow per forward (all device-resident, fp32 unless noted):
  pad+offline-norm (1 fused kernel per plane) -> TSSE (3 fused kernels) -> TCN ("tc": channel-last, every 1x1 convolution on
  the in-house tcgen05 GEMM with split-precision operands; "f32" / "tcp": fp32 SGEMM + 3 fused kernels per block)
  -> fused sub-band pack (unfold ++ cat ++ offline | cumulative norm ++ drop_band, written time-major [T',R,64]; fp16 for the
     tensor-core LSTM) -> 2-layer LSTM + fc kernels -> mask assembly kernel.
The [B,F,34,T'] sub-band tensor of the reference is never materialised."""
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .config import FullSubNetPlusConfig
from .modules import ChannelTimeSenseSELayer, SequenceModel

KP = 64  # padded feature width of the packed LSTM input (34 real features)
_IMPL = {"f32": 0, "tc": 1, "tcp": 2}   # tcp: split-precision tensor-core path (fp32-class accuracy)


class FullSubNet_Plus(nn.Module):
    def __init__(self, config: Optional[FullSubNetPlusConfig] = None, lstm_impl: str = "tc"):
        super().__init__()
        if config is None:
            config = FullSubNetPlusConfig()
        self.num_freqs = config.num_freqs
        self.look_ahead = config.look_ahead
        self.sequence_model = config.sequence_model
        self.fb_num_neighbors = config.fb_num_neighbors
        self.sb_num_neighbors = config.sb_num_neighbors
        self.norm_type = config.norm_type
        self.num_groups_in_drop_band = config.num_groups_in_drop_band
        self.output_size = config.output_size
        self.kersize = config.kersize
        self.lstm_impl = lstm_impl
        assert self.sequence_model in ("GRU", "LSTM", "TCN"), f"{self.__class__.__name__} only support GRU, LSTM and TCN."
        if self.sequence_model != "LSTM":
            raise NotImplementedError("only the LSTM sub-band model (every shipped config) is built")
        if config.subband_num != 1 or config.channel_attention_model != "TSSE" or config.fb_num_neighbors != 0:
            raise NotImplementedError("only subband_num=1, TSSE attention, fb_num_neighbors=0 (every shipped config) are built")
        if self.norm_type not in ("offline_laplace_norm", "cumulative_laplace_norm"):
            raise NotImplementedError(f"norm_type {self.norm_type}")
        # configuration fields the reference honours and this build does not: refuse them loudly instead of diverging silently
        if config.fb_output_activate_function != "ReLU":
            raise NotImplementedError("only fb_output_activate_function='ReLU' (every shipped config) is built, got "
                                      f"{config.fb_output_activate_function!r}")
        if config.weight_init:
            raise NotImplementedError("weight_init=True (fullsubnet_plus.py:140-141 re-initialises the weights) is not built: "
                                      "load a state_dict instead")
        if 2 * self.sb_num_neighbors + 4 > KP:
            raise NotImplementedError("sb_num_neighbors too large for the packed LSTM input")
        C = self.num_freqs
        self.channel_attention = ChannelTimeSenseSELayer(C, kersize=self.kersize)
        self.channel_attention_real = ChannelTimeSenseSELayer(C, kersize=self.kersize)
        self.channel_attention_imag = ChannelTimeSenseSELayer(C, kersize=self.kersize)
        self.fb_input_size = C
        self._build_fullband(C)
        self.sb_model = SequenceModel(input_size=(self.sb_num_neighbors * 2 + 1) + 3 * (self.fb_num_neighbors * 2 + 1),
                                      output_size=self.output_size, hidden_size=config.sb_model_hidden_size, num_layers=2,
                                      bidirectional=False, sequence_model="LSTM",
                                      output_activate_function=config.sb_output_activate_function)

    def _build_fullband(self, input_size):
        kw = dict(input_size=input_size, output_size=self.num_freqs, hidden_size=512, num_layers=2, bidirectional=False,
                  sequence_model="TCN", output_activate_function="ReLU")
        self.fb_model = SequenceModel(**kw)
        self.fb_model_real = SequenceModel(**kw)
        self.fb_model_imag = SequenceModel(**kw)
        import os
        # 1x1 convs on the tcgen05 GEMM (row N2) for "tc"; the accuracy-first "tcp" path keeps the fp32 TCN (fp32 SGEMM + the
        # fused kernels): on a digitally silent utterance the residual stream IS the sum of the block outputs, so the fp16
        # y1 / z / o intermediates of the channel-last path show up at 1.6e-3 in pred_crm (tests/golden/speech12.npz, #3),
        # and next to the split-precision LSTM the fp32 TCN costs < 10 % of the step
        tc = self.lstm_impl == "tc" and os.environ.get("NPPC_TCN_TC", "1") != "0"
        for m in (self.fb_model, self.fb_model_real, self.fb_model_imag):
            m.use_tc_convs = tc
            m.tc_split = os.environ.get("NPPC_TCN_SPLIT", "1") != "0"

    # -- helpers --------------------------------------------------------------------------------------
    def _pad_norm(self, x):
        """F.pad(x,[0,look_ahead]) + self.norm(x), x [B,1,F,T] -> [B,F,T']."""
        if self.norm_type == "offline_laplace_norm":
            return ops.pad_offline_laplace_norm(x, self.look_ahead)
        xp = torch.nn.functional.pad(x, [0, self.look_ahead]).contiguous()
        return ops.cumulative_laplace_norm(xp)[:, 0]

    def _subband(self, nbr_src, fb, fbr, fbi):
        """-> (y [R,O,T'], F') from the four [B,F,T'] planes."""
        B, F, Tp = fb.shape
        impl = _IMPL[self.lstm_impl]
        dt = torch.float16 if impl == 1 else torch.float32
        G = self.num_groups_in_drop_band
        # one fused kernel for both norm types: unfold ++ cat ++ (offline | cumulative) norm ++ drop_band -> time-major input
        xs, R = ops.subband_pack(nbr_src, fb, fbr, fbi, self.sb_num_neighbors, G, KP, dt, pad_rows=(impl == 2),
                                 cumulative=(self.norm_type == "cumulative_laplace_norm"))
        Fp = R // B
        return self.sb_model.lstm_forward(xs, impl, R), Fp

    @torch.no_grad()
    def forward(self, noisy_mag, noisy_real, noisy_imag):
        """[B,1,F,T] x3 -> compressed cRM [B,2,F',T] (fullsubnet_plus.py:143-230)."""
        assert noisy_mag.dim() == 4
        B, Cc, F, T = noisy_mag.shape
        assert Cc == 1, f"{self.__class__.__name__} takes the mag feature as inputs."
        fb_in = self.channel_attention(self._pad_norm(noisy_mag))
        fb_out = self.fb_model(fb_in)
        fbr_out = self.fb_model_real(self.channel_attention_real(self._pad_norm(noisy_real)))
        fbi_out = self.fb_model_imag(self.channel_attention_imag(self._pad_norm(noisy_imag)))
        y, Fp = self._subband(fb_in.contiguous(), fb_out.contiguous(), fbr_out.contiguous(), fbi_out.contiguous())
        return ops.assemble_mask(y, B, Fp, self.look_ahead)
