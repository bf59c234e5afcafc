"""Pydantic configs with the reference's field names and defaults, so reference yaml/dicts construct them
unchanged: FullSubNetPlusConfig (fullsubnet_plus.py:18-42), MultiDirectionConfig (networks.py:9-14),
AudioPCWrapperConfig (pc_wrapper.py:46-51), StftConfig (utils.py:14-17), NPPCModelConfig (nppc_model.py:13-22).
Extra knob (not in the reference): `lstm_impl` selects the LSTM kernel ("tc" = fp16-operand / fp32-accumulate tcgen05 kernels for the LSTM and the TCN 1x1 convolutions, "f32" = fp32 SIMT LSTM + fp32 TCN, "tcp" = split-precision tensor-core path (hi + lo fp16 operands: fp32-class accuracy on tcgen05), "auto" = "tc" with the
utterances whose normaliser means cancel re-run through "tcp" (NPPCModel.forward, DESIGN.md "Conditioning"))."""
from typing import List, Literal

import pydantic


class StftConfig(pydantic.BaseModel):
    nfft: int = 512
    hop_length: int = 256
    win_length: int = 512


class FullSubNetPlusConfig(pydantic.BaseModel):
    num_freqs: int = 257
    look_ahead: int = 2
    sequence_model: str = "LSTM"
    sb_num_neighbors: int = 15
    fb_num_neighbors: int = 0
    fb_output_activate_function: str = "ReLU"
    sb_output_activate_function: bool = False
    fb_model_hidden_size: int = 512
    sb_model_hidden_size: int = 384
    channel_attention_model: str = "TSSE"
    norm_type: str = "offline_laplace_norm"
    num_groups_in_drop_band: int = 1
    output_size: int = 2
    subband_num: int = 1
    kersize: List[int] = pydantic.Field(default_factory=lambda: [3, 5, 10])
    weight_init: bool = False

    @pydantic.field_validator("kersize", mode="before")
    @classmethod
    def _kersize_to_list(cls, v):
        # the reference converts omegaconf ListConfig -> list here (fullsubnet_plus.py:36-42)
        if not isinstance(v, list):
            try:
                v = list(v)
            except TypeError:
                raise ValueError(f"kersize must be a list of integers, got {type(v).__name__}")
        return v


class MultiDirectionConfig(FullSubNetPlusConfig):
    n_directions: int = 4


class AudioPCWrapperConfig(pydantic.BaseModel):
    multi_direction_configuration: MultiDirectionConfig


class NPPCModelConfig(pydantic.BaseModel):
    pretrained_restoration_model_configuration: FullSubNetPlusConfig
    pretrained_restoration_model_path: str
    audio_pc_wrapper_configuration: AudioPCWrapperConfig
    stft_configuration: StftConfig
    device: Literal["cpu", "cuda"] = "cuda"
    lstm_impl: Literal["tc", "f32", "tcp", "auto"] = "tc"
