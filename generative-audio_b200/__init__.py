"""generative-audio_b200 — B200-native (sm_100a) NPPC-over-FullSubNet+ speech-restoration hot path.

Drop-in surface of the reference's `nppc_audio` package for this path: NPPCModel, AudioPCWrapper,
MultiDirectionFullSubNet_Plus, FullSubNet_Plus, gram_schmidt_to_crm and their pydantic configs.
Import as `generative_audio_b200` (alias package at the repo root; a hyphen cannot be imported directly)."""
from . import _lib, inpainting, modules, ops, sharding, training  # noqa: F401
from .config import (AudioPCWrapperConfig, FullSubNetPlusConfig, MultiDirectionConfig, NPPCModelConfig,  # noqa: F401
                     StftConfig)
from .fullsubnet_plus import FullSubNet_Plus  # noqa: F401
from .networks import MultiDirectionFullSubNet_Plus  # noqa: F401
from .nppc_model import NPPCModel, load_pretrained_model  # noqa: F401
from .pc_wrapper import AudioPCWrapper, gram_schmidt_to_crm, gram_schmidt_to_spec_mag  # noqa: F401
from .trainer import NPPCAudioStep, second_moment_lambda  # noqa: F401
