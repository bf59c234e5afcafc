"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference (read-only at /root/reference).

Only `oracle/make_golden.py` (run in the dev container, where /root/reference exists) uses this, to
generate the committed fixtures under tests/golden/.  Nothing in the product package, the `-m gpu`
tests, smoke() or bench.py may import it: /root/reference does not exist on the GPU box.

The reference cannot be imported as shipped: `FullSubNet_plus/speech_enhance/utils/logger.py`
is missing and `omegaconf` / `librosa` are not installed.  We pre-register stub modules in sys.modules.
"""
import sys
import types

REF_ROOT = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_shims(trainer: bool = False):
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if "omegaconf" not in sys.modules:
        _stub("omegaconf", ListConfig=type("ListConfig", (list,), {}), DictConfig=dict, OmegaConf=object)
    if "librosa" not in sys.modules:
        _stub("librosa")
    _stub("FullSubNet_plus.speech_enhance.utils.logger", log=lambda *a, **k: None, init=lambda *a, **k: None)
    if trainer:
        for name in ("pesq", "pystoi", "torchinfo", "plotly", "plotly.graph_objects", "soundfile",
                     "matplotlib", "matplotlib.pyplot", "hydra"):
            if name not in sys.modules:
                try:
                    __import__(name)
                except Exception:
                    _stub(name, pesq=None, stoi=None, summary=None, main=lambda *a, **k: (lambda f: f))
        if "line_profiler" not in sys.modules:
            try:
                import line_profiler  # noqa: F401
            except Exception:
                _stub("line_profiler", LineProfiler=object)


def load_reference(trainer: bool = False):
    """Returns a namespace with the reference classes/functions on the hot path."""
    install_shims(trainer)
    import warnings
    warnings.filterwarnings("ignore")
    ns = types.SimpleNamespace()
    import utils as ref_utils
    from FullSubNet_plus.speech_enhance.fullsubnet_plus.model.fullsubnet_plus import FullSubNet_Plus, FullSubNetPlusConfig
    from FullSubNet_plus.speech_enhance.audio_zen.acoustics.mask import (decompress_cIRM, compress_cIRM,
                                                                         build_complex_ideal_ratio_mask)
    from FullSubNet_plus.speech_enhance.audio_zen.acoustics.feature import drop_band
    from FullSubNet_plus.speech_enhance.audio_zen.model.base_model import BaseModel
    from nppc_audio.networks import MultiDirectionConfig, MultiDirectionFullSubNet_Plus
    from nppc_audio.pc_wrapper import AudioPCWrapper, AudioPCWrapperConfig, gram_schmidt_to_crm
    from nppc_audio.nppc_model import NPPCModel, NPPCModelConfig
    ns.utils = ref_utils
    ns.FullSubNet_Plus, ns.FullSubNetPlusConfig = FullSubNet_Plus, FullSubNetPlusConfig
    ns.decompress_cIRM, ns.compress_cIRM = decompress_cIRM, compress_cIRM
    ns.build_complex_ideal_ratio_mask = build_complex_ideal_ratio_mask
    ns.drop_band, ns.BaseModel = drop_band, BaseModel
    ns.MultiDirectionConfig, ns.MultiDirectionFullSubNet_Plus = MultiDirectionConfig, MultiDirectionFullSubNet_Plus
    ns.AudioPCWrapper, ns.AudioPCWrapperConfig = AudioPCWrapper, AudioPCWrapperConfig
    ns.gram_schmidt_to_crm = gram_schmidt_to_crm
    ns.NPPCModel, ns.NPPCModelConfig = NPPCModel, NPPCModelConfig
    if trainer:
        from nppc_audio.trainer import NPPCAudioTrainer
        ns.NPPCAudioTrainer = NPPCAudioTrainer
    return ns
