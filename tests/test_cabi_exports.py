"""CPU checks of the drop-in boundary: libnppc_b200.so builds/loads without a GPU, exports every symbol that
include/nppc_b200.h declares (and the ctypes table lists each of them), and the host-side mirror keeps the reference's
state_dict contract (680 tensors, same keys and shapes as nppc_audio.NPPCModel)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "nppc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nppc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    import generative_audio_b200 as g
    lib = g._lib.load()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nppc_b200.h but not exported"
        assert n in g._lib.PROTOTYPES, f"{n} missing from the ctypes prototype table"
    assert set(g._lib.PROTOTYPES) <= set(names), set(g._lib.PROTOTYPES) - set(names)
    assert b"sm_100a" in lib.nppc_version()


def test_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call: usable on the CPU box."""
    import generative_audio_b200 as g
    lib = g._lib.load()
    assert lib.nppc_stft_mri(None, 1, 4096, 512, 256, None, None, None, None) == -1
    assert b"null pointer" in lib.nppc_last_error()
    assert lib.nppc_drop_band(1, 2, 3, 257, 7, 2, 1, None) == -1   # B must be > groups (feature.py:263)
    assert b"The batch size should larger than the num_groups" in lib.nppc_last_error()
    assert lib.nppc_gs_scratch_bytes(4, 5) > 0
    # round-2 entry points: shapes are validated before any CUDA call
    assert lib.nppc_gemm_f16_atb(1, 1, 100, 128, 64, 1, 1, 1, None) == -1 and b"rows % 64" in lib.nppc_last_error()
    assert lib.nppc_gemm_f16_tn_ex(16, 16, None, 16, 128, 128, 128, 192, 0, None) == -1 and b"KA" in lib.nppc_last_error()   # KA > K
    assert lib.nppc_conv3x3_tc(1, 48, None, 0, 1, 1, 1, 1, 8, 8, 64, 0.2, None) == -1 and b"multiples of 64" in lib.nppc_last_error()
    assert lib.nppc_conv1x1_out(1, 1, 10, 64, 1, 1, 17, 1, None) == -1
    assert lib.nppc_upsample2x_pad_nhwc(1, 1, 8, 8, 64, 15, 16, 1, None) == -1 and b"2h x 2w" in lib.nppc_last_error()
    assert lib.nppc_cancel_depth(None, 1, 10, 10.0, None, None, None) == -1
    assert lib.nppc_lstm_step_forward(None, None, 0, 1, 128, 1, 64, 0, 0, None, 0, None, None) == -1
    assert lib.nppc_lstm_step_workspace_bytes(34, 384, 10, 4096, 253, 64, 1, 0) > lib.nppc_lstm_step_workspace_bytes(34, 384, 10, 4096, 253, 64, 0, 0) > 0


def test_state_dict_contract():
    import json

    import generative_audio_b200 as g
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_manifest.json")))["entries"]
    bb = g.FullSubNet_Plus()
    head = g.MultiDirectionFullSubNet_Plus(g.MultiDirectionConfig(n_directions=5))
    sd = {"pretrained_restoration_model." + k: tuple(v.shape) for k, v in bb.state_dict().items()}
    sd.update({"audio_pc_wrapper.net." + k: tuple(v.shape) for k, v in head.state_dict().items()})
    ref = {k: tuple(10 if s == "2*n_dirs" else s for s in shp) for k, shp in man}
    assert sd == ref and len(sd) == 680


def test_cpu_tensors_and_cpu_device_are_refused():
    import torch

    import generative_audio_b200 as g
    with pytest.raises(RuntimeError):
        g.ops.stft_mri(torch.zeros(1, 4096))
    cfg = g.NPPCModelConfig(pretrained_restoration_model_configuration=g.FullSubNetPlusConfig(),
                            pretrained_restoration_model_path="/nonexistent.tar",
                            audio_pc_wrapper_configuration=g.AudioPCWrapperConfig(
                                multi_direction_configuration=g.MultiDirectionConfig(n_directions=5)),
                            stft_configuration=g.StftConfig(), device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g.NPPCModel(cfg)


def test_loss_lambda_schedule():
    import nppc_oracle as O
    import generative_audio_b200 as g
    for step in (0, 100, 250, 251, 400, 500, 600):
        assert g.second_moment_lambda(step, 500, 1.0) == O.second_moment_lambda(step, 500, 1.0)
    assert g.second_moment_lambda(0, 500, 1.0) == 1e-6 and g.second_moment_lambda(600, 500, 0.5) == 0.5


def test_derived_caches_are_not_copied_or_pickled_and_unsupported_configs_raise():
    """a SequenceModel / TCNBlock / UNet block that has already run holds derived caches (a ctypes LSTM-plan handle,
    packed / folded 16-bit weights); copy.deepcopy / torch.save must drop them, and configuration fields this build does not
    implement must raise instead of silently diverging from the reference."""
    import copy
    import ctypes
    import io

    import torch

    import generative_audio_b200 as g
    m = g.modules.SequenceModel(34, 2, 384, 2, False, "LSTM", False)
    m._plan, m._plan_key = ctypes.c_void_p(1234), ("key",)
    m2 = copy.deepcopy(m)
    assert m2._plan is None and m2._plan_key is None and m._plan is not None
    assert torch.equal(m2.fc_output_layer.weight, m.fc_output_layer.weight) and m2.fc_output_layer.weight is not m.fc_output_layer.weight
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    assert torch.load(buf, weights_only=False)._plan is None
    t = g.modules.SequenceModel(257, 257, 512, 2, False, "TCN", "ReLU")
    t._tplan, t.sequence_model[0]._fold = {"stale": 1}, ("stale",)
    t2 = copy.deepcopy(t)
    assert t2._tplan is None and t2.sequence_model[0]._fold is None
    u = g.inpainting.UNet(g.inpainting.UNetConfig())
    u.inc.conv._tc = ("stale",)
    assert copy.deepcopy(u).inc.conv._tc is None
    for bad in (dict(weight_init=True), dict(fb_output_activate_function="Tanh")):
        with pytest.raises(NotImplementedError):
            g.FullSubNet_Plus(g.FullSubNetPlusConfig(**bad))
