"""ctypes binding of libnppc_b200.so (C ABI declared in include/nppc_b200.h).

The product path has NO fallback: if the shared library is missing it is built with nvcc (build.py); if that
fails, or the symbols are missing, importing raises.  Nothing here touches oracle/.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnppc_b200.so")

_p, _i, _ll, _sz, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_size_t, C.c_float

# name -> (restype, argtypes); must list every symbol of include/nppc_b200.h
PROTOTYPES = {
    "nppc_last_error": (C.c_char_p, []),
    "nppc_version": (C.c_char_p, []),
    "nppc_launch_count": (_ll, []),
    "nppc_reset_launch_count": (None, []),
    "nppc_stft_mri": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "nppc_istft": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "nppc_crm_decompress_apply": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "nppc_decompress_cirm": (_i, [_p, _ll, _p, _p]),
    "nppc_build_cirm": (_i, [_p, _p, _p, _p, _i, _i, _p, _p]),
    "nppc_offline_laplace_norm": (_i, [_p, _i, _ll, _p, _p, _p]),
    "nppc_pad_offline_laplace_norm": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "nppc_cancel_depth": (_i, [_p, _i, _ll, C.c_double, _p, _p, _p]),
    "nppc_cumulative_laplace_norm": (_i, [_p, _i, _i, _i, _p, _p]),
    "nppc_unfold": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "nppc_drop_band": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "nppc_complex_lincomb": (_i, [_p, _p, _p, _i, _i, _ll, _p, _p, _p]),
    "nppc_gs_scratch_bytes": (_sz, [_i, _i]),
    "nppc_gram_schmidt_complex": (_i, [_p, _i, _i, _ll, _p, _p, _p]),
    "nppc_gram_schmidt_real": (_i, [_p, _i, _i, _ll, _p, _p, _p]),
    "nppc_gs_loss_fused": (_i, [_p, _p, _p, _i, _i, _ll, _p, _p, _p, _p, _p, _p, _p, _p]),
    "nppc_projection_loss": (_i, [_p, _p, _p, _i, _i, _ll, _p, _p, _p, _p, _p, _p, _p]),
    "nppc_gs_loss_fused_real": (_i, [_p, _p, _p, _i, _i, _ll, _p, _p, _p, _p, _p, _p, _p, _p]),
    "nppc_projection_loss_real": (_i, [_p, _p, _p, _i, _i, _ll, _p, _p, _p, _p, _p, _p, _p]),
    "nppc_logmag_stats": (_i, [_p, _i, _ll, _p, _p]),
    "nppc_logmag_apply": (_i, [_p, _i, _ll, _p, _ll, _p, _p]),
    "nppc_mask_blend": (_i, [_p, _i, _p, _p, _i, _i, _ll, _p, _p]),
    "nppc_pc_variations": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p]),
    "nppc_peak_normalize": (_i, [_p, _i, _i, _p]),
    "nppc_mix_scratch_bytes": (_sz, [_i]),
    "nppc_mix_with_snr": (_i, [_p, _p, _i, _i, _p, _p, _p, _p, _p, _p]),
    "nppc_time_to_spec_mask": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "nppc_subband_pack": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "nppc_tsse": (_i, [_p, _i, _i, _i, C.POINTER(_i), C.POINTER(_p), C.POINTER(_p)] + [_p] * 6 + [_i, _p, _p, _p]),
    "nppc_prelu_stats": (_i, [_p, _i, _i, _i, _p, _p, _p, _p]),
    "nppc_tcn_mid": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p]),
    "nppc_tcn_out": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "nppc_lstm_plan_create": (_i, [C.POINTER(_p), _i, _i, _i] + [_p] * 10 + [_p]),
    "nppc_lstm_plan_destroy": (None, [_p]),
    "nppc_lstm_workspace_bytes": (_sz, [_p, _i, _i, _i]),
    "nppc_lstm_forward": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _sz, _p, _p]),
    "nppc_lstm_step_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i, _i]),
    "nppc_lstm_step_forward": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _sz, _p, _p]),
    "nppc_lstm_step_backward": (_i, [_p, _p, _i, _i, _i, _i, _p, _sz, _p, _p, _p, _p]),
    "nppc_gemm_f16_atb": (_i, [_p, _p, _ll, _i, _i, _i, _p, _p, _p]),
    "nppc_conv3x3_pack_weights": (_i, [_p, _i, _i, _i, _p, _p]),
    "nppc_conv3x3_tc": (_i, [_p, _i, _p, _i, _p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "nppc_nchw_to_nhwc_f16": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "nppc_maxpool2x2_nhwc": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "nppc_upsample2x_pad_nhwc": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "nppc_conv1x1_out": (_i, [_p, _i, _i, _i, _p, _p, _i, _p, _p]),
    "nppc_assemble_mask": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "nppc_gemm_bf16_tn": (_i, [_p, _p, _p, _p, _ll, _i, _i, _p]),
    "nppc_gemm_f16_tn": (_i, [_p, _p, _p, _p, _ll, _i, _i, _p]),
    "nppc_tcn_cl_scale": (_i, [_p, _i, _ll, _p, _p, _p]),
    "nppc_gemm_f16_tn_ex": (_i, [_p, _p, _p, _p, _ll, _i, _i, _i, _i, _p]),
    "nppc_tcn_cl_pack": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _i, _p]),
    "nppc_tcn_cl_unpack": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p]),
    "nppc_prelu_stats_cl": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "nppc_tcn_mid_cl": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p]),
    "nppc_tcn_out_cl": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _i, _p]),
}

_lib = None


class NppcError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    # Always go through build.build(): it is a sha256 stamp compare of csrc/ + the header + the flags when the library is
    # current, and a rebuild when any source changed (a stale .so from an older commit must never be called with new
    # argument lists).  Without nvcc (a deployment box that received a prebuilt .so) the stamp is still checked.
    import importlib.util
    spec = importlib.util.spec_from_file_location("_nppc_build", os.path.join(HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if os.path.exists(mod.NVCC):
        mod.build(force=False, verbose=False)
    elif not mod.is_current():
        raise NppcError(f"{LIB_PATH} does not match the sources under csrc/ (stale build) and nvcc is not available to rebuild it")
    if not os.path.exists(LIB_PATH):
        raise NppcError(f"{LIB_PATH} is missing and could not be built; the CUDA path has no fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise NppcError(f"{LIB_PATH} does not export {name}; rebuild with generative-audio_b200/build.py --force")
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().nppc_last_error().decode(errors="replace")
        # the reference raises AssertionError for its shape/argument checks (feature.py:263, fullsubnet_plus.py:157)
        if rc == -1:
            raise AssertionError(msg)
        raise NppcError(f"{what} failed ({rc}): {msg}")
