"""Shared helpers for the GPU parity tests (test infrastructure: may import oracle/)."""
import os
import tempfile

import numpy as np
import torch

import weights


def wave(B, L, seed, scale=0.05):
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((rng.standard_normal((B, L)) * scale).astype(np.float32))


_MODELS = {}


def build_model(n_dirs=5, groups=1, impl="f32", seed=0, norm_type="offline_laplace_norm"):
    """Product NPPCModel with the deterministic synthetic weights used for tests/golden."""
    key = (n_dirs, groups, impl, seed, norm_type)
    if key in _MODELS:
        return _MODELS[key]
    import generative_audio_b200 as g
    sd = weights.synth_state_dict(n_dirs, seed)
    tmp = tempfile.mkdtemp()
    ck = os.path.join(tmp, "bb.tar")
    torch.save({"model": {k[len("pretrained_restoration_model."):]: v for k, v in sd.items()
                          if k.startswith("pretrained_restoration_model.")}}, ck)
    cfg = g.NPPCModelConfig(
        pretrained_restoration_model_configuration=g.FullSubNetPlusConfig(norm_type=norm_type),
        pretrained_restoration_model_path=ck,
        audio_pc_wrapper_configuration=g.AudioPCWrapperConfig(
            multi_direction_configuration=g.MultiDirectionConfig(n_directions=n_dirs, num_groups_in_drop_band=groups,
                                                                 norm_type=norm_type)),
        stft_configuration=g.StftConfig(), device="cuda", lstm_impl=impl)
    m = g.NPPCModel(cfg)
    missing = m.load_state_dict(sd, strict=True)
    m.eval()
    _MODELS[key] = (m, sd)
    return m, sd


def pca_samples(seed=0, K=50, B=3, D=2304):
    """Seeded MC-dropout-like samples [K,B,D] with a decaying spectrum (shared by oracle/make_golden_pca.py and the test)."""
    rng = np.random.default_rng(seed)
    basis = rng.standard_normal((B, 6, D))
    coef = rng.standard_normal((K, B, 6)) * np.array([30, 18, 10, 6, 3, 1.5])
    x = np.einsum("kbj,bjd->kbd", coef, basis) + 0.05 * rng.standard_normal((K, B, D)) + rng.standard_normal((1, B, D))
    return torch.from_numpy(x.astype(np.float32))
