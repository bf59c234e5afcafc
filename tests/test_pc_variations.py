"""N1 — the validator's consumer of the principal components (nppc_audio/validator.py:55-143,246-290): oracle pinned
against a fixture made from the reference's own functions; product kernels against fixture + oracle on the GPU."""
import pytest
import torch

import nppc_oracle as O
from conftest import load_golden, rel_err

torch.set_grad_enabled(False)


def test_oracle_pc_variations_vs_golden():
    g = load_golden("fn_pc_variations")
    L = int(g["length"][0])
    pr, pi, v = O.pc_variations(g["w_mat"], g["noisy_real"], g["noisy_imag"], g["enh_real"], g["enh_imag"], g["alphas"], L)
    assert rel_err(pr, g["pc_real"]) < 1e-5 and rel_err(pi, g["pc_imag"]) < 1e-5
    assert rel_err(v, g["variations"]) < 1e-4


@pytest.mark.gpu
def test_pc_variation_kernels_vs_golden():
    import generative_audio_b200 as gab
    g = load_golden("fn_pc_variations")
    L = int(g["length"][0])
    c = {k: v.cuda() for k, v in g.items()}
    pr, pi, vr, vi = gab.ops.pc_variations(c["w_mat"], c["noisy_real"], c["noisy_imag"], c["enh_real"], c["enh_imag"], c["alphas"])
    assert rel_err(pr.cpu(), g["pc_real"]) < 1e-4 and rel_err(pi.cpu(), g["pc_imag"]) < 1e-4
    B, n, A, Fq, T = vr.shape
    waves = gab.ops.istft(vr.reshape(B * n * A, Fq, T), vi.reshape(B * n * A, Fq, T), L)
    gab.ops.peak_normalize_(waves)
    assert rel_err(waves.reshape(B, n, A, L).cpu(), g["variations"]) < 1e-4
    x = torch.randn(5, 1000, device="cuda")
    ref = x / (x.abs().amax(dim=-1, keepdim=True) + 1e-8)
    assert rel_err(gab.ops.peak_normalize_(x.clone()).cpu(), ref.cpu()) < 1e-6


@pytest.mark.gpu
def test_model_pc_variations_consistent_with_forward():
    from helpers import build_model, wave
    m, sd = build_model(5, 1, "tc")
    x = wave(2, 4096, 9).cuda()
    out = m.pc_variations(x)
    w = m(x)
    assert torch.equal(out["w_mat"], w)
    assert out["variations"].shape == (2, 5, 6, 4096) and out["enhanced"].shape == (2, 4096)
    mag, real, imag = m._stft(x)
    pr, pi, v = O.pc_variations(w.cpu(), real[:, 0].cpu(), imag[:, 0].cpu(), out["enhanced_real"].cpu(), out["enhanced_imag"].cpu(),
                                out["alphas"], 4096)
    assert rel_err(out["pc_real"].cpu(), pr) < 1e-4
    assert rel_err(out["variations"].cpu(), v) < 1e-3
    assert out["variations"].abs().amax(dim=-1).sub(1).abs().max() < 1e-5
