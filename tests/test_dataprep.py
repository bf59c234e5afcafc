"""N3 — GPU data preparation: batched SNR mixing and the time -> spectrogram mask, against a fixture produced by the
reference's own dataset methods (oracle/make_golden_dataprep.py)."""
import pytest
import torch

import nppc_oracle as O
from conftest import load_golden, rel_err

torch.set_grad_enabled(False)


def _masks(g):
    L = int(g["mask_len"][0])
    m = torch.ones(g["mask_gaps"].shape[0], L)
    for b, (s, e) in enumerate(g["mask_gaps"].tolist()):
        m[b, s:e] = 0
    return m


def test_oracle_dataprep_vs_golden():
    g = load_golden("fn_dataprep")
    noisy, clean = O.mix_with_snr(g["clean"], g["noise"], g["snr"].tolist(), g["target"].tolist())
    assert rel_err(noisy, g["noisy_out"]) < 1e-6 and rel_err(clean, g["clean_out"]) < 1e-6
    T_frames, win, hop = g["stft"].tolist()
    m = _masks(g)
    assert torch.equal(O.time_to_spec_mask(m, T_frames, win, hop, True), g["spec_center"])
    assert torch.equal(O.time_to_spec_mask(m, T_frames, win, hop, False), g["spec_nocenter"])


@pytest.mark.gpu
def test_dataprep_kernels_vs_golden():
    import generative_audio_b200 as gab
    g = load_golden("fn_dataprep")
    noisy, clean = gab.ops.mix_with_snr(g["clean"].cuda(), g["noise"].cuda(), g["snr"].cuda(), g["target"].cuda())
    assert rel_err(noisy.cpu(), g["noisy_out"]) < 1e-5 and rel_err(clean.cpu(), g["clean_out"]) < 1e-5
    assert float(noisy[3].abs().max()) <= 0.99 + 1e-6          # the clipping-prevention branch
    T_frames, win, hop = g["stft"].tolist()
    m = _masks(g).cuda()
    assert torch.equal(gab.ops.time_to_spec_mask(m, T_frames, win, hop, True).cpu(), g["spec_center"])     # bit-exact
    assert torch.equal(gab.ops.time_to_spec_mask(m, T_frames, win, hop, False).cpu(), g["spec_nocenter"])
