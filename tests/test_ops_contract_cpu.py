"""The product's REAL `ops.py` wrappers against the C-ABI CONTRACT of include/nppc_b200.h, on CPU: the shared library is replaced
by tests/cabi_emulator.py (each entry point = the oracle function the header names as its reference, applied to the caller's
memory through the documented argument list and layouts).  What this pins without a GPU: argument order and sizes, output
allocation and shapes, layout conversions ([B,1,F,T] vs [B,F,T], planar (re, im), the permutes around build_cirm), the
rc = -1 -> AssertionError path, and the scratch-based Gram-Schmidt + objective autograd Function of the audio training step.
The same reference fixtures as the `-m gpu` kernel tests (tests/test_gpu_kernels.py), which check the kernels themselves."""
import pytest
import torch

import nppc_oracle as O
from conftest import load_golden, rel_err
from helpers import wave

TOL = 1e-5


@pytest.fixture()
def ops(monkeypatch):
    import cabi_emulator
    import generative_audio_b200 as g
    cabi_emulator.install_full(monkeypatch)
    return g.ops


def test_stft_istft_wrappers(ops):
    g = load_golden("fn_stft_crm_istft")
    mag, re, im = ops.stft_mri(g["wave"])
    assert mag.shape == g["mag"].shape == (g["wave"].shape[0], 1, 257, 1 + g["wave"].shape[1] // 256)
    assert rel_err(mag, g["mag"]) < TOL and rel_err(re, g["real"]) < TOL and rel_err(im, g["imag"]) < TOL
    x = wave(2, 2300, 5)                                           # ragged length, 1-D input promoted to a batch of one
    assert ops.stft_mri(x[0])[0].shape == (1, 1, 257, 9)
    _, r, i = ops.stft_mri(x)
    y = ops.istft(r[:, 0].contiguous(), i[:, 0].contiguous(), 2300)
    assert y.shape == (2, 2300) and rel_err(y[:, :2048], x[:, :2048]) < 1e-5 and torch.all(y[:, 2048:] == 0)
    with pytest.raises(AssertionError):
        ops.stft_mri(x, n_fft=256, hop=128, win=512)


@pytest.mark.parametrize("conj", [True, False])
def test_crm_wrappers(ops, conj):
    g = load_golden("fn_stft_crm_istft")
    mag, re, im = ops.crm_decompress_apply(g["mask"], g["real"], g["imag"], conj)          # [B,1,F,T] inputs are accepted as they are
    dec = O.decompress_cirm(g["mask"].permute(0, 2, 3, 1))
    omag, ore, oim = O.crm_apply(dec[..., 0], dec[..., 1], g["real"][:, 0], g["imag"][:, 0], conj)
    assert mag.shape == ore.shape and rel_err(re, ore) < TOL and rel_err(im, oim) < TOL and rel_err(mag, omag) < TOL
    if conj:
        assert rel_err(re, g["ereal"]) < 1e-4 and rel_err(im, g["eimag"]) < 1e-4 and rel_err(mag, g["emag"]) < 1e-4
    assert rel_err(ops.decompress_cirm(g["mask"]), g["dec"].permute(0, 3, 1, 2)) < TOL
    _, cr, ci = O.stft_mri(g["clean"])
    gt = ops.build_cirm(g["real"][:, 0].contiguous(), g["imag"][:, 0].contiguous(), cr[:, 0].contiguous(), ci[:, 0].contiguous())
    assert gt.shape == g["mask"].shape and rel_err(gt, g["gt_cirm"].permute(0, 3, 1, 2)) < 1e-4


def test_norm_wrappers(ops):
    g = load_golden("fn_norm")
    assert rel_err(ops.offline_laplace_norm(g["xpos"]), g["off_pos"]) < TOL
    assert rel_err(ops.cumulative_laplace_norm(g["xpos"]), g["cum_pos"]) < TOL
    x = torch.rand(3, 1, 33, 40) + 0.05
    y = ops.pad_offline_laplace_norm(x, 2)
    assert y.shape == (3, 33, 42) and rel_err(y, O.offline_laplace_norm(torch.nn.functional.pad(x, (0, 2)))[:, 0]) < TOL
    assert torch.equal(ops.pad_offline_laplace_norm(x[:, 0].contiguous(), 2), y)           # [B,F,T] form


def test_index_wrappers_bit_exact(ops):
    g = load_golden("fn_unfold")
    for n in (15, 0, 2):
        assert torch.equal(ops.unfold(g["x"], n), g[f"n{n}"])
    with pytest.raises(AssertionError):
        ops.unfold(g["x"][0], 2)                                                            # base_model.py:26: four dims
    d = load_golden("fn_drop_band")
    for G in (1, 2, 3):
        assert torch.equal(ops.drop_band(d["x"], G), d[f"g{G}"])
    with pytest.raises(AssertionError, match="batch size should larger"):
        ops.drop_band(d["x"][:2].contiguous(), 2)                                           # feature.py:263 through rc = -1


def test_gram_schmidt_and_loss_wrappers(ops):
    g = load_golden("fn_gram_schmidt")
    out = ops.gram_schmidt_complex(g["x"])
    assert rel_err(out, g["out"]) < 1e-4 and torch.equal(out[:, 0], g["x"][:, 0])
    gr = load_golden("fn_gram_schmidt_real")
    assert rel_err(ops.gram_schmidt_real(gr["x"]), gr["out"]) < 1e-4
    gen = torch.Generator().manual_seed(6)
    head = torch.randn(3, 5, 2, 16, 17, generator=gen)
    gt = torch.randn(3, 2, 16, 17, generator=gen)
    pred = gt + 0.2 * torch.randn(3, 2, 16, 17, generator=gen) + 0.1 * head[:, 1]
    w_ref = O.gram_schmidt_complex(head.double())
    ref = O.nppc_loss(w_ref, gt.double(), pred.double(), step=250, grace=500, lambda0=1.0)
    w, st = ops.gs_loss_fused(head, gt, pred)
    st2 = ops.projection_loss(w_ref.float(), gt, pred)
    assert rel_err(w, w_ref) < 1e-4
    for s in (st, st2):
        assert s["err_proj"].is_complex() and s["err_proj"].shape == (3, 5)
        for k in ("err_norm", "w_norms", "reconst_err", "second_moment_mse"):
            assert rel_err(s[k], ref[k]) < 1e-4, k
        assert rel_err(torch.view_as_real(s["err_proj"]), torch.view_as_real(ref["err_proj"])) < 1e-4


@pytest.mark.parametrize("lam", [1e-6, 0.4, 1.0])
def test_audio_gs_loss_function_backward_through_the_wrappers(ops, lam):
    """training.GsLossFn (the audio training step's Gram-Schmidt + objective node, GPU-verified against the reference's
    gradients): nppc_gs_loss_fused's scratch -> gs_loss_fused_with_gram -> coefficient-space backward -> nppc_complex_lincomb,
    against autograd of the oracle's restatement of trainer.py:259-298 on the reference's Gram-Schmidt."""
    from generative_audio_b200 import training
    gen = torch.Generator().manual_seed(2)
    head = torch.randn(2, 5, 2, 12, 9, generator=gen)
    gt = torch.randn(2, 2, 12, 9, generator=gen)
    pred = gt + 0.3 * torch.randn(2, 2, 12, 9, generator=gen) + 0.2 * head[:, 0]
    step = {1e-6: 0, 0.4: 350, 1.0: 600}[lam]
    with torch.enable_grad():
        a = head.clone().requires_grad_(True)
        obj, w, *_ = training.GsLossFn.apply(a, gt, pred, lam)
        obj.backward()
        b = head.double().requires_grad_(True)
        xs = torch.complex(b[:, :, 0], b[:, :, 1]).flatten(2)
        outs, hats = [], []
        for i in range(5):                                                      # pc_wrapper.py:20-44 with its detach()
            v = xs[:, i]
            for h in hats:
                v = v - h * (v.conj() * h).sum(dim=1, keepdim=True)
            hats.append(v.detach() / torch.linalg.vector_norm(v.detach(), dim=1, keepdim=True))
            outs.append(v)
        wc = torch.stack(outs, 1)
        wr = torch.stack([wc.real, wc.imag], 2).reshape(b.shape)
        ref = O.nppc_loss(wr, gt.double(), pred.double(), step=step, grace=500, lambda0=1.0)
        ref["objective"].backward()
    assert abs(obj.item() - ref["objective"].item()) < 1e-5 * abs(ref["objective"].item())
    assert rel_err(w, wr.detach()) < 1e-4
    assert rel_err(a.grad, b.grad) < 1e-4
