"""CPU checks of the hand-written training step's HOST-side mathematics (no GPU needed):
  * gs_backward.gs_loss_grad_coeffs — backward of Gram-Schmidt (conjugated coefficient, detached normaliser: pc_wrapper.py:8-44)
    + the NPPC objective (trainer.py:259-298,337-342, detached projection) in coefficient space — against torch.autograd of a
    literal restatement of the reference formulas;
  * training._dropband_maps — the row <-> (sample, frequency) permutation of drop_band (feature.py:254-285) against the oracle;
  * training._tsse_gate — the convolution-free TSSE squeeze (windowed sums) against the oracle's depthwise-conv restatement;
  * bench.py --impl reference — honours --steps / --warmup, prints our arm's config and says which sample it ran."""
import json
import os
import subprocess
import sys

import pytest
import torch

import nppc_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gs_ref(x):
    n = x.shape[1]
    v = torch.complex(x[:, :, 0], x[:, :, 1])
    outs, hats = [], []
    for i in range(n):
        w = v[:, i]
        for wh in hats:
            w = w - wh * (w.conj() * wh).sum(dim=1, keepdim=True)
        wd = w.detach()
        hats.append(wd / torch.linalg.vector_norm(wd, dim=1, keepdim=True))
        outs.append(w)
    out = torch.stack(outs, 1)
    return torch.stack([out.real, out.imag], 2)


def _loss_ref(w, gt, pred, lam):
    B, n = w.shape[:2]
    W = w.reshape(B, n, 2, -1)
    wn = torch.linalg.vector_norm(W, dim=(2, 3))
    wh = W / (wn[..., None, None] + 1e-8)
    err = (gt - pred).reshape(B, 2, -1)
    en = torch.linalg.vector_norm(err, dim=(1, 2))
    err = err / (en[:, None, None] + 1e-8)
    wn = wn / (en[:, None] + 1e-8)
    ep = (torch.complex(wh[:, :, 0], wh[:, :, 1]).conj() * torch.complex(err[:, 0], err[:, 1])[:, None]).sum(-1)
    mag = ep.abs()
    return (1 - mag.pow(2).sum(1)).mean() + lam * (wn.pow(2) - mag.detach().pow(2)).pow(2).mean()


@pytest.mark.parametrize("B,n,P,lam", [(3, 5, 200, 0.37), (1, 1, 50, 1.0), (2, 10, 64, 1e-6)])
def test_gs_loss_backward_in_coefficient_space_matches_autograd(B, n, P, lam):
    from generative_audio_b200.gs_backward import gs_loss_grad_coeffs
    g = torch.Generator().manual_seed(B * 100 + n)
    x = torch.randn(B, n, 2, P, dtype=torch.float64, generator=g)
    if n > 2:
        x[:, 2] = 0.6 * x[:, 0] + 0.4 * x[:, 2]            # correlated directions: the projections matter
    gt = torch.randn(B, 2, P, dtype=torch.float64, generator=g)
    pred = torch.randn(B, 2, P, dtype=torch.float64, generator=g)
    with torch.enable_grad():
        xr = x.clone().requires_grad_(True)
        _loss_ref(_gs_ref(xr), gt, pred, lam).backward()
    xc = torch.complex(x[:, :, 0], x[:, :, 1])
    e = torch.complex((gt - pred)[:, 0], (gt - pred)[:, 1])
    V = torch.cat([xc, e[:, None]], 1)
    G = torch.einsum("bjp,bkp->bjk", V.conj(), V)                                   # what the forward kernel's scratch holds
    with torch.no_grad():
        w = _gs_ref(x)
    wc = torch.complex(w[:, :, 0], w[:, :, 1])
    A = torch.linalg.lstsq(xc.transpose(1, 2), wc.transpose(1, 2)).solution.transpose(1, 2)   # w_i = sum_k A_ik x_k
    coef = gs_loss_grad_coeffs(G, A, lam)
    d = torch.einsum("bik,bkp->bip", coef, V)                                       # what nppc_complex_lincomb streams out
    got = torch.stack([d.real, d.imag], 2)
    assert ((got - xr.grad).abs().max() / xr.grad.abs().max()).item() < 1e-9
    # a device-scalar lambda and an upstream gradient scale go through unchanged
    coef2 = gs_loss_grad_coeffs(G, A, torch.tensor(lam, dtype=torch.float64), grad_scale=torch.tensor(2.0, dtype=torch.float64))
    assert torch.allclose(coef2, 2 * coef)


@pytest.mark.parametrize("B,Fq,G", [(4, 257, 2), (5, 11, 3), (3, 8, 1), (1, 9, 2)])
def test_dropband_row_maps_match_reference_order(B, Fq, G):
    from generative_audio_b200.training import _dropband_maps
    x = (torch.arange(B)[:, None] * 1000 + torch.arange(Fq)[None, :]).float()[:, None, :, None]    # [B,1,F,1]: value = 1000 b + f
    ref = O.drop_band(x, G) if B > 1 else x
    Ge = G if (G > 1 and B > 1) else 1
    sb, f, Fg = _dropband_maps(B, Fq, Ge, "cpu")
    assert Fg == ref.shape[2] and sb.numel() == ref.shape[0] * Fg
    assert torch.equal((sb * 1000 + f).float(), ref[:, 0, :, 0].reshape(-1))
    assert len({(int(a), int(b)) for a, b in zip(sb, f)}) == sb.numel()     # unique pairs: the backward scatter needs no atomics


def test_tsse_gate_without_convolutions_matches_oracle():
    import generative_audio_b200 as G
    import weights
    from generative_audio_b200.training import _tsse_gate
    sd = weights.synth_state_dict(5, 0, "audio_pc_wrapper.net.")
    att = G.modules.ChannelTimeSenseSELayer(257)
    att.load_state_dict({k[len("channel_attention."):]: v for k, v in sd.items() if k.startswith("channel_attention.")})
    x = torch.randn(2, 257, 40, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        got = x * _tsse_gate(att, x)[:, :, None]
        ref = O.tsse(x, sd, "channel_attention")
    assert ((got - ref).abs().max() / ref.abs().max()).item() < 1e-5


def test_reference_arm_reports_what_it_ran():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "0", "--cpu-batch", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 0          # K and W honoured, not clamped
    assert line["config"]["batch_per_gpu"] == 64                                               # our arm's config ...
    assert line["sample"]["batch_per_step"] == 1 and line["sample"]["of_batch"] == 64           # ... and the sample that actually ran
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0


def test_training_path_refuses_configurations_it_does_not_differentiate():
    """the differentiable head must not silently optimise a different network than inference evaluates — a
    norm_type it has no backward for raises before any kernel runs."""
    import types

    from generative_audio_b200 import training
    net = types.SimpleNamespace(norm_type="cumulative_laplace_norm", look_ahead=2)
    x = torch.zeros(2, 1, 257, 5)
    with pytest.raises(NotImplementedError, match="offline_laplace_norm"):
        training.head_forward_train(net, x, x, x, x, x, x)


def _gs_real_ref(x):
    """inpainting/nppc/pc_wrapper.py:43-59, literally."""
    sh = x.shape
    x = x.flatten(2)
    outs, hats = [], []
    for i in range(x.shape[1]):
        w = x[:, i, :]
        for w2 in hats:
            w = w - w2 * torch.sum(w * w2, dim=-1, keepdim=True)
        hats.append(w.detach() / w.detach().norm(dim=-1, keepdim=True))
        outs.append(w)
    return torch.stack(outs, dim=1).view(*sh)


def _loss_real_ref(w_mat, clean, pred, lam):
    """inpainting/trainer/nppc_trainer.py:347-373, literally (objective = reconst.mean() + lam * second_moment.mean())."""
    w_ = w_mat.flatten(2)
    w_norms = w_.norm(dim=2) + 1e-6
    w_hat = w_ / w_norms[:, :, None]
    err = (clean - pred).flatten(1)
    err_norm = err.norm(dim=1) + 1e-6
    err = err / err_norm[:, None]
    w_norms = w_norms / err_norm[:, None]
    err_proj = torch.einsum("bki,bi->bk", w_hat, err)
    reconst = 1 - err_proj.pow(2).sum(dim=1)
    second = (w_norms.pow(2) - err_proj.detach().pow(2)).pow(2)
    return reconst.mean() + lam * second.mean()


@pytest.mark.parametrize("B,n,P,lam", [(3, 5, 200, 0.37), (1, 1, 50, 1.0), (2, 10, 64, 1e-6), (2, 3, 30, 1.0)])
def test_real_gs_loss_backward_in_coefficient_space_matches_autograd(B, n, P, lam):
    """The inpainting head's backward (real Gram-Schmidt with detached normalisers + the 1e-6-regularised projection /
    second-moment objective) from the Gram matrix and the coefficient matrix the forward kernel leaves in its scratch."""
    from generative_audio_b200.gs_backward import gs_loss_grad_coeffs
    g = torch.Generator().manual_seed(B * 100 + n + 7)
    x = torch.randn(B, n, 1, P, dtype=torch.float64, generator=g) * 0.3        # small norms: the 1e-6 terms are visible
    if n > 2:
        x[:, 2] = 0.6 * x[:, 0] + 0.4 * x[:, 2]
    clean = torch.randn(B, 1, 1, P, dtype=torch.float64, generator=g)
    pred = torch.randn(B, 1, 1, P, dtype=torch.float64, generator=g)
    with torch.enable_grad():
        xr = x.clone().requires_grad_(True)
        _loss_real_ref(_gs_real_ref(xr), clean, pred, lam).backward()
    X = x.flatten(2)
    V = torch.cat([X, (clean - pred).flatten(1)[:, None]], 1)
    G = torch.einsum("bjp,bkp->bjk", V, V)
    with torch.no_grad():
        W = _gs_real_ref(x).flatten(2)
    A = torch.linalg.lstsq(X.transpose(1, 2), W.transpose(1, 2)).solution.transpose(1, 2)
    coef = gs_loss_grad_coeffs(G, A, lam, real=True)
    assert not coef.is_complex()
    got = torch.einsum("bik,bkp->bip", coef, V).view_as(x)
    assert ((got - xr.grad).abs().max() / xr.grad.abs().max()).item() < 1e-9


@pytest.mark.parametrize("B,n,P", [(2, 5, 120), (1, 1, 30), (3, 6, 64)])
def test_gram_schmidt_backward_for_an_arbitrary_upstream_gradient(B, n, P):
    """gs_backward.gs_grad_coeffs: the backward of Gram-Schmidt alone (what NPPCModel's differentiable forward needs when the
    loss is written in torch by the caller, as in the reference trainer) against autograd, complex and real."""
    from generative_audio_b200.gs_backward import gs_grad_coeffs
    g = torch.Generator().manual_seed(B * 10 + n)
    x = torch.randn(B, n, 2, P, dtype=torch.float64, generator=g)
    up = torch.randn(B, n, 2, P, dtype=torch.float64, generator=g)               # arbitrary d L / d w
    with torch.enable_grad():
        xr = x.clone().requires_grad_(True)
        (_gs_ref(xr) * up).sum().backward()
    xc, gc = torch.complex(x[:, :, 0], x[:, :, 1]), torch.complex(up[:, :, 0], up[:, :, 1])
    V = torch.cat([xc, gc], 1)
    G2 = torch.einsum("bjp,bkp->bjk", V.conj(), V)
    with torch.no_grad():
        w = _gs_ref(x)
    wc = torch.complex(w[:, :, 0], w[:, :, 1])
    A = torch.linalg.lstsq(xc.transpose(1, 2), wc.transpose(1, 2)).solution.transpose(1, 2)
    d = torch.einsum("bik,bkp->bip", gs_grad_coeffs(G2, A), V)
    got = torch.stack([d.real, d.imag], 2)
    assert ((got - xr.grad).abs().max() / xr.grad.abs().max()).item() < 1e-9
    # real variant (inpainting head)
    xre = torch.randn(B, n, 1, P, dtype=torch.float64, generator=g)
    upr = torch.randn(B, n, 1, P, dtype=torch.float64, generator=g)
    with torch.enable_grad():
        xg = xre.clone().requires_grad_(True)
        (_gs_real_ref(xg) * upr).sum().backward()
    Vr = torch.cat([xre.flatten(2), upr.flatten(2)], 1)
    with torch.no_grad():
        Wr = _gs_real_ref(xre).flatten(2)
    Ar = torch.linalg.lstsq(xre.flatten(2).transpose(1, 2), Wr.transpose(1, 2)).solution.transpose(1, 2)
    dr = torch.einsum("bik,bkp->bip", gs_grad_coeffs(torch.einsum("bjp,bkp->bjk", Vr, Vr), Ar), Vr).view_as(xre)
    assert ((dr - xg.grad).abs().max() / xg.grad.abs().max()).item() < 1e-9


def test_differentiable_gram_schmidt_function_glue_on_cpu(monkeypatch):
    """training.GramSchmidtFn with its three kernel calls replaced by CPU restatements: the Function's own plumbing (stacking
    [x; g], padding the coefficient matrix to the lincomb kernel's [2n, 2n+1] shape, slicing the n useful rows) reproduces
    autograd of the reference's Gram-Schmidt for a loss written in torch.  TEST-ONLY stubs."""
    import generative_audio_b200 as G
    from generative_audio_b200 import training
    ops = G.ops

    def cplx(t):
        return torch.complex(t[:, :, 0].double(), t[:, :, 1].double()).flatten(2)

    def with_coeffs(x):
        w = _gs_ref(x.double().flatten(3)).view_as(x)
        xc, wc = cplx(x), cplx(w)
        A = torch.linalg.lstsq(xc.transpose(1, 2), wc.transpose(1, 2)).solution.transpose(1, 2)
        return w.to(x.dtype), torch.einsum("bjp,bkp->bjk", xc.conj(), xc), A.to(torch.complex64)

    def lincomb(x, gt, pred, coef):
        B, n = x.shape[:2]
        assert tuple(coef.shape) == (B, n, n + 1) and gt.shape == pred.shape == (B, *x.shape[2:])
        e = torch.complex((gt - pred)[:, 0].double(), (gt - pred)[:, 1].double()).flatten(1)
        d = torch.einsum("bik,bkp->bip", coef.to(torch.complex128), torch.cat([cplx(x), e[:, None]], 1))
        return torch.stack([d.real, d.imag], 2).reshape(x.shape).to(x.dtype)

    monkeypatch.setattr(ops, "gram_schmidt_complex", lambda x: with_coeffs(x)[0])
    monkeypatch.setattr(ops, "gram_matrix_complex", lambda v: with_coeffs(v)[1])
    monkeypatch.setattr(ops, "complex_lincomb", lincomb)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 5, 2, 6, 9, dtype=torch.float64, generator=g)
    tgt = torch.randn(2, 5, 2, 6, 9, dtype=torch.float64, generator=g)
    with torch.enable_grad():
        a = x.clone().requires_grad_(True)
        w = training.GramSchmidtFn.apply(a)
        ((w - tgt) ** 2).sum().backward()                                           # any torch loss, as the reference trainer writes it
        b = x.clone().requires_grad_(True)
        wr = _gs_ref(b.flatten(3)).view_as(b)
        ((wr - tgt) ** 2).sum().backward()
    assert ((w.detach() - wr.detach()).abs().max() / wr.abs().max()).item() < 1e-12
    assert ((a.grad - b.grad).abs().max() / b.grad.abs().max()).item() < 1e-9       # coefficients replayed in fp64 from the Gram matrix
    with pytest.raises(NotImplementedError):
        training.GramSchmidtFn.apply(torch.zeros(1, 7, 2, 4, 4))


def test_gram_schmidt_scratch_decoders_follow_the_struct_layout():
    """ops._decode_gs_scratch(_real): the backward reads the Gram and coefficient matrices out of the forward kernel's scratch,
    `struct SampleScratch { double G[13*13*2]; float A[12*12*2]; }` per sample (csrc/gram_schmidt.cu: G row-major complex,
    UPPER triangle only, entry (j,k) at (j*13+k)*2 + {re,im}; A row i = direction i at (i*12+k)*2 + {re,im}).  A scratch buffer
    packed here with numpy in exactly that layout must come back as the full Hermitian / symmetric matrices."""
    import numpy as np

    import generative_audio_b200 as G_
    ops = G_.ops
    B, n = 3, 5
    rng = np.random.default_rng(0)
    assert ops.GS_SCRATCH_BYTES == 13 * 13 * 2 * 8 + 12 * 12 * 2 * 4
    for real in (False, True):
        nv = n + 1 if real else n
        M = rng.standard_normal((B, nv, nv)) + (0 if real else 1j) * rng.standard_normal((B, nv, nv))
        H = M + np.conj(np.transpose(M, (0, 2, 1)))                      # Hermitian (symmetric when real)
        A = np.tril(rng.standard_normal((B, n, n)) + (0 if real else 1j) * rng.standard_normal((B, n, n)))
        buf = np.zeros((B + 2, ops.GS_SCRATCH_BYTES), dtype=np.uint8)    # + trailing bytes, as nppc_gs_scratch_bytes over-allocates
        for b in range(B):
            Gs = np.full((13, 13, 2), 777.0)                             # garbage where the kernel never writes ...
            As = np.full((12, 12, 2), 555.0, dtype=np.float32)
            for j in range(nv):
                for k in range(j, nv):                                   # ... the upper triangle is all it fills
                    Gs[j, k] = (H[b, j, k].real, H[b, j, k].imag)
            Gs[np.arange(nv), np.arange(nv), 1] = 1e-9                   # rounding dust in Im of the diagonal must be dropped
            As[:n, :n, 0], As[:n, :n, 1] = A[b].real, A[b].imag
            buf[b, :13 * 13 * 16] = np.frombuffer(Gs.tobytes(), dtype=np.uint8)
            buf[b, 13 * 13 * 16:] = np.frombuffer(As.tobytes(), dtype=np.uint8)
        scr = torch.from_numpy(buf.reshape(-1))
        if real:
            Gd, Ad = ops._decode_gs_scratch_real(scr, B, n, nv)
            assert Gd.dtype == torch.float64 and Ad.dtype == torch.float32
        else:
            Gd, Ad = ops._decode_gs_scratch(scr, B, n)
            assert Gd.dtype == torch.complex128 and Ad.dtype == torch.complex64
            H = H - 1j * np.imag(H) * np.eye(nv)[None]
        assert np.allclose(Gd.numpy(), H, atol=1e-12)
        assert np.allclose(Ad.numpy(), A.astype(np.complex64 if not real else np.float32), atol=1e-6)


def test_differentiable_gram_schmidt_through_the_real_ops_wrappers_on_an_emulated_cabi(monkeypatch):
    """training.GramSchmidtFn through the product's real ops wrappers (gram_schmidt_complex_with_coeffs, gram_matrix_complex,
    complex_lincomb) with only the shared library replaced by tests/cabi_emulator.py: marshalling + scratch decode + the
    stacked [x; g] Gram pass + the padded coefficient matrix, against autograd of the reference's Gram-Schmidt."""
    import cabi_emulator
    from generative_audio_b200 import training
    lib = cabi_emulator.install(monkeypatch)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 5, 2, 6, 10, generator=g)
    tgt = torch.randn(2, 5, 2, 6, 10, generator=g)
    with torch.enable_grad():
        a = x.clone().requires_grad_(True)
        w = training.GramSchmidtFn.apply(a)
        ((w - tgt) ** 2).sum().backward()
        b = x.double().requires_grad_(True)
        wr = _gs_ref(b.flatten(3)).view_as(b)
        ((wr - tgt.double()) ** 2).sum().backward()
    assert lib.calls == ["nppc_gram_schmidt_complex", "nppc_gram_schmidt_complex", "nppc_complex_lincomb"]
    assert ((w.detach().double() - wr.detach()).abs().max() / wr.abs().max()).item() < 1e-5
    assert torch.equal(w.detach()[:, 0], x[:, 0])
    assert ((a.grad.double() - b.grad).abs().max() / b.grad.abs().max()).item() < 1e-4


def test_gram_schmidt_to_crm_keeps_the_graph_when_its_input_has_one(monkeypatch):
    """The reference's gram_schmidt_to_crm is differentiable; the drop-in must not silently return a graph-less tensor for an
    input that requires grad (emulated C ABI; plain kernel path when no graph is involved)."""
    import cabi_emulator
    import generative_audio_b200 as G
    lib = cabi_emulator.install(monkeypatch)
    x = torch.randn(1, 3, 2, 4, 5, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        w0 = G.gram_schmidt_to_crm(x)
    assert not w0.requires_grad and lib.calls == ["nppc_gram_schmidt_complex"]
    with torch.enable_grad():
        xr = x.clone().requires_grad_(True)
        w = G.gram_schmidt_to_crm(xr)
        assert w.requires_grad and torch.equal(w.detach(), w0)
        w.square().sum().backward()
        assert xr.grad is not None and xr.grad.abs().max() > 0
        with pytest.raises(NotImplementedError):
            G.gram_schmidt_to_crm(torch.zeros(1, 7, 2, 4, 5, requires_grad=True))


def test_real_gram_schmidt_keeps_the_graph_and_matches_autograd(monkeypatch):
    """gram_schmidt_to_spec_mag (inpainting): differentiable for an arbitrary torch loss through GramSchmidtRealFn — real wrappers
    on the emulated C ABI (Gram pass over the stacked [x; g] via nppc_gram_schmidt_real, linear combination on the two planes of
    nppc_complex_lincomb) against autograd of the reference's real Gram-Schmidt."""
    import cabi_emulator
    import generative_audio_b200 as G
    lib = cabi_emulator.install(monkeypatch)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 5, 6, 10, generator=g)
    tgt = torch.randn(2, 5, 6, 10, generator=g)
    with torch.enable_grad():
        a = x.clone().requires_grad_(True)
        w = G.gram_schmidt_to_spec_mag(a)
        ((w - tgt) ** 2).sum().backward()
        b = x.double().requires_grad_(True)
        wr = _gs_real_ref(b)
        ((wr - tgt.double()) ** 2).sum().backward()
    assert lib.calls == ["nppc_gram_schmidt_real", "nppc_gram_schmidt_real", "nppc_complex_lincomb"]
    assert ((w.detach().double() - wr.detach()).abs().max() / wr.abs().max()).item() < 1e-5
    assert ((a.grad.double() - b.grad).abs().max() / b.grad.abs().max()).item() < 1e-4
    with torch.no_grad():
        assert not G.inpainting.gram_schmidt_to_spec_mag(x).requires_grad
