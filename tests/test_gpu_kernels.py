"""GPU parity tests, kernel by kernel: every call goes through the C ABI (libnppc_b200.so via ctypes) and is
compared with the CPU oracle on the same seeded inputs and with the reference-generated golden fixtures.
Tolerances: bit-exact for index kernels, rel 1e-4 (max-norm) for fp32 kernels."""
import pytest
import torch

import nppc_oracle as O
from conftest import load_golden, rel_err
from helpers import wave

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ops():
    import generative_audio_b200 as g
    return g.ops


def cu(t):
    return t.cuda()


@pytest.mark.parametrize("B,L", [(1, 64000), (3, 4096), (2, 2048), (2, 700), (1, 64100)])
def test_stft(ops, B, L):
    x = wave(B, L, 3)
    mag, re, im = ops.stft_mri(cu(x))
    omag, ore, oim = O.stft_mri(x)
    assert mag.shape == omag.shape
    # fp64 oracle arbitrates: our error vs fp64 must be of the same order as torch's fp32 error
    dmag, dre, dim_ = O.stft_mri(x.double())
    for a, b, d in ((mag, omag, dmag), (re, ore, dre), (im, oim, dim_)):
        assert rel_err(a.cpu(), b) < TOL
        assert rel_err(a.cpu(), d) < TOL


def test_stft_golden(ops):
    g = load_golden("fn_stft_crm_istft")
    mag, re, im = ops.stft_mri(cu(g["wave"]))
    assert rel_err(mag.cpu(), g["mag"]) < TOL and rel_err(re.cpu(), g["real"]) < TOL and rel_err(im.cpu(), g["imag"]) < TOL


def test_stft_rejects_bad_args(ops):
    with pytest.raises(AssertionError):
        ops.stft_mri(cu(wave(1, 4096, 1)), n_fft=256, hop=128, win=256)
    with pytest.raises(RuntimeError):
        ops.stft_mri(wave(1, 4096, 1))  # CPU tensor: no CPU fallback


@pytest.mark.parametrize("B,L,length", [(2, 4096, 4096), (1, 64000, 64000), (2, 2048, 2000), (2, 2048, 2300)])
def test_istft(ops, B, L, length):
    x = wave(B, L, 5)
    _, re, im = O.stft_mri(x)
    g = torch.Generator().manual_seed(0)
    re = re[:, 0] * (1 + 0.3 * torch.randn(re[:, 0].shape, generator=g))
    im = im[:, 0] + 0.1 * torch.randn(im[:, 0].shape, generator=g)  # non-zero Im at DC/Nyquist on purpose
    if length <= 256 * (re.shape[-1] - 1):
        ref = torch.istft(torch.complex(re, im), 512, 256, 512, torch.hann_window(512), center=True, length=length)
        assert rel_err(O.istft(re, im, length), ref) < 1e-5
    out = ops.istft(cu(re.contiguous()), cu(im.contiguous()), length)
    assert rel_err(out.cpu(), O.istft(re, im, length)) < TOL


def test_stft_istft_round_trip(ops):
    x = wave(4, 64000, 9)
    _, re, im = ops.stft_mri(cu(x))
    y = ops.istft(re[:, 0].contiguous(), im[:, 0].contiguous(), 64000)
    assert rel_err(y.cpu(), x) < 1e-5


@pytest.mark.parametrize("conj", [True, False])
def test_crm_decompress_apply(ops, conj):
    g = load_golden("fn_stft_crm_istft")
    mag, re, im = ops.crm_decompress_apply(cu(g["mask"]), cu(g["real"][:, 0].contiguous()), cu(g["imag"][:, 0].contiguous()), conj)
    dec = O.decompress_cirm(g["mask"].permute(0, 2, 3, 1))
    omag, ore, oim = O.crm_apply(dec[..., 0], dec[..., 1], g["real"][:, 0], g["imag"][:, 0], conj)
    assert rel_err(re.cpu(), ore) < TOL and rel_err(im.cpu(), oim) < TOL and rel_err(mag.cpu(), omag) < TOL
    if conj:
        assert rel_err(re.cpu(), g["ereal"]) < TOL and rel_err(im.cpu(), g["eimag"]) < TOL and rel_err(mag.cpu(), g["emag"]) < TOL


def test_decompress_and_build_cirm(ops):
    g = load_golden("fn_stft_crm_istft")
    d = ops.decompress_cirm(cu(g["mask"]))
    assert rel_err(d.cpu(), g["dec"].permute(0, 3, 1, 2)) < TOL
    _, cr, ci = O.stft_mri(g["clean"])
    gt = ops.build_cirm(cu(g["real"][:, 0].contiguous()), cu(g["imag"][:, 0].contiguous()), cu(cr[:, 0].contiguous()),
                        cu(ci[:, 0].contiguous()))
    assert rel_err(gt.cpu(), g["gt_cirm"].permute(0, 3, 1, 2)) < TOL


def test_offline_laplace_norm(ops):
    g = load_golden("fn_norm")
    assert rel_err(ops.offline_laplace_norm(cu(g["xpos"])).cpu(), g["off_pos"]) < TOL
    # signed input: the divisor is a cancelling sum; compare against the fp64 oracle
    ref64 = O.offline_laplace_norm(g["xsigned"].double())
    assert rel_err(ops.offline_laplace_norm(cu(g["xsigned"])).cpu(), ref64) < TOL
    x = torch.rand(3, 1, 257, 251) + 0.05
    y = ops.pad_offline_laplace_norm(cu(x), 2)
    ref = O.offline_laplace_norm(torch.nn.functional.pad(x, (0, 2)))[:, 0]
    assert y.shape == ref.shape and rel_err(y.cpu(), ref) < TOL
    big = torch.rand(2, 257, 34, 253)
    assert rel_err(ops.offline_laplace_norm(cu(big)).cpu(), O.offline_laplace_norm(big)) < TOL


def test_cumulative_laplace_norm(ops):
    g = load_golden("fn_norm")
    assert rel_err(ops.cumulative_laplace_norm(cu(g["xpos"])).cpu(), g["cum_pos"]) < TOL
    x = torch.rand(2, 3, 257, 600) + 0.05  # T > block size: exercises the carry across scan tiles
    assert rel_err(ops.cumulative_laplace_norm(cu(x)).cpu(), O.cumulative_laplace_norm(x)) < TOL


def test_unfold_bit_exact(ops):
    g = load_golden("fn_unfold")
    for n in (15, 0, 2):
        assert torch.equal(ops.unfold(cu(g["x"]), n).cpu(), g[f"n{n}"])
    x = torch.randn(2, 2, 257, 253)
    assert torch.equal(ops.unfold(cu(x), 15).cpu(), O.unfold(x, 15))


def test_drop_band_bit_exact(ops):
    g = load_golden("fn_drop_band")
    for G in (1, 2, 3):
        assert torch.equal(ops.drop_band(cu(g["x"]), G).cpu(), g[f"g{G}"])
    with pytest.raises(AssertionError):
        ops.drop_band(cu(g["x"][:2].contiguous()), 2)  # B must be > groups (feature.py:263)
    with pytest.raises(AssertionError):
        ops.drop_band(cu(g["x"][:1].contiguous()), 1)


def test_gram_schmidt_complex(ops):
    g = load_golden("fn_gram_schmidt")
    out = ops.gram_schmidt_complex(cu(g["x"]))
    assert rel_err(out.cpu(), g["out"]) < TOL
    assert torch.equal(out[:, 0].cpu(), g["x"][:, 0])
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(3, 5, 2, 257, 251, generator=gen)
    x[:, 3] = 0.9 * x[:, 1] + 0.1 * x[:, 3]
    ref = O.gram_schmidt_complex(x.double())
    assert rel_err(ops.gram_schmidt_complex(cu(x)).cpu(), ref) < TOL
    for n in (1, 2, 8, 10, 12):
        xs = torch.randn(2, n, 2, 40, 33, generator=gen)
        assert rel_err(ops.gram_schmidt_complex(cu(xs)).cpu(), O.gram_schmidt_complex(xs.double())) < TOL, n


def test_gram_schmidt_real(ops):
    g = load_golden("fn_gram_schmidt_real")
    assert rel_err(ops.gram_schmidt_real(cu(g["x"])).cpu(), g["out"]) < TOL
    x = torch.randn(2, 10, 128, 500)
    assert rel_err(ops.gram_schmidt_real(cu(x)).cpu(), O.gram_schmidt_real(x.double())) < TOL


def test_projection_loss_and_fused(ops):
    gen = torch.Generator().manual_seed(6)
    head = torch.randn(4, 5, 2, 128, 17, generator=gen)
    gt = torch.randn(4, 2, 128, 17, generator=gen)
    pred = gt + 0.2 * torch.randn(4, 2, 128, 17, generator=gen) + 0.1 * head[:, 1]
    w_ref = O.gram_schmidt_complex(head.double())
    ref = O.nppc_loss(w_ref, gt.double(), pred.double(), step=250, grace=500, lambda0=1.0)
    w, st = ops.gs_loss_fused(cu(head), cu(gt), cu(pred))
    assert rel_err(w.cpu(), w_ref) < TOL
    st2 = ops.projection_loss(cu(w_ref.float()), cu(gt), cu(pred))
    for s in (st, st2):
        assert rel_err(s["err_norm"].cpu(), ref["err_norm"]) < TOL
        assert rel_err(s["w_norms"].cpu(), ref["w_norms"]) < TOL
        assert rel_err(torch.view_as_real(s["err_proj"].cpu()), torch.view_as_real(ref["err_proj"])) < 1e-3
        assert rel_err(s["reconst_err"].cpu(), ref["reconst_err"]) < TOL
        assert rel_err(s["second_moment_mse"].cpu(), ref["second_moment_mse"]) < 1e-3


@pytest.mark.parametrize("B,G", [(1, 1), (3, 1), (4, 2), (5, 2)])
def test_subband_pack(ops, B, G):
    gen = torch.Generator().manual_seed(8)
    Tp = 37
    nbr, fb, fbr, fbi = (torch.rand(B, 257, Tp, generator=gen) for _ in range(4))
    parts = [O.unfold(nbr[:, None], 15).reshape(B, 257, 31, Tp)] + [v[:, :, None] for v in (fb, fbr, fbi)]
    sb = O.offline_laplace_norm(torch.cat(parts, dim=2))
    if B > 1:
        sb = O.drop_band(sb.permute(0, 2, 1, 3), G).permute(0, 2, 1, 3)
    ref = sb.reshape(-1, 34, Tp).permute(2, 0, 1)  # [T', R, 34]
    xs, R = ops.subband_pack(cu(nbr), cu(fb), cu(fbr), cu(fbi), 15, G, 64, torch.float32)
    xs = xs.cpu()
    assert R == ref.shape[1] and xs.shape == (Tp, R, 64)
    assert rel_err(xs[:, :, :34], ref) < TOL
    assert torch.count_nonzero(xs[:, :, 34:]) == 0
    xb, R2 = ops.subband_pack(cu(nbr), cu(fb), cu(fbr), cu(fbi), 15, G, 64, torch.float16)
    xb = xb.cpu().float()
    assert R2 == R and xb.shape[1] % 128 == 0 and xb.shape[1] >= R
    assert rel_err(xb[:, :R, :34], ref) < 1e-2
    assert torch.count_nonzero(xb[:, R:, :]) == 0


@pytest.mark.parametrize("B,G,Tp", [(2, 1, 37), (3, 2, 70), (1, 1, 253)])
def test_subband_pack_cumulative_norm(ops, B, G, Tp):
    """norm_type = cumulative_laplace_norm fused into the packer (base_model.py:227-257 on the [B,F,S,T'] tensor): the running
    mean crosses tile boundaries (Tp > 32) and, for signed inputs, zero — the fp64 oracle arbitrates like in the model test."""
    gen = torch.Generator().manual_seed(18)
    nbr, fb, fbr, fbi = (torch.rand(B, 257, Tp, generator=gen) + 0.05 for _ in range(4))
    def ref_of(dt):
        parts = [O.unfold(nbr.to(dt)[:, None], 15).reshape(B, 257, 31, Tp)] + [v.to(dt)[:, :, None] for v in (fb, fbr, fbi)]
        sb = O.cumulative_laplace_norm(torch.cat(parts, dim=2))
        if B > 1:
            sb = O.drop_band(sb.permute(0, 2, 1, 3), G).permute(0, 2, 1, 3)
        return sb.reshape(-1, 34, Tp).permute(2, 0, 1)
    ref64, ref32 = ref_of(torch.float64), ref_of(torch.float32)
    xs, R = ops.subband_pack(cu(nbr), cu(fb), cu(fbr), cu(fbi), 15, G, 64, torch.float32, cumulative=True)
    xs = xs.cpu()
    assert R == ref64.shape[1] and xs.shape == (Tp, R, 64)
    assert rel_err(xs[:, :, :34], ref64) < max(TOL, 2 * rel_err(ref32, ref64))
    assert torch.count_nonzero(xs[:, :, 34:]) == 0
    xh, _ = ops.subband_pack(cu(nbr), cu(fb), cu(fbr), cu(fbi), 15, G, 64, torch.float16, cumulative=True)
    assert rel_err(xh.cpu().float()[:, :R, :34], ref64) < 1e-2 and torch.count_nonzero(xh[:, R:, :]) == 0


def test_lstm_f32(ops):
    import weights
    p = weights.synth_state_dict(5, 0, "pretrained_restoration_model.")
    gen = torch.Generator().manual_seed(10)
    R, Tp = 50, 23
    x = torch.randn(R, 34, Tp, generator=gen)
    ref = O.lstm_fc(x, p, "sb_model")
    lp = "sb_model.sequence_model."
    plan = ops.LstmPlan(*[cu(p[lp + f"{k}_l{l}"]) for l in (0, 1) for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")],
                        cu(p["sb_model.fc_output_layer.weight"]), cu(p["sb_model.fc_output_layer.bias"]))
    xs = torch.zeros(Tp, R, 64)
    xs[:, :, :34] = x.permute(2, 0, 1)
    y = plan.forward(cu(xs), 0)
    assert y.shape == ref.shape
    assert rel_err(y.cpu(), ref) < TOL


def test_assemble_mask(ops):
    y = torch.randn(3 * 7, 4, 11)
    ref = y.reshape(3, 7, 4, 11).permute(0, 2, 1, 3)[..., 2:]
    assert torch.equal(ops.assemble_mask(cu(y), 3, 7, 2).cpu(), ref)


def test_tsse_vs_oracle(ops):
    import weights
    import generative_audio_b200 as g
    p = weights.synth_state_dict(5, 0, "pretrained_restoration_model.")
    gen = torch.Generator().manual_seed(12)
    for B, T in ((3, 253), (2, 19), (1, 10)):
        x = torch.randn(B, 257, T, generator=gen)
        ref = O.tsse(x, p, "channel_attention_real")
        m = g.modules.ChannelTimeSenseSELayer(257)
        m.load_state_dict({k[len("channel_attention_real."):]: v for k, v in p.items() if k.startswith("channel_attention_real.")})
        m.cuda()
        assert rel_err(m(cu(x)).cpu(), ref) < TOL


@pytest.mark.parametrize("C,d", [(257, 1), (514, 9), (257, 5)])
def test_tcn_block_vs_oracle(ops, C, d):
    import weights
    import generative_audio_b200 as g
    pre = "pretrained_restoration_model." if C == 257 else "audio_pc_wrapper.net."
    p = weights.synth_state_dict(5, 0, pre)
    idx = {1: 0, 2: 1, 5: 2, 9: 3}[d]
    blk = f"fb_model_imag.sequence_model.{idx}"
    gen = torch.Generator().manual_seed(13)
    x = torch.randn(3, C, 253, generator=gen)
    ref = O.tcn_block(x, p, blk, d)
    m = g.modules.TCNBlock(C, 512, C, dilation=d)
    m.load_state_dict({k[len(blk) + 1:]: v for k, v in p.items() if k.startswith(blk + ".")})
    m.cuda()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out = m(cu(x))
    finally:
        torch.backends.cudnn.allow_tf32 = True
    assert rel_err(out.cpu(), ref) < TOL
