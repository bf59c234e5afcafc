"""GPU parity tests of the tcgen05 paths (bf16 GEMM, tensor-core LSTM, full model with lstm_impl="tc").
Tolerance for bf16-GEMM paths: rel 1e-2 (max-norm), SURVEY.md §8d / BASELINE.json north_star."""
import pytest
import torch

import nppc_oracle as O
import weights
from conftest import load_golden, rel_err
from helpers import build_model, wave

pytestmark = pytest.mark.gpu
TOL_BF16 = 1e-2


@pytest.fixture(scope="module")
def ops():
    import generative_audio_b200 as g
    return g.ops


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1000, 1536, 64), (300, 1536, 384), (20000, 256, 128), (77, 128, 64)])
def test_gemm_bf16_tn(ops, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.1).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    c = ops.gemm_bf16_tn(a.cuda(), w.cuda(), bias.cuda())
    ref = a.float() @ w.float().T + bias
    assert rel_err(c.float().cpu(), ref) < 6e-3  # bf16 output rounding (2^-9) only: accumulation is fp32
    c2 = ops.gemm_bf16_tn(a.cuda(), w.cuda(), None)
    assert rel_err(c2.float().cpu(), a.float() @ w.float().T) < 6e-3


def _plan(ops, p, pre="sb_model"):
    lp = pre + ".sequence_model."
    return ops.LstmPlan(*[p[lp + f"{k}_l{l}"].cuda() for l in (0, 1) for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")],
                        p[pre + ".fc_output_layer.weight"].cuda(), p[pre + ".fc_output_layer.bias"].cuda())


@pytest.mark.parametrize("R,Tp,which", [(50, 23, "backbone"), (300, 40, "head"), (257, 253, "backbone"), (200, 30, "head10"),
                                          (1, 1, "head"), (130, 3, "backbone")])
def test_lstm_tc_vs_oracle(ops, R, Tp, which):
    """O = 2 (backbone) and O = 10 (head, n_dirs = 5): fc fused into the recurrent kernel; O = 20 (n_dirs = 10): separate fc kernel."""
    pre = "pretrained_restoration_model." if which == "backbone" else "audio_pc_wrapper.net."
    p = weights.synth_state_dict(10 if which == "head10" else 5, 0, pre)
    g = torch.Generator().manual_seed(R)
    x = torch.randn(R, 34, Tp, generator=g)
    ref = O.lstm_fc(x, p, "sb_model", fast=True)
    plan = _plan(ops, p)
    xs = torch.zeros(Tp, R, 64)
    xs[:, :, :34] = x.permute(2, 0, 1)
    RS = ops.padded_rows(R, torch.float16)
    xsb = torch.zeros(Tp, RS, 64, dtype=torch.float16)
    xsb[:, :R] = xs.to(torch.float16)
    y = plan.forward(xsb.cuda(), 1, R)
    y0 = plan.forward(xs.cuda(), 0)
    assert y.shape == ref.shape
    e_tc, e_f32 = rel_err(y.cpu(), ref), rel_err(y0.cpu(), ref)
    print(f"lstm tc rel_err={e_tc:.3e}  f32 rel_err={e_f32:.3e}")
    assert e_f32 < 1e-4
    assert e_tc < TOL_BF16


def test_model_small_b2_tc():
    g = load_golden("model_small_b2")
    m, sd = build_model(5, 1, "tc")
    head, crm = m.forward_stages(g["wave"].cuda())
    e_crm = rel_err(crm.cpu(), g["pred_crm"])
    e_head = rel_err(head.reshape(2, 10, 257, 17).cpu(), g["head"])
    w = m(g["wave"].cuda())
    e_w = rel_err(w.cpu(), g["w_mat"])
    print(f"tc model small: pred_crm {e_crm:.3e} head {e_head:.3e} w_mat {e_w:.3e}")
    assert e_crm < TOL_BF16 and e_head < TOL_BF16 and e_w < TOL_BF16


def test_model_full_b1_tc():
    g = load_golden("model_full_b1")
    m, sd = build_model(5, 1, "tc")
    w = m(g["wave"].cuda())
    e_crm = rel_err(m.get_pred_crm(g["wave"].cuda()).cpu(), g["pred_crm"])
    e_w = rel_err(w.cpu(), g["w_mat"])
    e_enh = rel_err(m.enhance(g["wave"].cuda()).cpu(), g["enhanced_wave"])
    print(f"tc model full: pred_crm {e_crm:.3e} w_mat {e_w:.3e} enhanced {e_enh:.3e}")
    assert e_crm < TOL_BF16 and e_w < TOL_BF16 and e_enh < TOL_BF16


def test_model_tc_matches_f32_at_bench_like_batch():
    """size-independent property at a larger batch: the tensor-core model agrees with the fp32-kernel model."""
    x = wave(8, 64000, 77).cuda()
    mt, _ = build_model(5, 1, "tc")
    mf, _ = build_model(5, 1, "f32")
    wt, wf = mt(x), mf(x)
    e = rel_err(wt.cpu(), wf.cpu())
    print(f"tc vs f32 (B=8 x 4 s): {e:.3e}")
    assert e < TOL_BF16


@pytest.mark.parametrize("C", [257, 514])
def test_tcn_stack_tcgen05_vs_fp32_path(C):
    """Row N2: channel-last TCN stack with its 1x1 convolutions on the tcgen05 fp16 GEMM vs the fp32 channel-first path
    (library convolutions + tcn.cu kernels) and vs the CPU oracle."""
    import generative_audio_b200 as g
    import nppc_oracle as O
    import weights
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = g.modules.SequenceModel(input_size=C, output_size=257, hidden_size=512, num_layers=2, bidirectional=False,
                                sequence_model="TCN", output_activate_function="ReLU")
    sd = {k: weights.synth_tensor("pretrained_restoration_model.fb_model." + k, tuple(v.shape), 3) for k, v in m.state_dict().items()}
    m.load_state_dict(sd)
    m = m.cuda().eval()
    gen = torch.Generator().manual_seed(11)
    x = torch.rand(3, C, 61, generator=gen) * 2.0
    x[1] = (x[1] - 1.0) * 3.0e4   # the real / imag streams: signed and far outside the fp16 range
    with torch.no_grad():
        ref = m(x.cuda())
        m.use_tc_convs = True
        out = m(x.cuda())
        cpu = O.tcn_sequence_model(x, {"fb." + k: v for k, v in sd.items()}, "fb")
    e_ref, e_tc = rel_err(ref.cpu(), cpu), rel_err(out.cpu(), cpu)
    print(f"tcn C={C}: fp32 path vs oracle {e_ref:.3e}, tcgen05 path vs oracle {e_tc:.3e}")
    assert e_ref < 1e-4
    assert e_tc < 5e-3


def test_lstm_tc_bitwise_deterministic(ops):
    """The recurrent kernel hands h_t between warps, CTAs and proxies (TMEM, shared memory written by the epilogue and read
    by the tensor core, TMA store -> TMA reload): any race would show up as run-to-run differences.  Three runs, many CTA
    pairs, must agree bit for bit."""
    p = weights.synth_state_dict(5, 0, "audio_pc_wrapper.net.")
    plan = _plan(ops, p)
    R, Tp = 3000, 48
    g = torch.Generator().manual_seed(1)
    RS = ops.padded_rows(R, torch.float16)
    xs = torch.zeros(Tp, RS, 64, dtype=torch.float16)
    xs[:, :R, :34] = torch.randn(Tp, R, 34, generator=g).half()
    xs = xs.cuda()
    y0 = plan.forward(xs, 1, R).clone()
    for _ in range(2):
        assert torch.equal(plan.forward(xs, 1, R), y0)
