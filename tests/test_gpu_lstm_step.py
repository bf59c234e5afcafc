"""Stepwise tensor-core LSTM (csrc/lstm_step.cu): forward (fast and split-precision), BPTT and the MN-major weight-gradient
GEMM against plain torch (fp64 CPU for the forward, torch autograd for the gradients).  Shapes include the original
FullSubNet full-band LSTM 257 -> 512 (fullsubnet.py:39-47), which the persistent H = 384 kernel is not built for (row J1)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _params(I, H, O, seed, dev="cuda"):
    g = torch.Generator().manual_seed(seed)
    b = 1.0 / H ** 0.5
    u = lambda *s: (torch.rand(*s, generator=g) * 2 - 1) * b
    return [t.to(dev) for t in (u(4 * H, I), u(4 * H, H), u(4 * H), u(4 * H), u(4 * H, H), u(4 * H, H), u(4 * H), u(4 * H), u(O, H), u(O))]


def _ref_lstm(params, x):
    """x [R, T, I] -> y [R, O, T] with torch.lstm in the dtype / device of x."""
    ps = [p.to(x.device, x.dtype) for p in params]
    R = x.shape[0]
    H = ps[1].shape[1]
    z = torch.zeros(2, R, H, dtype=x.dtype, device=x.device)
    out = torch.lstm(x, (z, z), ps[:8], True, 2, 0.0, False, False, True)[0]
    return (out @ ps[8].T + ps[9]).transpose(1, 2)


def _pack(x, KP, dtype):
    """x [R, T, I] -> time-major [T, RS, KP]"""
    R, T, I = x.shape
    RS = -(-R // 128) * 128
    xs = torch.zeros(T, RS, KP, device="cuda", dtype=dtype)
    xs[:, :R, :I] = x.permute(1, 0, 2).to(dtype)
    return xs


@pytest.mark.parametrize("rows,mo,no,splits", [(256, 128, 64, 1), (4096, 1536, 384, 2), (1024, 384, 64, 4), (8192, 256, 128, None)])
def test_gemm_atb_matches_torch(rows, mo, no, splits):
    import generative_audio_b200 as G
    g = torch.Generator().manual_seed(rows + mo)
    a = torch.randn(rows, mo, generator=g).cuda().half()
    b = torch.randn(rows, no, generator=g).cuda().half()
    c = G.ops.gemm_f16_atb(a, b, splits)
    ref = a.double().T @ b.double()
    assert rel_err(c.cpu(), ref.cpu()) < 2e-3
    assert torch.equal(c, G.ops.gemm_f16_atb(a, b, splits))   # fixed summation order


@pytest.mark.parametrize("I,H,O,R,T", [(34, 384, 10, 300, 9), (257, 512, 4, 130, 7), (34, 384, 2, 128, 33)])
def test_lstm_step_forward_fast_and_precise(I, H, O, R, T):
    import generative_audio_b200 as G
    params = _params(I, H, O, 3)
    x = torch.randn(R, T, I, generator=torch.Generator().manual_seed(5))
    ref64 = _ref_lstm([p.cpu() for p in params], x.double())
    ref32 = _ref_lstm([p.cpu() for p in params], x)
    KP = -(-I // 64) * 64
    y, _ = G.ops.lstm_step_forward(params, _pack(x.cuda(), KP, torch.float16), R)
    assert y.shape == (R, O, T)
    assert rel_err(y.cpu(), ref64) < 3e-3                      # fp16 operands, tanh.approx
    yp, _ = G.ops.lstm_step_forward(params, _pack(x.cuda(), KP, torch.float32), R, precise=True)
    e32 = rel_err(ref32, ref64)
    assert rel_err(yp.cpu(), ref64) < max(2e-5, 4 * e32)       # split-precision operands: fp32-class


@pytest.mark.parametrize("I,H,O,R,T", [(34, 384, 10, 260, 12), (40, 128, 3, 128, 5)])
def test_lstm_step_backward_matches_autograd(I, H, O, R, T):
    import generative_audio_b200 as G
    params = _params(I, H, O, 7)
    x = torch.randn(R, T, I, generator=torch.Generator().manual_seed(9)).cuda()
    dy = (torch.randn(R, O, T, generator=torch.Generator().manual_seed(11)) * 1e-4).cuda()
    KP = -(-I // 64) * 64
    xs = _pack(x, KP, torch.float16)
    y, ws = G.ops.lstm_step_forward(params, xs, R, train=True)
    grads, dxs = G.ops.lstm_step_backward(params, xs, R, ws, dy)
    # reference: fp64 autograd on the fp16-rounded input (what the kernels saw)
    with torch.enable_grad():                                              # (another test module may have switched autograd off)
        ps = [p.double().cpu().requires_grad_(True) for p in params]      # CPU: cuDNN's RNN backward needs training mode
        xr = xs[:, :R, :I].permute(1, 0, 2).double().cpu().requires_grad_(True)
        yr = _ref_lstm(ps, xr)
        yr.backward(dy.double().cpu())
    assert rel_err(y.cpu(), yr.detach().cpu()) < 3e-3
    names = "w_ih0 w_hh0 b_ih0 b_hh0 w_ih1 w_hh1 b_ih1 b_hh1 fc_w fc_b".split()
    for n, g, p in zip(names, grads, ps):
        assert rel_err(g.cpu(), p.grad.cpu()) < 2e-2, n
    gx = dxs[:, :R, :I].permute(1, 0, 2)
    assert rel_err(gx.cpu(), xr.grad.cpu()) < 2e-2
