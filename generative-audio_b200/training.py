"""Training step of the PC head (BASELINE config 3; reference: NPPCAudioTrainer.base_step + backward, nppc_audio/trainer.py:100-106,
234-317) with a HAND-WRITTEN backward.  What runs where:

  * frozen FullSubNet+ backbone, all STFTs, cIRM build, cRM decompress/apply, drop_band, the six input normalisers -> the
    inference kernels (no gradient flows there: the reference wraps the backbone in no_grad, nppc_model.py:94);
  * sub-band LSTM (2 layers + fc), forward AND backward -> the stepwise tcgen05 kernels of csrc/lstm_step.cu (SubbandLstmFn:
    fused pack -> LSTM forward with saved gates; BPTT; weight gradients as MN-major tcgen05 GEMMs; gradient of the packed
    input scattered back through drop_band / offline norm deterministically);
  * Gram-Schmidt + objective, forward -> nppc_gs_loss_fused; backward -> coefficient-space solve (gs_backward.py) + ONE
    streaming kernel (nppc_complex_lincomb), honouring both detach()s (pc_wrapper.py:37, trainer.py:295);
  * TCN stacks: channel-last; every 1x1 convolution and the output Linear, forward AND both backward GEMMs (dX = dY W,
    dW = dY^T X), on the in-house tcgen05 GEMMs (Conv1x1TCFn); PReLU / GroupNorm(1, C) / the dilated depthwise 3-tap
    convolution between them are elementwise torch ops (shifted adds — no F.conv1d, no cuDNN);
  * TSSE attention: windowed sums via cumsum + the tiny excitation MLP, elementwise torch ops.
  No nn.LSTM / cuDNN RNN / F.conv1d launches in the step.  fp16 gradient operands carry a power-of-two loss scale.
  * data parallelism: one process per GPU; gradients are all-reduced per bucket as soon as the bucket's last gradient has been
    accumulated (GradBucketReducer: NCCL all-reduce overlapped with the rest of the backward; mean of per-rank objectives).
"""
import math
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops
from .gs_backward import gs_coeffs_from_gram, gs_grad_coeffs, gs_loss_grad_coeffs
from .modules import TCN_DILATIONS, TCNBlock


def _pow2_scale(t: torch.Tensor, target_log2: int) -> torch.Tensor:
    """device scalar 2^k with max|t| * 2^k in [2^(target-1), 2^target) (1 when t == 0): no host synchronisation."""
    m = t.detach().abs().max().clamp_min(1e-30)
    return torch.exp2(torch.floor(target_log2 - torch.log2(m)) ).clamp(2.0 ** -60, 2.0 ** 60)


# ---- 1x1 convolution / Linear on the in-house tcgen05 GEMMs, forward + backward --------------------------------------
def _split16(t):
    """fp32 -> (hi, lo) fp16 halves with hi + lo = t to ~22 mantissa bits."""
    hi = t.half()
    return hi, (t - hi.float()).half()


class Conv1x1TCFn(torch.autograd.Function):
    """y [M, Np] = x [M, Kp] W^T (bias is added by the caller).  x: rows padded to a multiple of 128, columns to a multiple of
    128 (zeros); W [N, K] fp32 master weight.  All three GEMMs (Y = X W^T, dX = dY W, dW = dY^T X) run on the in-house tcgen05
    kernels with SPLIT-PRECISION fp16 operands (hi*hi + lo*hi + hi*lo, fp32 accumulate; power-of-two range scales): the
    residual stream of the real / imag branches spans 5 decades and scalar gradients (PReLU slopes, the 3-tap attention mix)
    are heavily cancelling sums — plain fp16 operands measured 5e-2 on them against the reference's fp32 autograd."""

    @staticmethod
    def forward(ctx, x, w):
        M, Kp = x.shape
        N, K = w.shape
        Np = -(-N // 128) * 128
        sx = _pow2_scale(x, 14)
        xh, xl = _split16(x * sx)
        wp = torch.zeros(Np, Kp, device=x.device, dtype=torch.float32)
        wp[:N, :K] = w.detach().clamp(-65504, 65504)
        wh, wl = _split16(wp)
        y = ops.gemm_f16_tn_ex(torch.cat([xh, xl], dim=1), torch.cat([wh, wh, wl], dim=1), out_f32=True)
        y.mul_(1.0 / sx)
        ctx.save_for_backward(xh, xl, wh, wl, sx)
        ctx.shape = (N, K)
        return y

    @staticmethod
    def backward(ctx, dy):
        xh, xl, wh, wl, sx = ctx.saved_tensors
        N, K = ctx.shape
        sy = _pow2_scale(dy, 10)
        dh, dl = _split16(dy * sy)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            wth, wtl = wh.t().contiguous(), wl.t().contiguous()
            dx = ops.gemm_f16_tn_ex(torch.cat([dh, dl], dim=1), torch.cat([wth, wth, wtl], dim=1), out_f32=True)   # dY W
            dx.mul_(1.0 / sy)
        if ctx.needs_input_grad[1]:
            dw = ops.gemm_f16_atb(torch.cat([dh, dl, dh], dim=0), torch.cat([xh, xh, xl], dim=0))[:N, :K] * (1.0 / (sy * sx))   # dY^T X
        return dx, dw


def conv1x1_tc(x, w):
    return Conv1x1TCFn.apply(x, w)


# ---- TSSE attention without convolutions -----------------------------------------------------------------------------
def _tsse_gate(m, x):
    """ChannelTimeSenseSELayer (attention_model.py:78-98) gate [B, C] for x [B, C, T] (x carries no gradient): the depthwise
    valid convolution followed by the time average is a weighted sum of windowed sums of x."""
    B, C, T = x.shape
    cs = F.pad(torch.cumsum(x.double(), dim=-1), (1, 0))                              # cs[t] = sum_{u < t} x[u]
    feats = []
    for conv in (m.smallConv1d[0], m.middleConv1d[0], m.largeConv1d[0]):
        k = conv.kernel_size[0]
        L = T - k + 1
        S = torch.stack([(cs[..., j + L] - cs[..., j]) for j in range(k)], dim=-1).float() / L    # [B, C, k] window means
        feats.append(torch.relu((S * conv.weight[:, 0, :][None]).sum(-1) + conv.bias[None]))
    s = m.feature_concate_fc(torch.stack(feats, dim=-1))[..., 0]
    return torch.sigmoid(m.fc2(torch.relu(m.fc1(s))))


# ---- TCN stack, channel-last, 1x1 convolutions on the tcgen05 GEMMs -----------------------------------------------------
def _groupnorm1_cl(y, B, T, gamma, beta, eps=1e-8):
    """GroupNorm(1, C) on channel-last rows y [B*T, C] (per-sample moments over all (T, C), biased variance): the
    normalisation itself is layout-agnostic (one fused native kernel each way on the [B, 1, T*C] view), the per-channel
    affine follows on the channel-last rows."""
    v = F.group_norm(y.reshape(B, 1, -1), 1, None, None, eps)
    return torch.addcmul(beta[None, :], v.reshape(y.shape), gamma[None, :])


def _dwconv3_cl(y, B, T, weight, bias, d):
    """depthwise Conv1d(k=3, dilation d, zero padding d) along time on channel-last rows y [B*T, C] as three shifted adds."""
    C = y.shape[1]
    v = y.reshape(B, T, C)
    p = F.pad(v, (0, 0, d, d))
    k = weight[:, 0, :]                                                             # [C, 3]
    out = p[:, 0:T] * k[None, None, :, 0] + v * k[None, None, :, 1] + p[:, 2 * d:2 * d + T] * k[None, None, :, 2] + bias[None, None, :]
    return out.reshape(B * T, C)


def _prelu(v, a):
    return F.prelu(v, a)


def _pad_rows(v, Mp):
    return v if v.shape[0] == Mp else F.pad(v, (0, 0, 0, Mp - v.shape[0]))


def tcn_forward_train(seq_model, x):
    """SequenceModel TCN branch (sequence_model.py:47-58,106-112; causal_conv.py:96-108) for x [B, C, T] -> [B, O, T]."""
    if seq_model.output_activate_function not in ("ReLU",):
        raise NotImplementedError("training path: the full-band models use output_activate_function='ReLU' (every reference config)")
    B, C, T = x.shape
    M = B * T
    Mp, Cp = -(-M // 128) * 128, -(-C // 128) * 128
    xr = F.pad(x.permute(0, 2, 1).reshape(M, C), (0, Cp - C))                       # residual stream [M, Cp], zero pad columns
    blocks = [m for m in seq_model.sequence_model if isinstance(m, TCNBlock)]
    for blk in blocks:
        y = conv1x1_tc(_pad_rows(xr, Mp), blk.conv1x1.weight[:, :, 0])[:M, :512] + blk.conv1x1.bias[None, :]
        y = _groupnorm1_cl(_prelu(y, blk.prelu1.weight), B, T, blk.norm1.weight, blk.norm1.bias)
        y = _dwconv3_cl(y, B, T, blk.depthwise_conv.weight, blk.depthwise_conv.bias, blk.dilation)
        y = _groupnorm1_cl(_prelu(y, blk.prelu2.weight), B, T, blk.norm2.weight, blk.norm2.bias)
        o = conv1x1_tc(_pad_rows(y, Mp), blk.sconv.weight[:, :, 0])[:M, :C] + blk.sconv.bias[None, :]
        xr = xr + F.pad(o, (0, Cp - C))
    O = seq_model.fc_output_layer.weight.shape[0]
    o = conv1x1_tc(_pad_rows(torch.relu(xr), Mp), seq_model.fc_output_layer.weight)[:M, :O] + seq_model.fc_output_layer.bias[None, :]
    return torch.relu(o).reshape(B, T, O).permute(0, 2, 1).contiguous()


# ---- sub-band pack + stepwise LSTM, forward + backward ------------------------------------------------------------------
def _dropband_maps(B, Fq, G, device):
    """row -> (sample, frequency) of the packed LSTM input after drop_band (feature.py:254-285): output batches are ordered by
    group g = sample % G; row = ob * F' + j holds frequency g + G j of sample sb(ob)."""
    if G <= 1 or B <= 1:
        sb = torch.arange(B, device=device).repeat_interleave(Fq)
        f = torch.arange(Fq, device=device).repeat(B)
        return sb, f, Fq
    Fg = Fq // G
    order = torch.cat([torch.arange(g, B, G, device=device) for g in range(G)])   # sample of output batch ob
    sb = order.repeat_interleave(Fg)
    f = (order % G).repeat_interleave(Fg) + G * torch.arange(Fg, device=device).repeat(B)
    return sb, f, Fg


class SubbandLstmFn(torch.autograd.Function):
    """(fb, fbr, fbi [B, F, T'] with gradient; nbr_src [B, F, T'] constant; ten LSTM / fc parameters) -> y [R, O, T'].
    forward : fused unfold ++ cat ++ offline_laplace_norm ++ drop_band -> fp16 time-major input (nppc_subband_pack) ->
              stepwise tcgen05 LSTM with saved gates (nppc_lstm_step_forward, train = 1)
    backward: BPTT + weight-gradient GEMMs (nppc_lstm_step_backward) -> gradient of the packed input -> back through
              drop_band (a permutation: unique indices, no atomics) and the per-sample mean of offline_laplace_norm."""

    @staticmethod
    def forward(ctx, fb, fbr, fbi, nbr_src, nn_, groups, *params):
        B, Fq, Tp = fb.shape
        xs, R, sums = ops.subband_pack(nbr_src, fb, fbr, fbi, nn_, groups, 64, torch.float16, want_sums=True)
        y, ws = ops.lstm_step_forward(params, xs, R, train=True)
        ctx.save_for_backward(xs, ws, sums, *params)
        ctx.meta = (B, Fq, Tp, nn_, groups, R)
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, ws, sums, *params = ctx.saved_tensors
        B, Fq, Tp, nn_, groups, R = ctx.meta
        need_x = any(ctx.needs_input_grad[:3])
        grads, dxs = ops.lstm_step_backward(params, xs, R, ws, dy.contiguous(), want_dxs=need_x)
        dfb = dfbr = dfbi = None
        if need_x:
            S = 2 * nn_ + 4
            G = groups if (groups > 1 and B > 1) else 1
            sb, f, Fg = _dropband_maps(B, Fq, G, dy.device)
            N = float(Fq * S * Tp)
            den = (sums / N).float() + 1e-5                                            # [B], the forward's divisor
            g = dxs[:, :R, :S]                                                          # [T', R, S]
            # sum_i g_i y_i per sample (y = the normalised input itself): rows of one output batch are contiguous
            D_row = (g * xs[:, :R, :S].float()).sum(dim=(0, 2))                         # [R]
            D_ob = D_row.reshape(-1, Fg).sum(dim=1)                                     # [B'] in output-batch order
            D = torch.zeros(B, device=dy.device, dtype=torch.float32)
            D[sb.reshape(-1, Fg)[:, 0]] = D_ob                                          # a permutation of the samples
            base = (-D / (N * den))[:, None, None]                                      # mean term: every (f, t) of the sample
            outs = []
            for k in range(3):
                d = base.expand(B, Fq, Tp).clone()
                d[sb, f] = d[sb, f] + g[:, :, S - 3 + k].t() / den[sb][:, None]          # kept (sample, frequency) pairs are unique
                outs.append(d)
            dfb, dfbr, dfbi = outs
        return (dfb, dfbr, dfbi, None, None, None, *grads)


# ---- Gram-Schmidt + objective -------------------------------------------------------------------------------------------
class GsLossFn(torch.autograd.Function):
    """head [B, n, 2, F', T] -> objective (trainer.py:337-342) with w_mat and the statistics as non-differentiable outputs."""

    @staticmethod
    def forward(ctx, head, gt, pred, lam):
        w, st, G, A = ops.gs_loss_fused_with_gram(head, gt, pred)
        lam = torch.as_tensor(lam, dtype=torch.float64, device=head.device)      # device scalar: a captured step follows the schedule
        objective = st["reconst_err"].mean() + (lam * st["second_moment_mse"].double().mean()).float()
        ctx.save_for_backward(head, gt, pred, G, A, lam)
        outs = (w, st["err_norm"], torch.view_as_real(st["err_proj"]), st["w_norms"], st["reconst_err"], st["second_moment_mse"])
        ctx.mark_non_differentiable(*outs)
        return (objective, *outs)

    @staticmethod
    def backward(ctx, g_obj, *unused):
        head, gt, pred, G, A, lam = ctx.saved_tensors
        coef = gs_loss_grad_coeffs(G, A.to(torch.complex128), lam) * g_obj.double()
        return ops.complex_lincomb(head, gt, pred, coef), None, None, None


class GramSchmidtFn(torch.autograd.Function):
    """gram_schmidt_to_crm (pc_wrapper.py:8-44) with a backward for an ARBITRARY upstream gradient: the reference trainer's own
    pattern — `w_mat = nppc_model(noisy)`, a loss written in torch, `.backward()` (trainer.py:255-298) — when the fused
    GsLossFn is not what the caller uses.  head [B, n, 2, F', T] -> w_mat, n <= 6.
    Backward: one Gram pass over the 2n vectors (x_0 .. x_{n-1}, g_0 .. g_{n-1}), the coefficient recursion of
    gs_backward.gs_grad_coeffs on 2n-vectors, one streaming linear combination — all on the Gram-Schmidt kernels."""

    @staticmethod
    def forward(ctx, head):
        n = head.shape[1]
        if n > 6:
            raise NotImplementedError("differentiable Gram-Schmidt: n_dirs <= 6 (the backward stacks 2 n vectors; kernels take 12)")
        ctx.save_for_backward(head)
        return ops.gram_schmidt_complex(head)

    @staticmethod
    def backward(ctx, g):
        (head,) = ctx.saved_tensors
        B, n = head.shape[:2]
        S = torch.cat([head, g.to(head.dtype)], dim=1).contiguous()           # [B, 2n, 2, F', T]
        G2 = ops.gram_matrix_complex(S)                                        # fp64 Gram pass over (x; g); its top-left block is G_x
        coef = gs_grad_coeffs(G2, gs_coeffs_from_gram(G2, n))                  # w = A x replayed in fp64 from G_x; C [B, n, 2n]
        full = torch.zeros(B, 2 * n, 2 * n + 1, dtype=coef.dtype, device=coef.device)   # rows n..2n-1 and the error column stay 0
        full[:, :n, :2 * n] = coef
        zero = torch.zeros(B, *head.shape[2:], device=head.device, dtype=head.dtype)
        return ops.complex_lincomb(S, zero, zero, full)[:, :n].contiguous()


# ---- the PC head, training forward ----------------------------------------------------------------------------------------
def head_forward_train(net, nmag, nreal, nimag, emag, ereal, eimag):
    """MultiDirectionFullSubNet_Plus.forward (networks.py:63-163) with an autograd graph made of the Functions above ->
    head [B', n, 2, F', T].  The six inputs [B, 1, F, T] carry no gradient (frozen backbone)."""
    if net.norm_type != "offline_laplace_norm":
        raise NotImplementedError("training path: norm_type='offline_laplace_norm' only (the head's shipped configuration); "
                                  f"got {net.norm_type!r}")
    la = net.look_ahead
    B, _, Fq, T = nmag.shape
    with torch.no_grad():
        xn = [ops.pad_offline_laplace_norm(v, la) for v in (nmag, nreal, nimag, emag, ereal, eimag)]   # [B, F, T'] each
        raw = F.pad(nmag[:, 0], [0, la]).contiguous()

    def stream(a, b, att, model):
        xa = a * _tsse_gate(att, a)[:, :, None]
        xb = b * _tsse_gate(att, b)[:, :, None]
        return tcn_forward_train(model, torch.cat([xa, xb], dim=1))

    fb = stream(xn[0], xn[3], net.channel_attention, net.fb_model)
    fbr = stream(xn[1], xn[4], net.channel_attention_real, net.fb_model_real)
    fbi = stream(xn[2], xn[5], net.channel_attention_imag, net.fb_model_imag)
    y = SubbandLstmFn.apply(fb, fbr, fbi, raw, net.sb_num_neighbors, net.num_groups_in_drop_band, *net.sb_model.lstm_params())
    R, O, Tp = y.shape
    Fp = R // B
    n = net.n_directions
    return y.reshape(B, Fp, n, 2, Tp).permute(0, 2, 3, 1, 4)[..., la:].contiguous()


# ---- data-parallel gradient exchange ------------------------------------------------------------------------------------
class GradBucketReducer:
    """DDP-style overlap without DDP: parameters are grouped into buckets in REVERSE registration order (the order their
    gradients become final during backward: loss -> LSTM -> TCN -> attention); a post-accumulate hook counts a bucket down and
    launches its all-reduce (async, NCCL stream) the moment its last gradient is in place, so the exchange of the LSTM
    gradients rides under the TCN backward.  finish() waits, averages and scatters back.  'Mean of per-rank objectives'
    semantics (upstream DDP averages the same way, nppc/auxil.py:297-302)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 16 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._pending: List[int] = []
        self._work = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.reset()

    def reset(self):
        self._pending = [len(b) for b in self.buckets]
        self._work = []

    def _active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _on_grad(self, p):
        i = self._bucket_of[id(p)]
        self._pending[i] -= 1
        if self._pending[i] == 0 and self._active():
            self._launch(i, self.buckets[i])

    def _launch(self, i, members):
        flat = torch.cat([q.grad.reshape(-1) for q in members])
        self._work.append((i, members, flat, dist.all_reduce(flat, async_op=True)))

    def finish(self) -> int:
        """Wait for the in-flight buckets (and exchange any whose hook never fired), write the averaged gradients back."""
        if not self._active():
            self.reset()
            return 0
        world = dist.get_world_size()
        done = {i for i, _, _, _ in self._work}
        for i, b in enumerate(self.buckets):
            if i not in done:
                # a bucket whose count-down never finished holds parameters this step's graph did not reach: exchange the ones
                # that do have a gradient (every rank runs the same graph, so every rank picks the same members) instead of
                # silently leaving the whole bucket un-averaged
                members = [q for q in b if q.grad is not None]
                if members:
                    self._launch(i, members)
        n = len(self._work)
        for i, members, flat, work in self._work:
            work.wait()
            flat.div_(world)
            off = 0
            for q in members:
                q.grad.copy_(flat[off:off + q.numel()].view_as(q.grad))
                off += q.numel()
        self.reset()
        return n

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20) -> int:
    """Blocking variant (after backward): mean-all-reduce in flat buckets.  Returns the number of collectives."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    calls, bucket, size = 0, [], 0

    def flush():
        nonlocal calls, bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        calls += 1
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return calls
