"""Training step of the PC head (BASELINE config 3; reference: NPPCAudioTrainer.base_step, nppc_audio/trainer.py:234-317).

What runs where in round 1:
  * frozen FullSubNet+ backbone, all STFTs, cIRM build, cRM decompress/apply, drop_band  -> hand-written kernels (no grads
    needed there: the reference wraps the backbone in no_grad, nppc_model.py:94);
  * PC head forward + backward (TSSE, TCN, sub-band LSTM, Gram-Schmidt, projection loss)  -> torch autograd on the SAME
    nn.Module parameters (cuDNN LSTM BPTT etc.).  The hand-written backward of the head is round-2 work; this keeps
    `base_step(...)[1].backward()` + optimizer usable as a drop-in today.  The frozen backbone and the noisy STFT run
    ONCE per step (the reference runs the backbone twice and the noisy STFT three times).
  * data parallelism: one process per GPU, flat-bucket NCCL all-reduce (mean) of the head gradients after backward.
"""
from typing import Iterable, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops
from .modules import TCN_DILATIONS


# ---- differentiable torch mirror of the PC head (same parameters as the kernels use) ---------------------------------
def _offline_norm(x):
    mu = x.reshape(x.shape[0], -1).mean(dim=1).reshape(-1, *([1] * (x.dim() - 1)))
    return x / (mu + 1e-5)


def _tsse(m, x):
    feats = [torch.relu(c[0](x).mean(dim=-1)) for c in (m.smallConv1d, m.middleConv1d, m.largeConv1d)]
    s = m.feature_concate_fc(torch.stack(feats, dim=-1))[..., 0]
    g = torch.sigmoid(m.fc2(torch.relu(m.fc1(s))))
    return x * g[:, :, None]


def _tcn(seq_model, x):
    for blk in list(seq_model.sequence_model)[:len(TCN_DILATIONS)]:
        y = blk.norm1(blk.prelu1(blk.conv1x1(x)))
        y = blk.norm2(blk.prelu2(blk.depthwise_conv(y)))
        x = x + blk.sconv(y)
    o = seq_model.fc_output_layer(torch.relu(x).permute(0, 2, 1))
    return torch.relu(o).permute(0, 2, 1)


def _unfold(x, n):
    """[B,F,T] -> [B,F,2n+1,T] with reflect padding along F (base_model.py:33-46)."""
    Fq = x.shape[1]
    idx = torch.arange(Fq, device=x.device)[:, None] + torch.arange(2 * n + 1, device=x.device)[None, :] - n
    idx = torch.where(idx < 0, -idx, idx)
    idx = torch.where(idx > Fq - 1, 2 * (Fq - 1) - idx, idx)
    return x[:, idx, :]


def _drop_band(x, G):
    B, _, Fq, _ = x.shape
    assert B > G, f"Batch size = {B}, num_groups = {G}. The batch size should larger than the num_groups."
    if G <= 1:
        return x
    Fq -= Fq % G
    return torch.cat([x[g::G, :, g:Fq:G, :] for g in range(G)], dim=0)


def head_forward_autograd(net, nmag, nreal, nimag, emag, ereal, eimag, amp_dtype=None):
    """MultiDirectionFullSubNet_Plus.forward (networks.py:63-163) in differentiable torch ops -> [B, n, 2, F', T].
    amp_dtype (e.g. torch.bfloat16, BASELINE config 3): the GEMM-shaped parts (TCN convolutions, LSTM, fc) run under
    torch.autocast in that dtype; normalisers, attention, Gram-Schmidt and the loss stay fp32."""
    amp = lambda: torch.autocast("cuda", dtype=amp_dtype, enabled=amp_dtype is not None)
    la = net.look_ahead
    nmag, nreal, nimag, emag, ereal, eimag = (F.pad(v, [0, la]) for v in (nmag, nreal, nimag, emag, ereal, eimag))
    B, _, Fq, Tp = nmag.shape

    def stream(noisy, enh, att, model):
        a = _tsse(att, _offline_norm(noisy).reshape(B, Fq, Tp))
        b = _tsse(att, _offline_norm(enh).reshape(B, Fq, Tp))
        with amp():
            return _tcn(model, torch.cat([a, b], dim=1)).float()

    fb = stream(nmag, emag, net.channel_attention, net.fb_model)
    fbr = stream(nreal, ereal, net.channel_attention_real, net.fb_model_real)
    fbi = stream(nimag, eimag, net.channel_attention_imag, net.fb_model_imag)
    sb = torch.cat([_unfold(nmag[:, 0], net.sb_num_neighbors), fb[:, :, None], fbr[:, :, None], fbi[:, :, None]], dim=2)
    sb = _offline_norm(sb)
    if B > 1:
        sb = _drop_band(sb.permute(0, 2, 1, 3), net.num_groups_in_drop_band).permute(0, 2, 1, 3)
    Fp, S = sb.shape[1], sb.shape[2]
    seq = sb.reshape(B * Fp, S, Tp).permute(0, 2, 1).contiguous()
    lstm = net.sb_model.sequence_model
    was_training = lstm.training
    lstm.train(True)   # cuDNN's RNN backward needs the training-mode forward (no dropout here: identical numerics)
    try:
        with amp():
            o, _ = lstm(seq)
            y = net.sb_model.fc_output_layer(o).float()
    finally:
        lstm.train(was_training)
    y = y.permute(0, 2, 1)  # [B*F', 2n, T']
    n = net.n_directions
    return y.reshape(B, Fp, n, 2, Tp).permute(0, 2, 3, 1, 4)[..., la:]


def gram_schmidt_autograd(x):
    """pc_wrapper.py:8-44 incl. the conjugated coefficient and the detached normaliser."""
    B, n, _, Fq, T = x.shape
    v = torch.complex(x[:, :, 0], x[:, :, 1]).reshape(B, n, -1)
    outs, hats = [], []
    for i in range(n):
        w = v[:, i]
        for wh in hats:
            w = w - wh * (w.conj() * wh).sum(dim=1, keepdim=True)
        wd = w.detach()
        hats.append(wd / torch.linalg.vector_norm(wd, dim=1, keepdim=True))
        outs.append(w)
    out = torch.stack(outs, dim=1).reshape(B, n, Fq, T)
    return torch.stack([out.real, out.imag], dim=2)


def nppc_loss_autograd(w_mat, gt, pred, lam):
    """trainer.py:259-298, 337-342."""
    B, n = w_mat.shape[:2]
    W = w_mat.reshape(B, n, 2, -1)
    w_norms = torch.linalg.vector_norm(W, dim=(2, 3))
    w_hat = W / (w_norms[..., None, None] + 1e-8)
    err = (gt - pred).reshape(B, 2, -1)
    err_norm = torch.linalg.vector_norm(err, dim=(1, 2))
    err = err / (err_norm[:, None, None] + 1e-8)
    w_norms = w_norms / (err_norm[:, None] + 1e-8)
    err_proj = (torch.complex(w_hat[:, :, 0], w_hat[:, :, 1]).conj() * torch.complex(err[:, 0], err[:, 1])[:, None]).sum(-1)
    mag = err_proj.abs()
    reconst_err = 1 - mag.pow(2).sum(dim=1)
    second_moment_mse = (w_norms.pow(2) - mag.detach().pow(2)).pow(2)
    objective = reconst_err.mean() + lam * second_moment_mse.mean()
    return dict(err_norm=err_norm, err_proj=err_proj, err_proj_mag=mag, w_norms=w_norms, reconst_err=reconst_err,
                second_moment_mse=second_moment_mse, objective=objective)


# ---- data-parallel gradient exchange ------------------------------------------------------------------------------------
def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20) -> int:
    """Mean-all-reduce the gradients of `params` over the default process group in flat buckets (NCCL over NVLink on GPUs,
    gloo in the CPU tests).  DP parity is 'mean of per-rank objectives' (SURVEY.md §8e).  Returns the number of collectives."""
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
