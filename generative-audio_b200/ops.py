"""Tensor-level wrappers over the C ABI: torch supplies device memory and the current stream (plumbing);
every op below is one or more hand-written sm_100a kernels.  CUDA tensors only — CPU tensors raise."""
import torch

from . import _lib

_BF16 = torch.bfloat16
_F16 = torch.float16  # operand dtype of the tensor-core LSTM path


def _chk(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("generative-audio_b200 ops are CUDA-only (no CPU fallback); got a CPU tensor")
        if not t.is_contiguous():
            raise RuntimeError("generative-audio_b200 ops need contiguous tensors")


def _f32(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(_lib.load().nppc_launch_count())


def reset_launch_count():
    _lib.load().nppc_reset_launch_count()


def stft_mri(wave: torch.Tensor, n_fft: int = 512, hop: int = 256, win: int = 512):
    """utils.prepare_input_from_waveform (utils.py:107-147): wave [B,L] -> mag, real, imag each [B,1,F,T]."""
    if wave.dim() == 1:
        wave = wave[None]
    if win != n_fft:
        raise AssertionError("win_length must equal n_fft (all reference configs do)")
    wave = _f32(wave)
    _chk(wave)
    B, L = wave.shape
    F, T = n_fft // 2 + 1, 1 + L // hop
    out = torch.empty(3, B, 1, F, T, device=wave.device, dtype=torch.float32)
    _lib.check(_lib.load().nppc_stft_mri(wave.data_ptr(), B, L, n_fft, hop, out[0].data_ptr(), out[1].data_ptr(),
                                         out[2].data_ptr(), _stream()), "nppc_stft_mri")
    return out[0], out[1], out[2]


def istft(real: torch.Tensor, imag: torch.Tensor, length: int, n_fft: int = 512, hop: int = 256):
    """torch.istft(center=True, window=hann, length=length) (utils.py:60-70): [B,F,T] x2 -> [B,length]."""
    real, imag = _f32(real), _f32(imag)
    _chk(real, imag)
    B, F, T = real.shape
    out = torch.empty(B, length, device=real.device, dtype=torch.float32)
    _lib.check(_lib.load().nppc_istft(real.data_ptr(), imag.data_ptr(), B, T, n_fft, hop, length, out.data_ptr(),
                                      _stream()), "nppc_istft")
    return out


def crm_decompress_apply(crm: torch.Tensor, real: torch.Tensor, imag: torch.Tensor, conj: bool, want_mag=True):
    """decompress_cIRM + mask apply. crm [B,2,F,T] compressed; real/imag [B,(1,)F,T] -> (mag, real, imag) [B,F,T]."""
    crm, real, imag = _f32(crm), _f32(real), _f32(imag)
    _chk(crm, real, imag)
    B, two, F, T = crm.shape
    assert two == 2
    out = torch.empty(3, B, F, T, device=crm.device, dtype=torch.float32)
    _lib.check(_lib.load().nppc_crm_decompress_apply(crm.data_ptr(), real.data_ptr(), imag.data_ptr(), B, F * T,
                                                     1 if conj else 0, out[0].data_ptr() if want_mag else 0,
                                                     out[1].data_ptr(), out[2].data_ptr(), _stream()),
               "nppc_crm_decompress_apply")
    return out[0], out[1], out[2]


def decompress_cirm(m: torch.Tensor):
    m = _f32(m)
    _chk(m)
    out = torch.empty_like(m)
    _lib.check(_lib.load().nppc_decompress_cirm(m.data_ptr(), m.numel(), out.data_ptr(), _stream()), "nppc_decompress_cirm")
    return out


def build_cirm(nr, ni, cr, ci):
    """build_complex_ideal_ratio_mask (mask.py:24-41): [B,F,T] x4 -> compressed gt [B,2,F,T]."""
    nr, ni, cr, ci = (_f32(v) for v in (nr, ni, cr, ci))
    _chk(nr, ni, cr, ci)
    B, F, T = nr.shape
    gt = torch.empty(B, 2, F, T, device=nr.device, dtype=torch.float32)
    _lib.check(_lib.load().nppc_build_cirm(nr.data_ptr(), ni.data_ptr(), cr.data_ptr(), ci.data_ptr(), B, F * T,
                                           gt.data_ptr(), _stream()), "nppc_build_cirm")
    return gt


def offline_laplace_norm(x: torch.Tensor):
    """BaseModel.offline_laplace_norm (base_model.py:210-224) for x [B, ...]."""
    x = _f32(x)
    _chk(x)
    B = x.shape[0]
    y = torch.empty_like(x)
    sums = torch.empty(B, device=x.device, dtype=torch.float64)
    _lib.check(_lib.load().nppc_offline_laplace_norm(x.data_ptr(), B, x.numel() // B, sums.data_ptr(), y.data_ptr(),
                                                     _stream()), "nppc_offline_laplace_norm")
    return y


def pad_offline_laplace_norm(x: torch.Tensor, look_ahead: int):
    """F.pad(x,[0,la]) + offline_laplace_norm fused: x [B,1,F,T] or [B,F,T] -> [B,F,T+la]."""
    x = _f32(x)
    _chk(x)
    if x.dim() == 4:
        x = x[:, 0]
    B, F, T = x.shape
    y = torch.empty(B, F, T + look_ahead, device=x.device, dtype=torch.float32)
    sums = torch.empty(B, device=x.device, dtype=torch.float64)
    _lib.check(_lib.load().nppc_pad_offline_laplace_norm(x.data_ptr(), B, F, T, look_ahead, sums.data_ptr(),
                                                         y.data_ptr(), _stream()), "nppc_pad_offline_laplace_norm")
    return y


def cancel_depth(x: torch.Tensor, count: float, depth: torch.Tensor = None):
    """depth[b] = max(depth[b], cancellation depth of the offline normaliser's mean over x[b]) (see nppc_cancel_depth)."""
    x = _f32(x)
    _chk(x, depth)
    B = x.shape[0]
    if depth is None:
        depth = torch.zeros(B, device=x.device, dtype=torch.float32)
    sums = torch.empty(2 * B, device=x.device, dtype=torch.float64)
    _lib.check(_lib.load().nppc_cancel_depth(x.data_ptr(), B, x[0].numel(), float(count), sums.data_ptr(), depth.data_ptr(), _stream()),
               "nppc_cancel_depth")
    return depth


def cumulative_laplace_norm(x: torch.Tensor):
    """BaseModel.cumulative_laplace_norm (base_model.py:227-257): x [B,C,F,T]."""
    x = _f32(x)
    _chk(x)
    B, Cc, F, T = x.shape
    y = torch.empty_like(x)
    _lib.check(_lib.load().nppc_cumulative_laplace_norm(x.data_ptr(), B * Cc, F, T, y.data_ptr(), _stream()),
               "nppc_cumulative_laplace_norm")
    return y


def unfold(x: torch.Tensor, num_neighbor: int):
    """BaseModel.unfold (base_model.py:15-46): [B,C,F,T] -> [B,F,C,2n+1,T], bit-exact."""
    assert x.dim() == 4, f"The dim of input is {x.dim()}. It should be four dim."
    x = _f32(x)
    _chk(x)
    B, Cc, F, T = x.shape
    out = torch.empty(B, F, Cc, 2 * num_neighbor + 1, T, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().nppc_unfold(x.data_ptr(), B, Cc, F, T, num_neighbor, out.data_ptr(), _stream()), "nppc_unfold")
    return out


def drop_band(x: torch.Tensor, num_groups: int = 2):
    """drop_band (feature.py:254-285): [B,C,F,T] -> [B,C,F//G,T], bit-exact; asserts B > G."""
    x = _f32(x)
    _chk(x)
    B, Cc, F, T = x.shape
    G = max(num_groups, 1)
    out = torch.empty(B, Cc, F // G, T, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().nppc_drop_band(x.data_ptr(), B, Cc, F, T, num_groups, out.data_ptr(), _stream()), "nppc_drop_band")
    return out


def _gs_scratch(B, n, device):
    nbytes = _lib.load().nppc_gs_scratch_bytes(B, n)
    return torch.empty(nbytes, device=device, dtype=torch.uint8)


def gram_schmidt_complex(x: torch.Tensor):
    """gram_schmidt_to_crm (pc_wrapper.py:8-44): x [B,n,2,F,T] -> same."""
    x = _f32(x)
    _chk(x)
    B, n = x.shape[:2]
    assert x.shape[2] == 2
    P = x[0, 0, 0].numel()
    out = torch.empty_like(x)
    scr = _gs_scratch(B, n, x.device)
    _lib.check(_lib.load().nppc_gram_schmidt_complex(x.data_ptr(), B, n, P, scr.data_ptr(), out.data_ptr(), _stream()),
               "nppc_gram_schmidt_complex")
    return out


def gram_schmidt_real(x: torch.Tensor):
    """gram_schmidt_to_spec_mag (inpainting/nppc/pc_wrapper.py:43-59): x [B,n,...] -> same."""
    x = _f32(x)
    _chk(x)
    B, n = x.shape[:2]
    P = x[0, 0].numel()
    out = torch.empty_like(x)
    scr = _gs_scratch(B, n, x.device)
    _lib.check(_lib.load().nppc_gram_schmidt_real(x.data_ptr(), B, n, P, scr.data_ptr(), out.data_ptr(), _stream()),
               "nppc_gram_schmidt_real")
    return out


def _loss_outputs(B, n, device):
    f = dict(device=device, dtype=torch.float32)
    return (torch.empty(B, **f), torch.empty(B, n, 2, **f), torch.empty(B, n, **f), torch.empty(B, **f),
            torch.empty(B, n, **f))


def gs_loss_fused(head: torch.Tensor, gt: torch.Tensor, pred: torch.Tensor):
    """Gram-Schmidt + loss statistics in two HBM passes. head [B,n,2,F,T]; gt/pred [B,2,F,T]."""
    head, gt, pred = _f32(head), _f32(gt), _f32(pred)
    _chk(head, gt, pred)
    B, n = head.shape[:2]
    P = head[0, 0, 0].numel()
    assert gt.shape == pred.shape and gt[0, 0].numel() == P
    w = torch.empty_like(head)
    err_norm, err_proj, w_norms, reconst, sm = _loss_outputs(B, n, head.device)
    scr = _gs_scratch(B, n, head.device)
    _lib.check(_lib.load().nppc_gs_loss_fused(head.data_ptr(), gt.data_ptr(), pred.data_ptr(), B, n, P, scr.data_ptr(),
                                              w.data_ptr(), err_norm.data_ptr(), err_proj.data_ptr(), w_norms.data_ptr(),
                                              reconst.data_ptr(), sm.data_ptr(), _stream()), "nppc_gs_loss_fused")
    return w, dict(err_norm=err_norm, err_proj=torch.view_as_complex(err_proj), w_norms=w_norms, reconst_err=reconst,
                   second_moment_mse=sm)


GS_SCRATCH_BYTES = 13 * 13 * 2 * 8 + 12 * 12 * 2 * 4   # SampleScratch of gram_schmidt.cu


def gs_loss_fused_with_gram(head: torch.Tensor, gt: torch.Tensor, pred: torch.Tensor):
    """gs_loss_fused that also returns what the backward needs from the forward's scratch: the Hermitian Gram matrix
    G [B, n+1, n+1] complex128 of (x_0 .. x_{n-1}, gt - pred) and the coefficient matrix A [B, n, n] complex64 (w = A x)."""
    head, gt, pred = _f32(head), _f32(gt), _f32(pred)
    _chk(head, gt, pred)
    B, n = head.shape[:2]
    P = head[0, 0, 0].numel()
    w = torch.empty_like(head)
    err_norm, err_proj, w_norms, reconst, sm = _loss_outputs(B, n, head.device)
    scr = _gs_scratch(B, n, head.device)
    assert scr.numel() >= B * GS_SCRATCH_BYTES
    _lib.check(_lib.load().nppc_gs_loss_fused(head.data_ptr(), gt.data_ptr(), pred.data_ptr(), B, n, P, scr.data_ptr(),
                                              w.data_ptr(), err_norm.data_ptr(), err_proj.data_ptr(), w_norms.data_ptr(),
                                              reconst.data_ptr(), sm.data_ptr(), _stream()), "nppc_gs_loss_fused")
    scr = scr[:B * GS_SCRATCH_BYTES].reshape(B, GS_SCRATCH_BYTES)
    Gu = torch.view_as_complex(scr[:, :13 * 13 * 16].contiguous().view(torch.float64).reshape(B, 13, 13, 2))[:, :n + 1, :n + 1]
    up = torch.triu(Gu, diagonal=1)
    G = torch.diag_embed(torch.diagonal(Gu, dim1=1, dim2=2).real.to(Gu.dtype)) + up + up.conj().transpose(1, 2)
    A = torch.view_as_complex(scr[:, 13 * 13 * 16:].contiguous().view(torch.float32).reshape(B, 12, 12, 2))[:, :n, :n]
    st = dict(err_norm=err_norm, err_proj=torch.view_as_complex(err_proj), w_norms=w_norms, reconst_err=reconst, second_moment_mse=sm)
    return w, st, G, A


def complex_lincomb(x: torch.Tensor, gt: torch.Tensor, pred: torch.Tensor, coef: torch.Tensor):
    """out[b,i] = sum_k coef[b,i,k] x[b,k] + coef[b,i,n] (gt - pred)[b]; x [B,n,2,...], coef [B,n,n+1] complex."""
    x, gt, pred = _f32(x), _f32(gt), _f32(pred)
    c = torch.view_as_real(coef.to(torch.complex64).contiguous()).contiguous()
    _chk(x, gt, pred, c)
    B, n = x.shape[:2]
    P = x[0, 0, 0].numel()
    out = torch.empty_like(x)
    _lib.check(_lib.load().nppc_complex_lincomb(x.data_ptr(), gt.data_ptr(), pred.data_ptr(), B, n, P, c.data_ptr(), out.data_ptr(),
                                                _stream()), "nppc_complex_lincomb")
    return out


def _decode_gs_scratch(scr: torch.Tensor, B: int, n: int):
    """SampleScratch of gram_schmidt.cu -> (Hermitian Gram matrix [B,n,n] complex128 of the n input vectors, coefficient
    matrix A [B,n,n] complex64 with w = A x)."""
    scr = scr[:B * GS_SCRATCH_BYTES].reshape(B, GS_SCRATCH_BYTES)
    Gu = torch.view_as_complex(scr[:, :13 * 13 * 16].contiguous().view(torch.float64).reshape(B, 13, 13, 2))[:, :n, :n]
    up = torch.triu(Gu, diagonal=1)
    G = torch.diag_embed(torch.diagonal(Gu, dim1=1, dim2=2).real.to(Gu.dtype)) + up + up.conj().transpose(1, 2)
    A = torch.view_as_complex(scr[:, 13 * 13 * 16:].contiguous().view(torch.float32).reshape(B, 12, 12, 2))[:, :n, :n]
    return G, A


def gram_schmidt_complex_with_coeffs(x: torch.Tensor):
    """gram_schmidt_complex that also returns (G, A) from the kernel's scratch: the Gram matrix of the inputs and the
    coefficients of w = A x — what the backward of a differentiable Gram-Schmidt needs (training.GramSchmidtFn)."""
    x = _f32(x)
    _chk(x)
    B, n = x.shape[:2]
    assert x.shape[2] == 2
    P = x[0, 0, 0].numel()
    out = torch.empty_like(x)
    scr = _gs_scratch(B, n, x.device)
    assert scr.numel() >= B * GS_SCRATCH_BYTES
    _lib.check(_lib.load().nppc_gram_schmidt_complex(x.data_ptr(), B, n, P, scr.data_ptr(), out.data_ptr(), _stream()),
               "nppc_gram_schmidt_complex")
    G, A = _decode_gs_scratch(scr, B, n)
    return out, G, A


def gram_matrix_complex(v: torch.Tensor):
    """Hermitian Gram matrix G [B,m,m] complex128 (G[j,k] = v_j^H v_k) of m <= 12 complex vectors v [B,m,2,...]: the fp64 Gram
    pass of the Gram-Schmidt kernels (one read of v), taken from the scratch of a nppc_gram_schmidt_complex call whose
    orthogonalised output is discarded."""
    return gram_schmidt_complex_with_coeffs(v)[1]


def projection_loss(w_mat: torch.Tensor, gt: torch.Tensor, pred: torch.Tensor):
    """Loss statistics of NPPCAudioTrainer.base_step (trainer.py:259-298) for an explicit w_mat."""
    w_mat, gt, pred = _f32(w_mat), _f32(gt), _f32(pred)
    _chk(w_mat, gt, pred)
    B, n = w_mat.shape[:2]
    P = w_mat[0, 0, 0].numel()
    err_norm, err_proj, w_norms, reconst, sm = _loss_outputs(B, n, w_mat.device)
    scr = _gs_scratch(B, n, w_mat.device)
    _lib.check(_lib.load().nppc_projection_loss(w_mat.data_ptr(), gt.data_ptr(), pred.data_ptr(), B, n, P, scr.data_ptr(),
                                                err_norm.data_ptr(), err_proj.data_ptr(), w_norms.data_ptr(),
                                                reconst.data_ptr(), sm.data_ptr(), _stream()), "nppc_projection_loss")
    return dict(err_norm=err_norm, err_proj=torch.view_as_complex(err_proj), w_norms=w_norms, reconst_err=reconst,
                second_moment_mse=sm)


def _loss_outputs_real(B, n, device):
    f = dict(device=device, dtype=torch.float32)
    return (torch.empty(B, **f), torch.empty(B, n, **f), torch.empty(B, n, **f), torch.empty(B, **f), torch.empty(B, n, **f))


def gs_loss_fused_real(head: torch.Tensor, gt: torch.Tensor, pred: torch.Tensor):
    """Real Gram-Schmidt + the inpainting trainer's loss statistics (nppc_trainer.py:338-385). head [B,n,F,T]; gt/pred [B,1,F,T]."""
    head, gt, pred = _f32(head), _f32(gt), _f32(pred)
    _chk(head, gt, pred)
    B, n = head.shape[:2]
    P = head[0, 0].numel()
    assert gt.shape == pred.shape and gt[0].numel() == P
    w = torch.empty_like(head)
    err_norm, err_proj, w_norms, reconst, sm = _loss_outputs_real(B, n, head.device)
    scr = _gs_scratch(B, n, head.device)
    _lib.check(_lib.load().nppc_gs_loss_fused_real(head.data_ptr(), gt.data_ptr(), pred.data_ptr(), B, n, P, scr.data_ptr(),
                                                   w.data_ptr(), err_norm.data_ptr(), err_proj.data_ptr(),
                                                   w_norms.data_ptr(), reconst.data_ptr(), sm.data_ptr(), _stream()),
               "nppc_gs_loss_fused_real")
    return w, dict(err_norm=err_norm, err_proj=err_proj, w_norms=w_norms, reconst_err=reconst, second_moment_mse=sm)


def gs_loss_fused_real_with_gram(head: torch.Tensor, gt: torch.Tensor, pred: torch.Tensor):
    """gs_loss_fused_real that also returns what the backward needs from the forward's scratch: the symmetric Gram matrix
    G [B, n+1, n+1] float64 of (x_0 .. x_{n-1}, gt - pred) and the coefficient matrix A [B, n, n] float32 (w = A x).
    The real kernels use the .re slots of the same SampleScratch the complex ones fill (gram_schmidt.cu)."""
    head, gt, pred = _f32(head), _f32(gt), _f32(pred)
    _chk(head, gt, pred)
    B, n = head.shape[:2]
    P = head[0, 0].numel()
    assert gt.shape == pred.shape and gt[0].numel() == P
    w = torch.empty_like(head)
    err_norm, err_proj, w_norms, reconst, sm = _loss_outputs_real(B, n, head.device)
    scr = _gs_scratch(B, n, head.device)
    assert scr.numel() >= B * GS_SCRATCH_BYTES
    _lib.check(_lib.load().nppc_gs_loss_fused_real(head.data_ptr(), gt.data_ptr(), pred.data_ptr(), B, n, P, scr.data_ptr(),
                                                   w.data_ptr(), err_norm.data_ptr(), err_proj.data_ptr(),
                                                   w_norms.data_ptr(), reconst.data_ptr(), sm.data_ptr(), _stream()),
               "nppc_gs_loss_fused_real")
    G, A = _decode_gs_scratch_real(scr, B, n, n + 1)
    return w, dict(err_norm=err_norm, err_proj=err_proj, w_norms=w_norms, reconst_err=reconst, second_moment_mse=sm), G, A


def _decode_gs_scratch_real(scr: torch.Tensor, B: int, n: int, nv: int):
    """SampleScratch of the REAL kernels (the .re slots of the complex layout) -> (symmetric Gram matrix [B,nv,nv] float64 of the
    nv vectors of the Gram pass, coefficient matrix A [B,n,n] float32 with w = A x)."""
    scr = scr[:B * GS_SCRATCH_BYTES].reshape(B, GS_SCRATCH_BYTES)
    Gu = scr[:, :13 * 13 * 16].contiguous().view(torch.float64).reshape(B, 13, 13, 2)[:, :nv, :nv, 0]
    up = torch.triu(Gu, diagonal=1)
    G = torch.diag_embed(torch.diagonal(Gu, dim1=1, dim2=2)) + up + up.transpose(1, 2)
    A = scr[:, 13 * 13 * 16:].contiguous().view(torch.float32).reshape(B, 12, 12, 2)[:, :n, :n, 0].contiguous()
    return G, A


def gram_matrix_real(v: torch.Tensor):
    """Symmetric Gram matrix G [B,m,m] float64 of m <= 12 real vectors v [B,m,...]: the fp64 Gram pass of the real Gram-Schmidt
    kernel (one read of v), taken from the scratch of a nppc_gram_schmidt_real call whose output is discarded."""
    v = _f32(v)
    _chk(v)
    B, m = v.shape[:2]
    P = v[0, 0].numel()
    out = torch.empty_like(v)
    scr = _gs_scratch(B, m, v.device)
    assert scr.numel() >= B * GS_SCRATCH_BYTES
    _lib.check(_lib.load().nppc_gram_schmidt_real(v.data_ptr(), B, m, P, scr.data_ptr(), out.data_ptr(), _stream()),
               "nppc_gram_schmidt_real")
    return _decode_gs_scratch_real(scr, B, m, m)[0]


def real_lincomb(x: torch.Tensor, gt: torch.Tensor, pred: torch.Tensor, coef: torch.Tensor):
    """out[b,i] = sum_k coef[b,i,k] x[b,k] + coef[b,i,n] (gt - pred)[b] for REAL vectors: x [B,n,...], gt / pred [B,...],
    coef [B,n,n+1] real — the streaming pass of the inpainting head's Gram-Schmidt + loss backward.  Runs on
    nppc_complex_lincomb: with purely real coefficients the planar complex kernel treats its two planes independently, so the
    two halves of every real vector ride as its "real" and "imaginary" planes (needs an even vector length)."""
    x, gt, pred = _f32(x), _f32(gt), _f32(pred)
    B, n = x.shape[:2]
    P = x[0, 0].numel()
    if P % 2:
        raise ValueError(f"real_lincomb: vector length must be even (got {P})")
    assert gt[0].numel() == P and pred[0].numel() == P and tuple(coef.shape) == (B, n, n + 1) and not coef.is_complex()
    c = torch.complex(coef.float(), torch.zeros_like(coef, dtype=torch.float32))
    out = complex_lincomb(x.reshape(B, n, 2, P // 2), gt.reshape(B, 2, P // 2), pred.reshape(B, 2, P // 2), c)
    return out.reshape(x.shape)


def projection_loss_real(w_mat: torch.Tensor, gt: torch.Tensor, pred: torch.Tensor):
    """Loss statistics of the inpainting NPPC trainer's base_step for an explicit w_mat [B,n,F,T]."""
    w_mat, gt, pred = _f32(w_mat), _f32(gt), _f32(pred)
    _chk(w_mat, gt, pred)
    B, n = w_mat.shape[:2]
    P = w_mat[0, 0].numel()
    err_norm, err_proj, w_norms, reconst, sm = _loss_outputs_real(B, n, w_mat.device)
    scr = _gs_scratch(B, n, w_mat.device)
    _lib.check(_lib.load().nppc_projection_loss_real(w_mat.data_ptr(), gt.data_ptr(), pred.data_ptr(), B, n, P,
                                                     scr.data_ptr(), err_norm.data_ptr(), err_proj.data_ptr(),
                                                     w_norms.data_ptr(), reconst.data_ptr(), sm.data_ptr(), _stream()),
               "nppc_projection_loss_real")
    return dict(err_norm=err_norm, err_proj=err_proj, w_norms=w_norms, reconst_err=reconst, second_moment_mse=sm)


def logmag_normalize(clean_spec: torch.Tensor, masked_spec: torch.Tensor):
    """utils.preprocess_data (utils.py:294-306) without the mask expansion: spec [B,2,F,T] (re, im) ->
    (clean_log_norm [B,1,F,T], masked_log_norm [B,1,F,T], mean, std) with ONE global mean / unbiased std of the clean batch."""
    clean_spec, masked_spec = _f32(clean_spec), _f32(masked_spec)
    _chk(clean_spec, masked_spec)
    B, two, Fq, T = clean_spec.shape
    assert two == 2 and masked_spec.shape == clean_spec.shape
    P = Fq * T
    lib = _lib.load()
    sums = torch.empty(2, device=clean_spec.device, dtype=torch.float64)
    _lib.check(lib.nppc_logmag_stats(clean_spec.data_ptr(), B, P, sums.data_ptr(), _stream()), "nppc_logmag_stats")
    outs = []
    for sp in (clean_spec, masked_spec):
        o = torch.empty(B, 1, Fq, T, device=sp.device, dtype=torch.float32)
        _lib.check(lib.nppc_logmag_apply(sp.data_ptr(), B, P, sums.data_ptr(), B * P, o.data_ptr(), _stream()), "nppc_logmag_apply")
        outs.append(o)
    n = float(B * P)
    mean = sums[0] / n
    std = torch.sqrt((sums[1] - n * mean * mean) / (n - 1.0))
    return outs[0], outs[1], mean.float(), std.float()


def mask_blend(x_in, x: torch.Tensor, mask: torch.Tensor):
    """x_in[:, :1] * mask + x * (1 - mask)  (RestorationWrapper.forward, unet.py:298-313); x_in=None -> x * (1 - mask).
    x [B,C,F,T], mask [B,1,F,T] (or anything with B*F*T elements), x_in [B,Cin,F,T]."""
    x, mask = _f32(x), _f32(mask)
    _chk(x, mask)
    B, Cc = x.shape[:2]
    P = x[0, 0].numel()
    assert mask.numel() == B * P, "mask must be [B,1,F,T]"
    out = torch.empty_like(x)
    if x_in is not None:
        x_in = _f32(x_in)
        _chk(x_in)
        assert x_in.shape[0] == B and x_in[0, 0].numel() == P
    _lib.check(_lib.load().nppc_mask_blend(x_in.data_ptr() if x_in is not None else None, x_in.shape[1] if x_in is not None else 0,
                                           x.data_ptr(), mask.data_ptr(), B, Cc, P, out.data_ptr(), _stream()), "nppc_mask_blend")
    return out


def pc_variations(w_mat: torch.Tensor, noisy_real, noisy_imag, enh_real, enh_imag, alphas: torch.Tensor, want_pc=True):
    """N1: pc[b,d] = decompress(w_mat[b,d]) * noisy[b], var[b,d,a] = enhanced[b] + alphas[a] * pc[b,d] (validator.py:55-102,
    246-290) in one pass.  w_mat [B,n,2,F,T]; noisy/enh [B,F,T] (or [B,1,F,T]); returns (pc_re, pc_im [B,n,F,T] | None,
    var_re, var_im [B,n,A,F,T])."""
    w_mat, noisy_real, noisy_imag, enh_real, enh_imag = (_f32(t) for t in (w_mat, noisy_real, noisy_imag, enh_real, enh_imag))
    alphas = alphas.to(device=w_mat.device, dtype=torch.float32).contiguous()
    _chk(w_mat, noisy_real, noisy_imag, enh_real, enh_imag, alphas)
    B, n, two, Fq, T = w_mat.shape
    assert two == 2
    FT, A = Fq * T, alphas.numel()
    for t in (noisy_real, noisy_imag, enh_real, enh_imag):
        assert t.numel() == B * FT
    f = dict(device=w_mat.device, dtype=torch.float32)
    pc_re = torch.empty(B, n, Fq, T, **f) if want_pc else None
    pc_im = torch.empty(B, n, Fq, T, **f) if want_pc else None
    var_re, var_im = torch.empty(B, n, A, Fq, T, **f), torch.empty(B, n, A, Fq, T, **f)
    _lib.check(_lib.load().nppc_pc_variations(w_mat.data_ptr(), noisy_real.data_ptr(), noisy_imag.data_ptr(), enh_real.data_ptr(),
                                              enh_imag.data_ptr(), B, n, FT, alphas.data_ptr(), A,
                                              pc_re.data_ptr() if want_pc else None, pc_im.data_ptr() if want_pc else None,
                                              var_re.data_ptr(), var_im.data_ptr(), _stream()), "nppc_pc_variations")
    return pc_re, pc_im, var_re, var_im


def peak_normalize_(x: torch.Tensor):
    """In place x[r] /= max|x[r]| + 1e-8 over the last dimension (validator.py:118-134)."""
    _chk(x)
    L = x.shape[-1]
    _lib.check(_lib.load().nppc_peak_normalize(x.data_ptr(), x.numel() // L, L, _stream()), "nppc_peak_normalize")
    return x


# ---- N3: GPU data preparation ---------------------------------------------------------------------------------------------
def mix_with_snr(clean: torch.Tensor, noise: torch.Tensor, snr_db, target_db=-25.0):
    """Batched AudioDataset._mix_with_snr (dataset/audio_dataset.py:92-158): clean / noise [B,L]; snr_db, target_db: scalar or
    [B].  Returns (noisy [B,L], clean_normalised [B,L])."""
    clean, noise = _f32(clean), _f32(noise)
    _chk(clean, noise)
    B, L = clean.shape
    assert noise.shape == clean.shape
    as_vec = lambda v: (v.to(device=clean.device, dtype=torch.float32).reshape(-1).expand(B) if torch.is_tensor(v)
                        else torch.full((B,), float(v), device=clean.device, dtype=torch.float32)).contiguous()
    snr, tgt = as_vec(snr_db), as_vec(target_db)
    lib = _lib.load()
    scr = torch.empty(lib.nppc_mix_scratch_bytes(B), device=clean.device, dtype=torch.uint8)
    noisy, clean_out = torch.empty_like(clean), torch.empty_like(clean)
    _lib.check(lib.nppc_mix_with_snr(clean.data_ptr(), noise.data_ptr(), B, L, snr.data_ptr(), tgt.data_ptr(), scr.data_ptr(),
                                     noisy.data_ptr(), clean_out.data_ptr(), _stream()), "nppc_mix_with_snr")
    return noisy, clean_out


def time_to_spec_mask(mask_time: torch.Tensor, T_frames: int, win_length: int, hop_length: int, center: bool = True):
    """Batched AudioInpaintingDataset.time_to_spec_mask (dataset/audio_dataset_inpainting.py:223-251): mask_time [B,L] (1 = keep)
    -> [B,T_frames] with 1 where the whole analysis window of the frame is unmasked."""
    mask_time = _f32(mask_time)
    _chk(mask_time)
    B, L = mask_time.shape
    out = torch.empty(B, T_frames, device=mask_time.device, dtype=torch.float32)
    _lib.check(_lib.load().nppc_time_to_spec_mask(mask_time.data_ptr(), B, L, T_frames, win_length, hop_length, int(center),
                                                  out.data_ptr(), _stream()), "nppc_time_to_spec_mask")
    return out


TC_ROW_TILE = 128  # the tensor-core LSTM owns 128 sequences per CTA; its time-major buffers pad rows to this


def padded_rows(R: int, dtype) -> int:
    return -(-R // TC_ROW_TILE) * TC_ROW_TILE if dtype == _F16 else R


def subband_pack(nbr_src, fb, fbr, fbi, num_neighbor: int, groups: int, KP: int = 64, dtype=torch.float32, pad_rows: bool = False,
                 want_sums: bool = False, cumulative: bool = False):
    """Fused unfold ++ cat ++ offline_laplace_norm ++ drop_band -> time-major LSTM input [T', R_stride, KP]
    (R = B*F' real rows; rows R..R_stride-1 are zero padding for the tensor-core paths: always for fp16, for fp32 when
    pad_rows). Returns (xs, R) or (xs, R, sums [B] fp64 = per-sample sum of the un-normalised sub-band tensor).
    cumulative=True: norm_type "cumulative_laplace_norm" (running mean of each row's own features) instead of the offline norm."""
    nbr_src, fb, fbr, fbi = (_f32(v) for v in (nbr_src, fb, fbr, fbi))
    _chk(nbr_src, fb, fbr, fbi)
    B, F, Tp = nbr_src.shape
    G = groups if (groups > 1 and B > 1) else 1
    R = B * (F // G)
    RS = padded_rows(R, _F16 if pad_rows else dtype)
    xs = torch.empty(Tp, RS, KP, device=fb.device, dtype=dtype)
    sums = torch.empty(B, device=fb.device, dtype=torch.float64)
    _lib.check(_lib.load().nppc_subband_pack(nbr_src.data_ptr(), fb.data_ptr(), fbr.data_ptr(), fbi.data_ptr(), B, F, Tp,
                                             num_neighbor, groups, KP, RS, int(cumulative), sums.data_ptr(),
                                             xs.data_ptr() if dtype == torch.float32 else 0,
                                             xs.data_ptr() if dtype == _F16 else 0, _stream()), "nppc_subband_pack")
    return (xs, R, sums) if want_sums else (xs, R)


class LstmPlan:
    """Owns the re-packed weight caches of one 2-layer LSTM + fc (sequence_model.py:31-38,79)."""

    def __init__(self, w_ih0, w_hh0, b_ih0, b_hh0, w_ih1, w_hh1, b_ih1, b_hh1, fc_w, fc_b):
        import ctypes as C
        ts = [_f32(t.detach()) for t in (w_ih0, w_hh0, b_ih0, b_hh0, w_ih1, w_hh1, b_ih1, b_hh1, fc_w, fc_b)]
        _chk(*ts)
        self.H = ts[1].shape[1]
        self.I = ts[0].shape[1]
        self.O = ts[8].shape[0]
        h = C.c_void_p()
        _lib.check(_lib.load().nppc_lstm_plan_create(C.byref(h), self.I, self.H, self.O, *[t.data_ptr() for t in ts],
                                                     _stream()), "nppc_lstm_plan_create")
        torch.cuda.current_stream().synchronize()  # the source tensors may be temporaries
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().nppc_lstm_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def forward(self, xs: torch.Tensor, impl: int, R: int = None):
        """xs [T', R_stride, KP] (fp32 for impl 0; fp16 with R_stride = R rounded up to 128 for impl 1) -> y [R, O, T']."""
        _chk(xs)
        Tp, RS, KP = xs.shape
        R = RS if R is None else R
        if impl == 0 and xs.dtype != torch.float32 or impl == 1 and xs.dtype != _F16:
            raise RuntimeError("LstmPlan.forward: xs dtype does not match impl")
        lib = _lib.load()
        nbytes = lib.nppc_lstm_workspace_bytes(self._h, R, Tp, impl)
        ws = torch.empty(max(nbytes, 16), device=xs.device, dtype=torch.uint8)
        y = torch.empty(R, self.O, Tp, device=xs.device, dtype=torch.float32)
        _lib.check(lib.nppc_lstm_forward(self._h, xs.data_ptr(), R, RS, Tp, KP, impl, ws.data_ptr(), nbytes, y.data_ptr(),
                                         _stream()), "nppc_lstm_forward")
        return y


# ---- a7 implementation 2: stepwise tensor-core LSTM (generic I / H, saved state, BPTT) -----------------------------------
def _lstm_structs():
    import ctypes as C

    class W(C.Structure):
        _fields_ = [("w_ih", C.c_void_p * 2), ("w_hh", C.c_void_p * 2), ("b_ih", C.c_void_p * 2), ("b_hh", C.c_void_p * 2),
                    ("fc_w", C.c_void_p), ("fc_b", C.c_void_p), ("I", C.c_int), ("H", C.c_int), ("O", C.c_int)]

    class G(C.Structure):
        _fields_ = [("w_ih", C.c_void_p * 2), ("w_hh", C.c_void_p * 2), ("b_ih", C.c_void_p * 2), ("b_hh", C.c_void_p * 2),
                    ("fc_w", C.c_void_p), ("fc_b", C.c_void_p)]
    return W, G


def _lstm_wstruct(params):
    """params: (w_ih0, w_hh0, b_ih0, b_hh0, w_ih1, w_hh1, b_ih1, b_hh1, fc_w, fc_b) fp32 CUDA tensors (nn.LSTM layout)."""
    import ctypes as C
    W, _ = _lstm_structs()
    ts = [_f32(t.detach()) for t in params]
    _chk(*ts)
    w = W()
    w.w_ih = (C.c_void_p * 2)(ts[0].data_ptr(), ts[4].data_ptr())
    w.w_hh = (C.c_void_p * 2)(ts[1].data_ptr(), ts[5].data_ptr())
    w.b_ih = (C.c_void_p * 2)(ts[2].data_ptr(), ts[6].data_ptr())
    w.b_hh = (C.c_void_p * 2)(ts[3].data_ptr(), ts[7].data_ptr())
    w.fc_w, w.fc_b = ts[8].data_ptr(), ts[9].data_ptr()
    w.I, w.H, w.O = ts[0].shape[1], ts[1].shape[1], ts[8].shape[0]
    assert ts[0].shape[0] == 4 * w.H and ts[4].shape == (4 * w.H, w.H) and ts[8].shape[1] == w.H
    return w, ts


def lstm_step_forward(params, xs: torch.Tensor, R: int = None, train: bool = False, precise: bool = False):
    """Stepwise tensor-core LSTM + fc: xs [T', R_stride, KP] (fp16; fp32 when precise) -> (y [R, O, T'] f32, workspace).
    train=True keeps gates / c / h of every step in the returned workspace for lstm_step_backward."""
    import ctypes as C
    _chk(xs)
    Tp, RS, KP = xs.shape
    R = RS if R is None else R
    if xs.dtype != (torch.float32 if precise else _F16):
        raise RuntimeError("lstm_step_forward: xs must be fp16 (fast mode) or fp32 (precise mode)")
    w, keep = _lstm_wstruct(params)
    lib = _lib.load()
    nbytes = lib.nppc_lstm_step_workspace_bytes(w.I, w.H, w.O, RS, Tp, KP, int(train), int(precise))
    ws = torch.empty(nbytes, device=xs.device, dtype=torch.uint8)
    y = torch.empty(R, w.O, Tp, device=xs.device, dtype=torch.float32)
    _lib.check(lib.nppc_lstm_step_forward(C.byref(w), xs.data_ptr(), int(precise), R, RS, Tp, KP, int(train), int(precise),
                                          ws.data_ptr(), nbytes, y.data_ptr(), _stream()), "nppc_lstm_step_forward")
    return y, ws


def lstm_step_backward(params, xs: torch.Tensor, R: int, ws: torch.Tensor, dy: torch.Tensor, want_dxs: bool = True):
    """BPTT of lstm_step_forward(train=True): dy [R, O, T'] -> (list of 10 parameter gradients in `params` order, dxs or None)."""
    import ctypes as C
    dy = _f32(dy)
    _chk(xs, ws, dy)
    Tp, RS, KP = xs.shape
    w, keep = _lstm_wstruct(params)
    _, G = _lstm_structs()
    gs = [torch.empty_like(t) for t in keep]
    g = G()
    g.w_ih = (C.c_void_p * 2)(gs[0].data_ptr(), gs[4].data_ptr())
    g.w_hh = (C.c_void_p * 2)(gs[1].data_ptr(), gs[5].data_ptr())
    g.b_ih = (C.c_void_p * 2)(gs[2].data_ptr(), gs[6].data_ptr())
    g.b_hh = (C.c_void_p * 2)(gs[3].data_ptr(), gs[7].data_ptr())
    g.fc_w, g.fc_b = gs[8].data_ptr(), gs[9].data_ptr()
    dxs = torch.empty(Tp, RS, KP, device=xs.device, dtype=torch.float32) if want_dxs else None
    _lib.check(_lib.load().nppc_lstm_step_backward(C.byref(w), xs.data_ptr(), R, RS, Tp, KP, ws.data_ptr(), ws.numel(), dy.data_ptr(),
                                                   C.byref(g), _ptr(dxs), _stream()), "nppc_lstm_step_backward")
    return gs, dxs


def gemm_f16_atb(a: torch.Tensor, b: torch.Tensor, splits: int = None):
    """C [Mo, No] f32 = A^T B, A [rows, Mo], B [rows, No] fp16 row-major (tcgen05 MN-major operands, deterministic split-K)."""
    _chk(a, b)
    assert a.dtype == _F16 and b.dtype == _F16 and a.shape[0] == b.shape[0]
    rows, Mo = a.shape
    No = b.shape[1]
    if splits is None:
        tiles = (Mo // 128) * (No // (128 if No % 128 == 0 else 64))
        splits = max(1, min(64, 148 // max(tiles, 1), rows // 64))
    part = torch.empty(splits, Mo, No, device=a.device, dtype=torch.float32)
    c = torch.empty(Mo, No, device=a.device, dtype=torch.float32)
    _lib.check(_lib.load().nppc_gemm_f16_atb(a.data_ptr(), b.data_ptr(), rows, Mo, No, splits, part.data_ptr(), c.data_ptr(), _stream()),
               "nppc_gemm_f16_atb")
    return c


def tsse(x: torch.Tensor, kersize, conv_w, conv_b, fcat_w, fcat_b, fc1_w, fc1_b, fc2_w, fc2_b):
    """ChannelTimeSenseSELayer.forward (attention_model.py:78-98): x [B,C,T] -> x * gate."""
    import ctypes as C
    x = _f32(x)
    ts = [_f32(t.detach()) for t in (*conv_w, *conv_b, fcat_w, fcat_b, fc1_w, fc1_b, fc2_w, fc2_b)]
    _chk(x, *ts)
    B, Cc, T = x.shape
    y = torch.empty_like(x)
    scratch = torch.empty(2 * B * Cc, device=x.device, dtype=torch.float32)
    ks = (C.c_int * 3)(*[int(k) for k in kersize])
    cw = (C.c_void_p * 3)(*[t.data_ptr() for t in ts[0:3]])
    cb = (C.c_void_p * 3)(*[t.data_ptr() for t in ts[3:6]])
    _lib.check(_lib.load().nppc_tsse(x.data_ptr(), B, Cc, T, ks, cw, cb, *[t.data_ptr() for t in ts[6:12]],
                                     ts[8].shape[0], scratch.data_ptr(), y.data_ptr(), _stream()), "nppc_tsse")
    return y


def prelu_stats(y: torch.Tensor, prelu_a: torch.Tensor, bias: torch.Tensor = None):
    """[B,2] fp64 (sum, sum of squares) of PReLU(y + bias[c]) per sample, y [B,C,T]."""
    _chk(y, prelu_a)
    B, Cc, T = y.shape
    stats = torch.empty(B, 2, device=y.device, dtype=torch.float64)
    _lib.check(_lib.load().nppc_prelu_stats(y.data_ptr(), B, Cc, T, bias.data_ptr() if bias is not None else None,
                                            prelu_a.data_ptr(), stats.data_ptr(), _stream()), "nppc_prelu_stats")
    return stats


def tcn_mid(y1, prelu1_a, stats1, gamma1, beta1, dw_w, dw_b, dilation: int, prelu2_a, bias1=None):
    """z = PReLU2(depthwise(GroupNorm1(PReLU1(y1 + bias1)))) and the per-sample moments of z (causal_conv.py:100-104)."""
    _chk(y1, prelu1_a, stats1, gamma1, beta1, dw_w, dw_b, prelu2_a)
    B, Cc, T = y1.shape
    z = torch.empty_like(y1)
    stats2 = torch.empty(B, 2, device=y1.device, dtype=torch.float64)
    _lib.check(_lib.load().nppc_tcn_mid(y1.data_ptr(), B, Cc, T, bias1.data_ptr() if bias1 is not None else None,
                                        prelu1_a.data_ptr(), stats1.data_ptr(), gamma1.data_ptr(),
                                        beta1.data_ptr(), dw_w.data_ptr(), dw_b.data_ptr(), dilation, prelu2_a.data_ptr(),
                                        z.data_ptr(), stats2.data_ptr(), _stream()), "nppc_tcn_mid")
    return z, stats2


def tcn_out(o, x, c_hidden: int, stats2, u, vb):
    """x + sconv(GroupNorm2(z)) with the norm folded: x + o*rstd + vb - mean*rstd*u (causal_conv.py:105-108)."""
    _chk(o, x, stats2, u, vb)
    B, Cc, T = x.shape
    xn = torch.empty_like(x)
    _lib.check(_lib.load().nppc_tcn_out(o.data_ptr(), x.data_ptr(), B, Cc, T, c_hidden, stats2.data_ptr(), u.data_ptr(),
                                        vb.data_ptr(), xn.data_ptr(), _stream()), "nppc_tcn_out")
    return xn


# ---- N2: channel-last TCN stack on the tcgen05 GEMM ------------------------------------------------------------------
def gemm_f16_tn(a: torch.Tensor, w: torch.Tensor, bias=None):
    """C[M,N] fp16 = A[M,K] fp16 @ W[N,K]^T fp16 (+ bias f32), fp32 accumulate (tcgen05).  K % 64 == 0, N % 128 == 0."""
    _chk(a, w, bias)
    assert a.dtype == torch.float16 and w.dtype == torch.float16
    M, K = a.shape
    N = w.shape[0]
    c = torch.empty(M, N, device=a.device, dtype=torch.float16)
    _lib.check(_lib.load().nppc_gemm_f16_tn(a.data_ptr(), w.data_ptr(), _ptr(bias), c.data_ptr(), M, N, K, _stream()),
               "nppc_gemm_f16_tn")
    return c


def gemm_f16_tn_ex(a: torch.Tensor, w: torch.Tensor, bias=None, out_f32: bool = False):
    """C[M,N] = sum over the K-blocks kb of W[N,K]: A[M, KA] block (kb % (KA/64)) x W block kb (fp16 operands, fp32
    accumulate); out fp16 or fp32.  KA = a.shape[1] <= K = w.shape[1]."""
    _chk(a, w, bias)
    assert a.dtype == torch.float16 and w.dtype == torch.float16
    M, KA = a.shape
    N, K = w.shape
    c = torch.empty(M, N, device=a.device, dtype=torch.float32 if out_f32 else torch.float16)
    _lib.check(_lib.load().nppc_gemm_f16_tn_ex(a.data_ptr(), w.data_ptr(), _ptr(bias), c.data_ptr(), M, N, K, KA, int(out_f32),
                                               _stream()), "nppc_gemm_f16_tn_ex")
    return c


def tcn_cl_scale(x: torch.Tensor):
    """per-sample fp16 range scale of x [B, ...] f32: (scale [B], inv_scale [B]) with scale = max(max|x[b]|, 1)."""
    _chk(x)
    B = x.shape[0]
    scale = torch.empty(B, device=x.device, dtype=torch.float32)
    inv = torch.empty_like(scale)
    _lib.check(_lib.load().nppc_tcn_cl_scale(x.data_ptr(), B, x[0].numel(), scale.data_ptr(), inv.data_ptr(), _stream()),
               "nppc_tcn_cl_scale")
    return scale, inv


def tcn_cl_pack(x: torch.Tensor, Kp: int, inv_scale: torch.Tensor, x32: torch.Tensor, xh: torch.Tensor, split: bool = False):
    _chk(x, inv_scale, x32, xh)
    B, C, T = x.shape
    assert xh.shape[1] == (2 * Kp if split else Kp)
    _lib.check(_lib.load().nppc_tcn_cl_pack(x.data_ptr(), B, C, T, Kp, inv_scale.data_ptr(), x32.data_ptr(), xh.data_ptr(), int(split),
                                            _stream()), "nppc_tcn_cl_pack")


def tcn_cl_unpack(o: torch.Tensor, B: int, C: int, T: int, Np: int, scale, bias, relu: int):
    _chk(o, scale, bias)
    assert o.dtype in (torch.float16, torch.float32)
    out = torch.empty(B, C, T, device=o.device, dtype=torch.float32)
    _lib.check(_lib.load().nppc_tcn_cl_unpack(o.data_ptr(), int(o.dtype == torch.float32), B, C, T, Np, _ptr(scale), _ptr(bias),
                                              int(relu), out.data_ptr(), _stream()), "nppc_tcn_cl_unpack")
    return out


def prelu_stats_cl(y1: torch.Tensor, B: int, T: int, scale, bias, prelu_a):
    _chk(y1, scale, bias, prelu_a)
    stats = torch.empty(B, 2, device=y1.device, dtype=torch.float64)
    _lib.check(_lib.load().nppc_prelu_stats_cl(y1.data_ptr(), B, T, y1.shape[1], scale.data_ptr(), bias.data_ptr(), prelu_a.data_ptr(),
                                               stats.data_ptr(), _stream()), "nppc_prelu_stats_cl")
    return stats


def tcn_mid_cl(y1, B: int, T: int, scale, bias1, prelu1_a, stats1, gamma1, beta1, dw_w, dw_b, dilation: int, prelu2_a):
    _chk(y1, scale, bias1, prelu1_a, stats1, gamma1, beta1, dw_w, dw_b, prelu2_a)
    z = torch.empty_like(y1)
    stats2 = torch.empty(B, 2, device=y1.device, dtype=torch.float64)
    _lib.check(_lib.load().nppc_tcn_mid_cl(y1.data_ptr(), B, T, y1.shape[1], scale.data_ptr(), bias1.data_ptr(), prelu1_a.data_ptr(), stats1.data_ptr(),
                                           gamma1.data_ptr(), beta1.data_ptr(), dw_w.data_ptr(), dw_b.data_ptr(), dilation,
                                           prelu2_a.data_ptr(), z.data_ptr(), stats2.data_ptr(), _stream()), "nppc_tcn_mid_cl")
    return z, stats2


def tcn_out_cl(o, x32, B: int, T: int, C: int, Np: int, Kp: int, stats2, u, vb, inv_scale, xh, relu_h: bool, split: bool = False):
    _chk(o, x32, stats2, u, vb, inv_scale, xh)
    assert xh.shape[1] == (2 * Kp if split else Kp)
    _lib.check(_lib.load().nppc_tcn_out_cl(o.data_ptr(), x32.data_ptr(), B, T, C, Np, Kp, 512, stats2.data_ptr(), u.data_ptr(),
                                           vb.data_ptr(), inv_scale.data_ptr(), xh.data_ptr(), int(relu_h), int(split), _stream()),
               "nppc_tcn_out_cl")


# ---- N4: UNet convolutions on tcgen05 (NHWC fp16 implicit GEMM) --------------------------------------------------------------
def conv3x3_pack_weights(w: torch.Tensor, C0: int, C1: int = 0):
    """w [Cout, C0 + C1, 3, 3] fp32 (BatchNorm folded) -> fp16 [Cout, 9 * (C0p + C1p)] for conv3x3_tc."""
    w = _f32(w)
    _chk(w)
    Cout = w.shape[0]
    assert w.shape[1] == C0 + C1 and w.shape[2:] == (3, 3)
    C0p, C1p = -(-C0 // 64) * 64, -(-C1 // 64) * 64
    out = torch.empty(Cout, 9 * (C0p + C1p), device=w.device, dtype=_F16)
    _lib.check(_lib.load().nppc_conv3x3_pack_weights(w.data_ptr(), Cout, C0, C1, out.data_ptr(), _stream()), "nppc_conv3x3_pack_weights")
    return out


def conv3x3_tc(x0: torch.Tensor, x1, w_packed: torch.Tensor, bias: torch.Tensor, negative_slope: float = 0.2):
    """leaky_relu(conv3x3(cat(x0, x1), padding=1) + bias): x0 [B,H,W,C0p], x1 [B,H,W,C1p] or None (NHWC fp16) -> [B,H,W,Cout] fp16."""
    _chk(x0, x1, w_packed, bias)
    assert x0.dtype == _F16 and w_packed.dtype == _F16 and bias.dtype == torch.float32
    B, H, W, C0p = x0.shape
    C1p = 0 if x1 is None else x1.shape[3]
    assert x1 is None or (x1.dtype == _F16 and x1.shape[:3] == x0.shape[:3])
    Cout = w_packed.shape[0]
    assert w_packed.shape[1] == 9 * (C0p + C1p)
    y = torch.empty(B, H, W, Cout, device=x0.device, dtype=_F16)
    _lib.check(_lib.load().nppc_conv3x3_tc(x0.data_ptr(), C0p, _ptr(x1), C1p, w_packed.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, W, Cout,
                                           float(negative_slope), _stream()), "nppc_conv3x3_tc")
    return y


def nchw_to_nhwc_f16(x: torch.Tensor, Cp: int = 64):
    x = _f32(x)
    _chk(x)
    B, C, H, W = x.shape
    y = torch.empty(B, H, W, Cp, device=x.device, dtype=_F16)
    _lib.check(_lib.load().nppc_nchw_to_nhwc_f16(x.data_ptr(), B, C, H, W, Cp, y.data_ptr(), _stream()), "nppc_nchw_to_nhwc_f16")
    return y


def maxpool2x2_nhwc(x: torch.Tensor):
    _chk(x)
    B, H, W, C = x.shape
    assert x.dtype == _F16
    y = torch.empty(B, H // 2, W // 2, C, device=x.device, dtype=_F16)
    _lib.check(_lib.load().nppc_maxpool2x2_nhwc(x.data_ptr(), B, H, W, C, y.data_ptr(), _stream()), "nppc_maxpool2x2_nhwc")
    return y


def upsample2x_pad_nhwc(x: torch.Tensor, H: int, W: int):
    """bilinear x2 (align_corners=True) of x [B,h,w,C] fp16, zero-padded (centred like tmp_utils.py:76-82) to [B,H,W,C]."""
    _chk(x)
    B, h, w, C = x.shape
    assert x.dtype == _F16
    y = torch.empty(B, H, W, C, device=x.device, dtype=_F16)
    _lib.check(_lib.load().nppc_upsample2x_pad_nhwc(x.data_ptr(), B, h, w, C, H, W, y.data_ptr(), _stream()), "nppc_upsample2x_pad_nhwc")
    return y


def conv1x1_out(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor):
    """unet outc: x [B,H,W,Cin] fp16 NHWC, w [Cout,Cin(,1,1)] fp32 -> [B,Cout,H,W] fp32."""
    w, bias = _f32(w.reshape(w.shape[0], -1)), _f32(bias)
    _chk(x, w, bias)
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.shape[1] == Cin and x.dtype == _F16
    y = torch.empty(B, Cout, H, W, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().nppc_conv1x1_out(x.data_ptr(), B, H * W, Cin, w.data_ptr(), bias.data_ptr(), Cout, y.data_ptr(), _stream()),
               "nppc_conv1x1_out")
    return y


def assemble_mask(y: torch.Tensor, B: int, Fp: int, look_ahead: int):
    """y [B*F', O, T'] -> [B, O, F', T'-la] (fullsubnet_plus.py:227-229)."""
    _chk(y)
    R, O, Tp = y.shape
    assert R == B * Fp
    out = torch.empty(B, O, Fp, Tp - look_ahead, device=y.device, dtype=torch.float32)
    _lib.check(_lib.load().nppc_assemble_mask(y.data_ptr(), B, Fp, O, Tp, look_ahead, out.data_ptr(), _stream()),
               "nppc_assemble_mask")
    return out


def gemm_bf16_tn(a: torch.Tensor, w: torch.Tensor, bias=None):
    """C[M,N] bf16 = A[M,K] bf16 @ W[N,K]^T bf16 + bias (tcgen05)."""
    _chk(a, w, bias)
    M, K = a.shape
    N = w.shape[0]
    c = torch.empty(M, N, device=a.device, dtype=_BF16)
    _lib.check(_lib.load().nppc_gemm_bf16_tn(a.data_ptr(), w.data_ptr(), _ptr(bias), c.data_ptr(), M, N, K, _stream()),
               "nppc_gemm_bf16_tn")
    return c
