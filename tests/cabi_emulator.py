"""TEST INFRASTRUCTURE ONLY — a host-memory emulation of a handful of C-ABI entry points of libnppc_b200.so, written from the
CONTRACT in include/nppc_b200.h (argument order, layouts, the SampleScratch the Gram-Schmidt kernels leave behind), in numpy on
CPU tensors' memory.  Purpose: run the product's REAL `ops.py` wrappers (pointer marshalling, scratch decoding, reshapes) and
the autograd Functions above them without a GPU, so that code written while no GPU was available is exercised end to end
against an independent restatement of what the kernels promise.  Never imported by the product; nothing here is timed.

Emulated: nppc_mask_blend, nppc_logmag_stats / _apply, nppc_gram_schmidt_complex / _real, nppc_gs_loss_fused_real,
nppc_complex_lincomb.  Host-only entry points (nppc_gs_scratch_bytes, nppc_last_error, ...) go to the real library."""
import ctypes

import numpy as np

NV_MAX, NA = 13, 12
G_BYTES, A_BYTES = NV_MAX * NV_MAX * 2 * 8, NA * NA * 2 * 4
SCR_BYTES = G_BYTES + A_BYTES                      # sizeof(SampleScratch), csrc/gram_schmidt.cu


def _arr(ptr, count, ctype, dtype):
    if not ptr:
        return None
    return np.frombuffer((ctype * int(count)).from_address(int(ptr)), dtype=dtype)


def f32(ptr, count):
    return _arr(ptr, count, ctypes.c_float, np.float32)


def f64(ptr, count):
    return _arr(ptr, count, ctypes.c_double, np.float64)


def _solve(G, n, nv, has_err, cplx, eps_add):
    """The coefficient-space MGS of gs_solve (reference recurrences incl. the conjugated coefficient) -> A [n,n], stats."""
    A = np.zeros((n, n), dtype=np.complex128)
    ahat, v = [], []
    stats = dict(err_proj=np.zeros(n, np.complex128), w_norms=np.zeros(n), second=np.zeros(n))
    eps_n = np.sqrt(max(G[n, n].real, 0.0)) if has_err else 0.0
    acc = 0.0
    for i in range(n):
        a = np.zeros(n, np.complex128)
        a[i] = 1.0
        for j in range(i):
            c = np.sum(np.conj(a) * v[j])          # sum_p conj(w[p]) what_j[p] = sum_k conj(a[k]) (G ahat_j)[k]
            a = a - ahat[j] * c
        nrm = np.sqrt(max((np.conj(a) @ (G[:n, :n] @ a)).real, 0.0))
        A[i] = a
        ahat.append(a / nrm)
        v.append(G[:n, :n] @ (a / nrm))
        if has_err:
            pr = np.sum(np.conj(a) * G[:n, n]) / ((nrm + eps_add) * (eps_n + eps_add))
            wn = (nrm + eps_add) / (eps_n + eps_add) if not cplx else nrm / (eps_n + eps_add)
            pm2 = abs(pr) ** 2
            stats["err_proj"][i], stats["w_norms"][i], stats["second"][i] = pr, wn, (wn * wn - pm2) ** 2
            acc += pm2
    if has_err:
        stats["reconst"] = 1.0 - acc
        stats["err_norm"] = eps_n + (0.0 if cplx else eps_add)
    return A, stats


def _run_gs(x_ptr, gt_ptr, pred_ptr, B, n, P, scr_ptr, out_ptr, cplx, outs=None):
    comp = 2 if cplx else 1
    x = f32(x_ptr, B * n * comp * P).reshape(B, n, comp, P).astype(np.float64)
    has_err = bool(gt_ptr)
    if has_err:
        e = (f32(gt_ptr, B * comp * P).astype(np.float64) - f32(pred_ptr, B * comp * P).astype(np.float64)).reshape(B, 1, comp, P)
        vecs = np.concatenate([x, e], axis=1)
    else:
        vecs = x
    nv = vecs.shape[1]
    vc = vecs[:, :, 0] + (1j * vecs[:, :, 1] if cplx else 0)                   # [B, nv, P]
    scr = _arr(scr_ptr, B * SCR_BYTES, ctypes.c_uint8, np.uint8).reshape(B, SCR_BYTES)
    out = f32(out_ptr, B * n * comp * P).reshape(B, n, comp, P) if out_ptr else None
    for b in range(B):
        G = np.conj(vc[b]) @ vc[b].T                                           # G[j,k] = v_j^H v_k
        Gs = np.zeros((NV_MAX, NV_MAX, 2))
        for j in range(nv):
            for k in range(j, nv):                                             # the kernels fill the UPPER triangle only
                Gs[j, k] = (G[j, k].real, G[j, k].imag)
        A, st = _solve(G, n, nv, has_err, cplx, 1e-8 if cplx else 1e-6)
        As = np.zeros((NA, NA, 2), dtype=np.float32)
        As[:n, :n, 0], As[:n, :n, 1] = A.real, A.imag
        scr[b, :G_BYTES] = np.frombuffer(Gs.tobytes(), dtype=np.uint8)
        scr[b, G_BYTES:] = np.frombuffer(As.tobytes(), dtype=np.uint8)
        if out is not None:
            Af = As[:n, :n, 0].astype(np.float64) + 1j * As[:n, :n, 1].astype(np.float64)   # fp32 coefficients, as the apply pass
            w = Af @ vc[b, :n]
            w[0] = vc[b, 0]                                                    # direction 0 is returned untouched
            out[b, :, 0] = w.real
            if cplx:
                out[b, :, 1] = w.imag
        if has_err and outs is not None:
            err_norm, err_proj, w_norms, reconst, second = outs
            f32(err_norm, B)[b] = st["err_norm"]
            if cplx:
                f32(err_proj, B * n * 2).reshape(B, n, 2)[b] = np.stack([st["err_proj"].real, st["err_proj"].imag], -1)
            else:
                f32(err_proj, B * n).reshape(B, n)[b] = st["err_proj"].real
            f32(w_norms, B * n).reshape(B, n)[b] = st["w_norms"]
            f32(reconst, B)[b] = st["reconst"]
            f32(second, B * n).reshape(B, n)[b] = st["second"]
    return 0


class EmulatedLib:
    def __init__(self, real_lib):
        self._real = real_lib
        self.calls = []

    def __getattr__(self, name):                   # host-only entry points: the real library
        return getattr(self._real, name)

    def nppc_mask_blend(self, x_in, Cin, x, mask, B, C, P, out, stream):
        self.calls.append("nppc_mask_blend")
        xv, mv, ov = f32(x, B * C * P).reshape(B, C, P), f32(mask, B * P).reshape(B, 1, P), f32(out, B * C * P).reshape(B, C, P)
        ov[:] = xv * (1 - mv)
        if x_in:
            ov += f32(x_in, B * Cin * P).reshape(B, Cin, P)[:, :1] * mv
        return 0

    def nppc_logmag_stats(self, spec, B, P, sums, stream):
        self.calls.append("nppc_logmag_stats")
        s = f32(spec, B * 2 * P).reshape(B, 2, P).astype(np.float64)
        lm = np.log(np.sqrt(s[:, 0] ** 2 + s[:, 1] ** 2).astype(np.float32).astype(np.float64) + 1e-6)
        f64(sums, 2)[:] = (lm.sum(), (lm * lm).sum())
        return 0

    def nppc_logmag_apply(self, spec, B, P, sums, n_stat, out, stream):
        self.calls.append("nppc_logmag_apply")
        s = f32(spec, B * 2 * P).reshape(B, 2, P).astype(np.float64)
        lm = np.log(np.sqrt(s[:, 0] ** 2 + s[:, 1] ** 2).astype(np.float32).astype(np.float64) + 1e-6)
        sm = f64(sums, 2)
        mean = sm[0] / n_stat
        std = np.sqrt((sm[1] - n_stat * mean * mean) / (n_stat - 1.0))
        f32(out, B * P).reshape(B, P)[:] = (lm - mean) / std
        return 0

    def nppc_gram_schmidt_complex(self, x, B, n, P, scratch, out, stream):
        self.calls.append("nppc_gram_schmidt_complex")
        return _run_gs(x, 0, 0, B, n, P, scratch, out, True)

    def nppc_gram_schmidt_real(self, x, B, n, P, scratch, out, stream):
        self.calls.append("nppc_gram_schmidt_real")
        return _run_gs(x, 0, 0, B, n, P, scratch, out, False)

    def nppc_gs_loss_fused_real(self, x, gt, pred, B, n, P, scratch, w_mat, err_norm, err_proj, w_norms, reconst, second, stream):
        self.calls.append("nppc_gs_loss_fused_real")
        return _run_gs(x, gt, pred, B, n, P, scratch, w_mat, False, (err_norm, err_proj, w_norms, reconst, second))

    def nppc_complex_lincomb(self, x, gt, pred, B, n, P, coef, out, stream):
        self.calls.append("nppc_complex_lincomb")
        xv = f32(x, B * n * 2 * P).reshape(B, n, 2, P).astype(np.float64)
        e = (f32(gt, B * 2 * P).astype(np.float64) - f32(pred, B * 2 * P).astype(np.float64)).reshape(B, 1, 2, P)
        v = np.concatenate([xv, e], axis=1)
        vc = v[:, :, 0] + 1j * v[:, :, 1]
        c = f32(coef, B * n * (n + 1) * 2).reshape(B, n, n + 1, 2).astype(np.float64)
        cc = c[..., 0] + 1j * c[..., 1]
        o = np.einsum("bik,bkp->bip", cc, vc)
        ov = f32(out, B * n * 2 * P).reshape(B, n, 2, P)
        ov[:, :, 0], ov[:, :, 1] = o.real, o.imag
        return 0


def install(monkeypatch):
    """Route generative_audio_b200's C-ABI calls to the emulation and let its wrappers accept CPU tensors.  Returns the lib."""
    import torch

    import generative_audio_b200 as g
    lib = EmulatedLib(g._lib.load())
    monkeypatch.setattr(g._lib, "load", lambda: lib)

    def chk(*ts):
        for t in ts:
            if t is not None and not t.is_contiguous():
                raise RuntimeError("generative-audio_b200 ops need contiguous tensors")

    monkeypatch.setattr(g.ops, "_chk", chk)
    monkeypatch.setattr(g.ops, "_stream", lambda: 0)
    monkeypatch.setattr(g.inpainting.UNet, "forward", g.inpainting.UNet._forward)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    return lib
