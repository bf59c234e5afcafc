"""TEST INFRASTRUCTURE ONLY — a host-memory emulation of a handful of C-ABI entry points of libnppc_b200.so, written from the
CONTRACT in include/nppc_b200.h (argument order, layouts, the SampleScratch the Gram-Schmidt kernels leave behind), in numpy on
CPU tensors' memory.  Purpose: run the product's REAL `ops.py` wrappers (pointer marshalling, scratch decoding, reshapes) and
the autograd Functions above them without a GPU, so that code written while no GPU was available is exercised end to end
against an independent restatement of what the kernels promise.  Never imported by the product; nothing here is timed.

EmulatedLib: nppc_mask_blend, nppc_logmag_stats / _apply, nppc_gram_schmidt_complex / _real, nppc_gs_loss_fused_real,
nppc_complex_lincomb (numpy restatements incl. the scratch bytes).  EmulatedLibFull adds the STFT / iSTFT / cRM / cIRM / norm /
unfold / drop_band / complex loss entry points, each delegating to the oracle function the header cites.  Host-only entry points
(nppc_gs_scratch_bytes, ...) go to the real library."""
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


def _solve(G, n, nv, has_err, cplx, eps_add, do_gs=True):
    """The coefficient-space MGS of gs_solve (reference recurrences incl. the conjugated coefficient) -> A [n,n], stats."""
    A = np.zeros((n, n), dtype=np.complex128)
    ahat, v = [], []
    stats = dict(err_proj=np.zeros(n, np.complex128), w_norms=np.zeros(n), second=np.zeros(n))
    eps_n = np.sqrt(max(G[n, n].real, 0.0)) if has_err else 0.0
    acc = 0.0
    for i in range(n):
        a = np.zeros(n, np.complex128)
        a[i] = 1.0
        for j in range(i if do_gs else 0):
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


def _run_gs(x_ptr, gt_ptr, pred_ptr, B, n, P, scr_ptr, out_ptr, cplx, outs=None, do_gs=True):
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
        A, st = _solve(G, n, nv, has_err, cplx, 1e-8 if cplx else 1e-6, do_gs)
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


# ---- front / back end entry points, computed by the oracle (oracle/nppc_oracle.py) on views of the caller's memory -------------
def _t(ptr, shape):
    import torch
    n = int(np.prod(shape))
    return torch.from_numpy(f32(ptr, n).reshape(shape))


class EmulatedLibFull(EmulatedLib):
    """EmulatedLib + the elementwise / index / transform entry points of the hot path, each delegating to the oracle function
    that the header names as its reference (`include/nppc_b200.h`).  Used by tests/test_ops_contract_cpu.py."""

    def nppc_stft_mri(self, wave, B, L, n_fft, hop, mag, real, imag, stream):
        import nppc_oracle as O
        self.calls.append("nppc_stft_mri")
        F, T = n_fft // 2 + 1, 1 + L // hop
        m, r, i = O.stft_mri(_t(wave, (B, L)), n_fft, hop, n_fft)
        _t(mag, (B, F, T))[:], _t(real, (B, F, T))[:], _t(imag, (B, F, T))[:] = m[:, 0], r[:, 0], i[:, 0]
        return 0

    def nppc_istft(self, real, imag, B, T, n_fft, hop, length, wave, stream):
        import nppc_oracle as O
        self.calls.append("nppc_istft")
        F = n_fft // 2 + 1
        _t(wave, (B, length))[:] = O.istft(_t(real, (B, F, T)), _t(imag, (B, F, T)), length, n_fft, hop)
        return 0

    def nppc_crm_decompress_apply(self, crm, real, imag, B, FT, conj, out_mag, out_real, out_imag, stream):
        import nppc_oracle as O
        self.calls.append("nppc_crm_decompress_apply")
        m = O.decompress_cirm(_t(crm, (B, 2, FT)))
        mag, er, ei = O.crm_apply(m[:, 0], m[:, 1], _t(real, (B, FT)), _t(imag, (B, FT)), bool(conj))
        if out_mag:
            _t(out_mag, (B, FT))[:] = mag
        _t(out_real, (B, FT))[:], _t(out_imag, (B, FT))[:] = er, ei
        return 0

    def nppc_decompress_cirm(self, m, n, out, stream):
        import nppc_oracle as O
        self.calls.append("nppc_decompress_cirm")
        _t(out, (n,))[:] = O.decompress_cirm(_t(m, (n,)))
        return 0

    def nppc_build_cirm(self, nr, ni, cr, ci, B, FT, gt, stream):
        import nppc_oracle as O
        self.calls.append("nppc_build_cirm")
        g = O.build_cirm(*(_t(p, (B, FT)) for p in (nr, ni, cr, ci)))             # [B, FT, 2]
        _t(gt, (B, 2, FT))[:] = g.permute(0, 2, 1)
        return 0

    def nppc_offline_laplace_norm(self, x, B, n, sums, y, stream):
        import nppc_oracle as O
        self.calls.append("nppc_offline_laplace_norm")
        _t(y, (B, n))[:] = O.offline_laplace_norm(_t(x, (B, 1, 1, n)).clone())[:, 0, 0]
        return 0

    def nppc_pad_offline_laplace_norm(self, x, B, F, T, look_ahead, sums, y, stream):
        import torch

        import nppc_oracle as O
        self.calls.append("nppc_pad_offline_laplace_norm")
        xp = torch.nn.functional.pad(_t(x, (B, 1, F, T)), [0, look_ahead])
        _t(y, (B, F, T + look_ahead))[:] = O.offline_laplace_norm(xp)[:, 0]
        return 0

    def nppc_cumulative_laplace_norm(self, x, BC, F, T, y, stream):
        import nppc_oracle as O
        self.calls.append("nppc_cumulative_laplace_norm")
        _t(y, (BC, F, T))[:] = O.cumulative_laplace_norm(_t(x, (BC, 1, F, T)).clone())[:, 0]
        return 0

    def nppc_unfold(self, x, B, C, F, T, num_neighbor, out, stream):
        import nppc_oracle as O
        self.calls.append("nppc_unfold")
        _t(out, (B, F, C, 2 * num_neighbor + 1, T))[:] = O.unfold(_t(x, (B, C, F, T)), num_neighbor)
        return 0

    def nppc_drop_band(self, x, B, C, F, T, groups, out, stream):
        import nppc_oracle as O
        self.calls.append("nppc_drop_band")
        if not B > groups:                      # the library reports the reference's assertion (feature.py:263) as rc = -1
            self._msg = f"Batch size = {B}, num_groups = {groups}. The batch size should larger than the num_groups."
            return -1
        G = max(groups, 1)
        _t(out, (B, C, F // G, T))[:] = O.drop_band(_t(x, (B, C, F, T)), groups)
        return 0

    def nppc_last_error(self):
        return getattr(self, "_msg", "").encode()

    def nppc_gs_loss_fused(self, x, gt, pred, B, n, P, scratch, w_mat, err_norm, err_proj, w_norms, reconst, second, stream):
        self.calls.append("nppc_gs_loss_fused")
        return _run_gs(x, gt, pred, B, n, P, scratch, w_mat, True, (err_norm, err_proj, w_norms, reconst, second))

    def nppc_projection_loss(self, x, gt, pred, B, n, P, scratch, err_norm, err_proj, w_norms, reconst, second, stream):
        self.calls.append("nppc_projection_loss")
        # an explicit w_mat: no orthogonalisation (A = identity); statistics only
        return _run_gs(x, gt, pred, B, n, P, scratch, 0, True, (err_norm, err_proj, w_norms, reconst, second), do_gs=False)


def install_full(monkeypatch):
    lib = install(monkeypatch)
    full = EmulatedLibFull(lib._real)
    import generative_audio_b200 as g
    monkeypatch.setattr(g._lib, "load", lambda: full)
    return full
