"""Backward of Gram-Schmidt (pc_wrapper.py:8-44, with its conjugated coefficient and the detach() of the normaliser at :37) +
the NPPC projection / second-moment objective (trainer.py:259-298,337-342, with the detach() at :295), in COEFFICIENT SPACE.

Every vector involved — the directions w_i, the error e = gt - pred and therefore the gradients — lies in the span of the
n + 1 vectors {x_0 .. x_{n-1}, e} the forward pass already took the Gram matrix of, so the whole backward is O(n^3) complex
arithmetic per utterance on (n+1)-vectors plus ONE streaming pass that forms d objective / d x_i = sum_k C_ik x_k + C_in e
(ops.complex_lincomb).  Pure tensor arithmetic on tiny tensors: runs on whatever device G lives on (CPU in the unit test
against torch.autograd, CUDA in the training step).

Derivation (delta = 1e-8, <u, v>_R = Re u^H v):  with nu_k = |w_k|, eps = |e|, q_k = w_k^H e,
    pi_k^2 = |q_k|^2 / ((nu_k + delta)^2 (eps + delta)^2),   nu'_k = nu_k / (eps + delta)
    objective = mean_b (1 - sum_k pi_k^2) + lambda mean_{b,k} (nu'_k^2 - sg(pi_k^2))^2
    d objective / d w_k = alpha_k e + beta_k w_k
        alpha_k = -(1/B) 2 conj(q_k) / ((nu_k + delta)^2 (eps + delta)^2)
        beta_k  =  (1/B) 2 |q_k|^2 / ((nu_k + delta)^3 nu_k (eps + delta)^2) + (lambda / (B n)) 4 (nu'_k^2 - pi_k^2) / (eps + delta)^2
Gram-Schmidt with constant (detached) a_j = w_j / |w_j|:  w_i = M_{i-1} .. M_0 x_i,  M_j v = v - a_j conj(a_j^H v).  M_j is
symmetric for <.,.>_R, so d/dx_i = M_0 M_1 .. M_{i-1} (d/dw_i).

Real variant (`real=True`: the inpainting head, inpainting/nppc/pc_wrapper.py:43-59 + trainer/nppc_trainer.py:338-385): the same
algebra on real numbers with delta = 1e-6 added to BOTH norms before any division, i.e. nu'_k = (nu_k + delta) / (eps + delta),
which changes the second-moment term of beta_k to (lambda / (B n)) 4 (nu'_k^2 - pi_k^2) (nu_k + delta) / (nu_k (eps + delta)^2)."""
import torch


def gs_loss_grad_coeffs(G: torch.Tensor, A: torch.Tensor, lam: float, grad_scale=1.0, real: bool = False) -> torch.Tensor:
    """G [B, n+1, n+1] complex (G[j,k] = v_j^H v_k over the vectors x_0..x_{n-1}, e);  A [B, n, n] complex (w_i = sum_k A_ik x_k)
    -> C [B, n, n+1] complex with d objective / d x_i = sum_k C[i,k] v_k  (times grad_scale, the incoming d/d objective).
    real=True: G, A (and C) real, the inpainting trainer's epsilon placement (module docstring)."""
    B, n1, _ = G.shape
    n = n1 - 1
    cd = G.dtype
    Wc = torch.zeros(B, n, n1, dtype=cd, device=G.device)
    Wc[:, :, :n] = A.to(cd)
    GW = torch.einsum("bmk,bjk->bjm", G, Wc)                  # (G w_j)[m]
    nu2 = torch.einsum("bjm,bjm->bj", Wc.conj(), GW).real.clamp_min(0)
    nu = nu2.sqrt()
    eps = G[:, n, n].real.clamp_min(0).sqrt()
    q = GW[:, :, n].conj()                                    # conj(e^H w_k)... = w_k^H e
    q = torch.einsum("bjm,bm->bj", Wc.conj(), G[:, :, n])     # w_k^H e, explicit
    d = 1e-6 if real else 1e-8
    e2 = (eps + d) ** 2
    pi2 = (q.abs() ** 2) / ((nu + d) ** 2 * e2[:, None])
    alpha = -(1.0 / B) * 2 * q.conj() / ((nu + d) ** 2 * e2[:, None])
    if real:
        nup2 = (nu + d) ** 2 / e2[:, None]
        sm = (lam / (B * n)) * 4 * (nup2 - pi2) * (nu + d) / (nu * e2[:, None])
    else:
        nup2 = nu2 / e2[:, None]
        sm = (lam / (B * n)) * 4 * (nup2 - pi2) / e2[:, None]
    beta = (1.0 / B) * 2 * (q.abs() ** 2) / ((nu + d) ** 3 * nu * e2[:, None]) + sm
    a = Wc / nu[:, :, None].to(cd)                            # detached normalised directions (no epsilon, pc_wrapper.py:37)
    Ga = torch.einsum("bmk,bjk->bjm", G, a)                   # (G a_j)[m];  a_j^H v = sum_m conj((G a_j)[m]) ... see below
    out = torch.zeros(B, n, n1, dtype=cd, device=G.device)
    for i in range(n):
        v = beta[:, i, None].to(cd) * Wc[:, i]
        v[:, n] = v[:, n] + alpha[:, i]
        for j in range(i - 1, -1, -1):
            s = torch.einsum("bm,bm->b", Ga[:, j].conj(), v)  # a_j^H v = sum_k conj(a_j[m]) G[m,k] v[k] = (G a_j)^H v  (G Hermitian)
            v = v - a[:, j] * s.conj()[:, None]
        out[:, i] = v
    return out * grad_scale


def gs_grad_coeffs(G2: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
    """Backward of Gram-Schmidt ALONE, for an arbitrary upstream gradient (the reference trainer's own pattern:
    `w_mat = nppc_model(noisy)`, a loss written in torch, `.backward()`, trainer.py:255-298,100-106).

    G2 [B, 2n, 2n]: Gram matrix (G2[j,k] = v_j^H v_k) of the 2n vectors (x_0 .. x_{n-1}, g_0 .. g_{n-1}), g_i = d L / d w_i;
    A [B, n, n]: w_i = sum_k A_ik x_k from the forward.  Returns C [B, n, 2n] with d L / d x_i = sum_k C[i,k] v_k.
    d/dx_i = M_0 .. M_{i-1} g_i (module docstring): every M_j moves its argument along a_j only, so the result stays in
    span{g_i, x_0 .. x_{i-1}} and the whole recursion is arithmetic on 2n-vectors.  Works for real G2 / A as well (the
    inpainting head: conj() is then a no-op)."""
    B, n2, _ = G2.shape
    n = n2 // 2
    cd = G2.dtype
    Wc = torch.zeros(B, n, n2, dtype=cd, device=G2.device)
    Wc[:, :, :n] = A.to(cd)
    GW = torch.einsum("bmk,bjk->bjm", G2, Wc)
    nu = torch.einsum("bjm,bjm->bj", Wc.conj(), GW).real.clamp_min(0).sqrt()
    a = Wc / nu[:, :, None].to(cd)                            # detached normalised directions (no epsilon, pc_wrapper.py:37)
    Ga = torch.einsum("bmk,bjk->bjm", G2, a)
    out = torch.zeros(B, n, n2, dtype=cd, device=G2.device)
    for i in range(n):
        v = torch.zeros(B, n2, dtype=cd, device=G2.device)
        v[:, n + i] = 1
        for j in range(i - 1, -1, -1):
            s = torch.einsum("bm,bm->b", Ga[:, j].conj(), v)  # a_j^H v
            v = v - a[:, j] * s.conj()[:, None]
        out[:, i] = v
    return out


def gs_coeffs_from_gram(G: torch.Tensor, n: int) -> torch.Tensor:
    """The coefficient matrix A [B, n, n] of Gram-Schmidt (w_i = sum_k A_ik x_k) replayed from the Gram matrix G [B, >=n, >=n]
    of the inputs alone, in G's precision (fp64): the same recurrences the kernel's solve runs (conjugated coefficient,
    normaliser without epsilon), so the backward need not go through the fp32 copy of A the forward leaves in its scratch."""
    B = G.shape[0]
    Gx = G[:, :n, :n]
    A = torch.zeros(B, n, n, dtype=G.dtype, device=G.device)
    ahat, v = [], []
    for i in range(n):
        a = torch.zeros(B, n, dtype=G.dtype, device=G.device)
        a[:, i] = 1
        for j in range(i):
            c = torch.einsum("bk,bk->b", a.conj(), v[j])              # sum_p conj(w[p]) what_j[p]
            a = a - ahat[j] * c[:, None]
        nrm = torch.einsum("bk,bkl,bl->b", a.conj(), Gx, a).real.clamp_min(0).sqrt()
        A[:, i] = a
        ah = a / nrm[:, None].to(G.dtype)
        ahat.append(ah)
        v.append(torch.einsum("bkl,bl->bk", Gx, ah))
    return A
