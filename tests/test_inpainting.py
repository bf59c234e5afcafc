"""a16 — the inpainting variant.  CPU: the oracle restatement is pinned against outputs of the UNMODIFIED reference
(tests/golden/inpaint_model_b2.npz, made by oracle/make_golden_inpainting.py) and the product UNet keeps the
reference's state_dict keys.  GPU: product (C-ABI kernels + library convolutions) vs golden and oracle."""
import json
import os
import tempfile

import pytest
import torch

import nppc_oracle as O
import weights
from conftest import GOLD, load_golden, rel_err

torch.set_grad_enabled(False)
N_DIRS = 3


def _shapes(m):
    return [(k, tuple(v.shape)) for k, v in m.state_dict().items()]


def _params(in_ch, out_ch, tag):
    import generative_audio_b200 as g
    net = g.inpainting.UNet(g.inpainting.UNetConfig(in_channels=in_ch, out_channels=out_ch))
    return net, weights.synth_unet_state_dict(_shapes(net), 0, tag)


def test_unet_state_dict_keys_match_reference():
    net, _ = _params(1, 1, "rest.")
    with open(os.path.join(GOLD, "unet_manifest.json")) as f:
        ref = [(k, tuple(s)) for k, s in json.load(f)["entries"]]
    assert _shapes(net) == ref


def test_oracle_inpainting_vs_golden():
    g = load_golden("inpaint_model_b2")
    _, p_rest = _params(1, 1, "rest.")
    _, p_head = _params(2, N_DIRS, "head.")
    clean_n, m4, masked_n = O.inpaint_preprocess(g["clean_spec"], g["masked_spec"], g["mask"])
    assert rel_err(clean_n, g["clean_n"]) < 1e-6 and rel_err(masked_n, g["masked_n"]) < 1e-6
    taps = {}
    w = O.inpaint_forward(p_rest, p_head, masked_n, m4, taps)
    assert rel_err(taps["pred"], g["pred"]) < 1e-5
    assert rel_err(w, g["w_mat"]) < 1e-4
    for step in (0, 300, 600):
        st = O.inpaint_loss(g["w_mat"], g["clean_n"], g["pred"], step=step, grace=500, lambda0=1.0)
        for k in ("err_norm", "err_proj", "w_norms", "reconst_err", "second_moment_mse"):
            assert rel_err(st[k], g[f"s{step}_{k}"]) < 1e-5, (step, k)
        assert abs(st["objective"].item() - g[f"s{step}_objective"].item()) <= 1e-5 * abs(g[f"s{step}_objective"].item()) + 1e-7


def _product_model():
    import generative_audio_b200 as g
    I = g.inpainting
    rest, p_rest = _params(1, 1, "rest.")
    rest.load_state_dict(p_rest)
    ck = os.path.join(tempfile.mkdtemp(), "rest.pt")
    torch.save({"model_state_dict": rest.state_dict()}, ck)
    cfg = I.NPPCModelConfig(pretrained_restoration_model_configuration=I.UNetConfig(in_channels=1, out_channels=1),
                            pretrained_restoration_model_path=ck,
                            audio_pc_wrapper_configuration=I.AudioInpaintingPCWrapperConfig(
                                model_configuration=I.UNetConfig(in_channels=2, out_channels=N_DIRS), n_dirs=N_DIRS))
    m = I.NPPCModel(cfg)
    m.pc_wrapper.net.load_state_dict(weights.synth_unet_state_dict(_shapes(m.pc_wrapper.net), 0, "head."))
    return m


@pytest.mark.gpu
def test_inpainting_model_vs_golden():
    import generative_audio_b200 as g
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gd = load_golden("inpaint_model_b2")
    m = _product_model()
    clean_n, m4, masked_n = g.inpainting.preprocess_data(gd["clean_spec"].cuda(), gd["masked_spec"].cuda(), gd["mask"].cuda())
    assert rel_err(clean_n.cpu(), gd["clean_n"]) < 1e-5 and rel_err(masked_n.cpu(), gd["masked_n"]) < 1e-5
    pred = m.get_pred_spec_mag_norm(masked_n, m4)
    assert rel_err(pred.cpu(), gd["pred"]) < 1e-4
    w = m(masked_n, m4)
    assert w.shape == (2, N_DIRS, 40, 50)
    assert rel_err(w.cpu(), gd["w_mat"]) < 1e-3
    # PCs are supported on the gap only
    assert torch.all(w[:, :, :, :11] == 0)
    for step in (0, 300, 600):
        stepper = g.inpainting.InpaintingNPPCStep(m, 1.0, 500)
        stepper.step = step
        reconst, objective, log = stepper.base_step((gd["masked_spec"], gd["mask"], gd["clean_spec"]))
        for k in ("err_norm", "err_proj", "w_norms", "reconst_err", "second_moment_mse"):
            assert rel_err(log[k].cpu(), gd[f"s{step}_{k}"]) < 2e-3, (step, k)
        assert abs(objective.item() - gd[f"s{step}_objective"].item()) < 2e-3 * abs(gd[f"s{step}_objective"].item()) + 1e-6


@pytest.mark.gpu
def test_inpainting_kernels_vs_oracle():
    import generative_audio_b200 as g
    rng = torch.Generator().manual_seed(5)
    B, n, Fq, T = 3, 4, 33, 47
    head = torch.randn(B, n, Fq, T, generator=rng)
    gt, pred = torch.randn(B, 1, Fq, T, generator=rng), torch.randn(B, 1, Fq, T, generator=rng)
    mask = (torch.rand(B, 1, Fq, T, generator=rng) > 0.3).float()
    x_in = torch.randn(B, 2, Fq, T, generator=rng)
    assert torch.equal(g.ops.mask_blend(None, head.cuda(), mask.cuda()).cpu(), head * (1 - mask))
    ref = x_in[:, :1] * mask + head * (1 - mask)
    assert rel_err(g.ops.mask_blend(x_in.cuda(), head.cuda(), mask.cuda()).cpu(), ref) < 1e-6
    w, st = g.ops.gs_loss_fused_real(head.cuda(), gt.cuda(), pred.cuda())
    w_ref = O.gram_schmidt_real(head)
    assert rel_err(w.cpu(), w_ref) < 1e-4
    so = O.inpaint_loss(w_ref, gt, pred, step=600, grace=500, lambda0=1.0)
    for k in ("err_norm", "err_proj", "w_norms", "reconst_err", "second_moment_mse"):
        assert rel_err(st[k].cpu(), so[k]) < 1e-4, k
    st2 = g.ops.projection_loss_real(w_ref.cuda(), gt.cuda(), pred.cuda())
    for k in ("err_norm", "err_proj", "w_norms", "reconst_err", "second_moment_mse"):
        assert rel_err(st2[k].cpu(), so[k]) < 1e-4, k
    spec_c, spec_m = torch.randn(B, 2, Fq, T, generator=rng), torch.randn(B, 2, Fq, T, generator=rng)
    c, mm, mean, std = g.ops.logmag_normalize(spec_c.cuda(), spec_m.cuda())
    co, _, mo = O.inpaint_preprocess(spec_c, spec_m, torch.ones(B, T))
    assert rel_err(c.cpu(), co) < 1e-5 and rel_err(mm.cpu(), mo) < 1e-5


def test_pca_batch_vs_reference_sklearn():
    """Row N4 (second half): the batched SVD replacement of the reference's per-item sklearn PCA, on the fixture's seeded samples
    (CPU tensors: torch.linalg.svd runs on either device; the product path calls it on CUDA tensors)."""
    import generative_audio_b200 as g
    from helpers import pca_samples
    x = pca_samples()
    gd = load_golden("fn_pca_batch")
    pcs, scaled, weights, mean, svals = g.inpainting.pca_batch(x, 5)
    assert rel_err(svals, gd["svals"]) < 1e-4 and rel_err(weights, gd["weights"]) < 1e-4 and rel_err(mean, gd["mean"]) < 1e-5
    assert rel_err(pcs, gd["pcs"]) < 1e-3 and rel_err(scaled, gd["scaled"]) < 1e-3


@pytest.mark.gpu
def test_mc_dropout_baseline():
    import generative_audio_b200 as g
    I = g.inpainting
    gd = load_golden("inpaint_model_b2")
    net = I.UNet(I.UNetConfig(in_channels=1, out_channels=1, dropout=0.2))
    net.load_state_dict(weights.synth_unet_state_dict(_shapes(net), 0, "rest."))
    model = I.RestorationWrapper(net.cuda().eval())
    clean_n, m4, masked_n = I.preprocess_data(gd["clean_spec"].cuda(), gd["masked_spec"].cuda(), gd["mask"].cuda())
    det = model(masked_n, m4)
    assert rel_err(det.cpu(), gd["pred"]) < 1e-4          # dropout layers exist but are off: same as the p = 0 reference run
    out = I.calculate_unet_baseline(model, masked_n, m4, n_mc_samples=24, n_components=4)
    B, _, Fq, T = masked_n.shape
    pcs = out["principal_components"]
    assert pcs.shape == (B, 4, Fq, T) and out["mean_prediction"].shape == (B, 1, Fq, T)
    gap = (m4 == 0)
    assert torch.all(pcs[~gap.expand_as(pcs)] == 0) and torch.all(out["mean_prediction"][~gap] == 0)
    gram = torch.einsum("bif,bjf->bij", pcs.flatten(2), pcs.flatten(2))
    assert rel_err(gram.cpu(), torch.eye(4).expand(B, 4, 4)) < 1e-4                  # orthonormal directions
    assert torch.all(out["singular_vals"][:, :-1] >= out["singular_vals"][:, 1:]) and torch.all(out["singular_vals"] > 0)
    assert rel_err(out["importance_weights"].sum(dim=1).cpu(), torch.ones(B)) < 1e-5
    # the MC mean stays close to the deterministic restoration inside the gap, and dropout is switched off again afterwards
    assert rel_err(out["mean_prediction"].cpu(), (det * gap).cpu()) < 0.5
    assert torch.equal(model(masked_n, m4), det)


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,C0,C1,Cout", [(2, 8, 31, 64, 0, 64), (1, 40, 50, 3, 0, 128), (2, 16, 62, 128, 64, 256), (1, 9, 17, 64, 64, 64)])
def test_conv3x3_tcgen05_vs_conv2d(B, H, W, C0, C1, Cout):
    """Row N4: the implicit-GEMM 3x3 convolution (TMA halo by out-of-bounds zero fill, two-tensor K loop for the decoder's
    cat) against F.conv2d in fp64 on the fp16-rounded operands; partial border tiles, channel padding, both BN tile widths."""
    import generative_audio_b200 as g
    ops = g.ops
    gen = torch.Generator().manual_seed(B * 1000 + W)
    x = torch.randn(B, C0 + C1, H, W, generator=gen)
    w = torch.randn(Cout, C0 + C1, 3, 3, generator=gen) / (3 * (C0 + C1) ** 0.5)
    bias = torch.randn(Cout, generator=gen) * 0.1
    C0p, C1p = -(-C0 // 64) * 64, -(-C1 // 64) * 64
    x0 = ops.nchw_to_nhwc_f16(x[:, :C0].contiguous().cuda(), C0p)
    x1 = ops.nchw_to_nhwc_f16(x[:, C0:].contiguous().cuda(), C1p) if C1 else None
    y = ops.conv3x3_tc(x0, x1, ops.conv3x3_pack_weights(w.cuda(), C0, C1), bias.cuda(), 0.2)
    assert y.shape == (B, H, W, Cout) and y.dtype == torch.float16
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x.half().double(), w.half().double(), bias.double(), padding=1), 0.2)
    assert rel_err(y.float().cpu().permute(0, 3, 1, 2), ref) < 2e-3


@pytest.mark.gpu
def test_inpainting_unet_on_tcgen05_matches_fp32_path():
    """set_compute_dtype(model, "tc"): every 3x3 convolution of both UNets on the in-house tcgen05 kernel (no cuDNN conv in
    the launch list); w_mat within 5e-3 of the fp32 golden."""
    import generative_audio_b200 as g
    from torch.profiler import ProfilerActivity, profile
    gd = load_golden("inpaint_model_b2")
    m = _product_model()
    _, m4, masked_n = g.inpainting.preprocess_data(gd["clean_spec"].cuda(), gd["masked_spec"].cuda(), gd["mask"].cuda())
    g.inpainting.set_compute_dtype(m, "tc")
    try:
        w = m(masked_n, m4)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            w = m(masked_n, m4)
            torch.cuda.synchronize()
    finally:
        g.inpainting.set_compute_dtype(m, None)
    assert rel_err(w.cpu(), gd["w_mat"]) < 5e-3
    names = " ".join(e.key for e in prof.key_averages())
    assert "conv3x3_tc_kernel" in names
    lib = " ".join(n for n in names.split() if "nppc" not in n and "anonymous" not in n).lower()
    assert "cudnn" not in lib and "implicit_convolve" not in lib and "conv2d" not in lib


@pytest.mark.gpu
def test_nhwc_pool_and_upsample_kernels_vs_torch():
    """The NHWC fp16 glue of the tcgen05 UNet path: MaxPool2d(2) (bit-exact) and bilinear x2 (align_corners=True) + the pad to
    the skip tensor's size (tmp_utils.py:59-82), incl. odd sizes (125 -> 62 columns, 62 -> 124 padded to 125)."""
    import generative_audio_b200 as g
    F = torch.nn.functional
    x = torch.randn(2, 64, 33, 125, generator=torch.Generator().manual_seed(3)).cuda().half()
    xh = x.permute(0, 2, 3, 1).contiguous()
    p = g.ops.maxpool2x2_nhwc(xh)
    assert torch.equal(p.permute(0, 3, 1, 2), F.max_pool2d(x, 2))
    small = x[:, :, :16, :62].contiguous()
    u = g.ops.upsample2x_pad_nhwc(small.permute(0, 2, 3, 1).contiguous(), 33, 125)
    ref = F.interpolate(small.float(), scale_factor=2, mode="bilinear", align_corners=True)
    dy, dx = 33 - ref.size(2), 125 - ref.size(3)
    ref = F.pad(ref, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
    assert rel_err(u.float().permute(0, 3, 1, 2).cpu(), ref.cpu()) < 1e-3


# ---- the training step (a16 grad path): PC head in train mode, hand-written masking / Gram-Schmidt / objective backward ----
PICKS = {"outc.conv.weight": 1, "outc.conv.bias": 1, "inc.conv.conv.0.weight": 1, "inc.conv.conv.1.weight": 1,
         "inc.conv.conv.4.bias": 1, "down4.mpconv.1.conv.3.weight": 4099, "down4.mpconv.1.conv.4.weight": 1,
         "up1.conv.conv.0.weight": 8191, "up4.conv.conv.4.bias": 1}        # == oracle/make_golden_inpainting_grads.py


def _check_step_against_reference(net, gg, step, objective, grad_norm, before, tol):
    """Gradients (taken BEFORE clipping: `before` holds them), objective, clip norm, BatchNorm running mean and the parameters
    after the clipped Adam step against the unmodified reference's training step."""
    params = dict(net.named_parameters())
    assert abs(float(objective) - gg[f"s{step}_objective"].item()) < tol * abs(gg[f"s{step}_objective"].item())
    assert abs(float(grad_norm) - gg[f"s{step}_grad_norm"].item()) < tol * gg[f"s{step}_grad_norm"].item()
    for k, s in PICKS.items():
        ref = gg[f"s{step}_grad_{k}"]
        assert rel_err(before[k].flatten()[::s].cpu(), ref) < tol, (step, k)
        # Adam's first step moves every weight by lr * g / (|g| + eps): compare the UPDATE where the gradient is not noise
        w0 = weights.synth_unet_tensor("head." + k, params[k].shape, 0).flatten()[::s]
        upd_ref = gg[f"s{step}_after_{k}"] - w0
        upd = params[k].detach().flatten()[::s].cpu() - w0
        sel = ref.abs() > 0.05 * ref.abs().max()          # well above any gradient noise: no sign flips, |g| >> Adam's eps
        assert sel.any() and (upd[sel] - upd_ref[sel]).abs().max().item() < 0.02 * 1e-4, (step, k)
    rm = dict(net.named_buffers())["inc.conv.conv.1.running_mean"]
    assert rel_err(rm.cpu(), gg[f"s{step}_bn_running_mean"]) < max(tol, 1e-5)


@pytest.mark.parametrize("step", [0, 600])
def test_inpainting_training_step_host_math_vs_reference_gradients(step):
    """CPU: everything of InpaintingNPPCStep.train_step that is NOT a kernel — the product UNet in train mode (module tree,
    BatchNorm batch statistics, state_dict naming), the coefficient-space backward of masking + real Gram-Schmidt + objective
    (gs_backward.gs_loss_grad_coeffs(real=True)) fed with the Gram / coefficient matrices the forward kernel leaves in its
    scratch (restated here in fp64), and the sync-free clip_grad_norm_ — against ONE training step of the unmodified reference
    (tests/golden/inpaint_step_b2_grads.npz)."""
    import generative_audio_b200 as g
    from generative_audio_b200.gs_backward import gs_loss_grad_coeffs
    from generative_audio_b200.inpainting import second_moment_lambda
    from generative_audio_b200.inpainting_training import clip_grad_norm_
    gd, gg = load_golden("inpaint_model_b2"), load_golden("inpaint_step_b2_grads")
    net, p_head = _params(2, N_DIRS, "head.")
    net.load_state_dict(p_head)
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.5, 0.999))
    m4 = gd["mask"][:, None, None, :].expand(-1, 1, gd["clean_n"].shape[2], -1)
    with torch.enable_grad():
        head = net._forward(torch.cat((gd["masked_n"], gd["pred"]), dim=1)) * (1 - m4)
        X = head.detach().double().flatten(2)
        V = torch.cat([X, (gd["clean_n"] - gd["pred"]).double().flatten(1)[:, None]], 1)
        G = torch.einsum("bjp,bkp->bjk", V, V)
        W = O.gram_schmidt_real(head.detach().double()).flatten(2)
        # w = A x with A lower triangular (the gap holds 9 x 40 bins: the three directions are independent)
        A = torch.linalg.lstsq(X.transpose(1, 2), W.transpose(1, 2)).solution.transpose(1, 2)
        lam = second_moment_lambda(step, 500, 1.0)
        coef = gs_loss_grad_coeffs(G, A, lam, real=True)
        head.backward(torch.einsum("bik,bkp->bip", coef, V).view_as(head).float())
    st = O.inpaint_loss(W.view_as(head).float(), gd["clean_n"], gd["pred"], step=step, grace=500, lambda0=1.0)
    before = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    norm = clip_grad_norm_(list(net.parameters()), 1.0)
    opt.step()
    _check_step_against_reference(net, gg, step, st["objective"], norm, before, 2e-4)
    assert not gg[f"s{step}_rest_has_grad"].item()


def _cpu_kernel_stubs(monkeypatch):
    """CPU restatements (oracle functions, fp64 Gram / coefficient matrices) of the kernels InpaintingNPPCStep calls, installed
    over generative_audio_b200.ops, so that the Python glue of the training step (autograd Functions, argument order, shapes,
    non-differentiable outputs, log keys, optimizer plumbing) can be rehearsed end to end without a GPU.  TEST ONLY."""
    import generative_audio_b200 as g
    ops, I = g.ops, g.inpainting

    def logmag_normalize(clean_spec, masked_spec):
        c, _, m = O.inpaint_preprocess(clean_spec, masked_spec, torch.ones(clean_spec.shape[0], clean_spec.shape[3]))
        return c, m, None, None

    def mask_blend(x_in, x, mask):
        mask = mask.reshape(x.shape[0], 1, *x.shape[2:])
        return x * (1 - mask) if x_in is None else x_in[:, :1] * mask + x * (1 - mask)

    def with_gram(head, gt, pred):
        X = head.double().flatten(2)
        V = torch.cat([X, (gt - pred).double().flatten(1)[:, None]], 1)
        w = O.gram_schmidt_real(head.double())
        A = torch.linalg.lstsq(X.transpose(1, 2), w.flatten(2).transpose(1, 2)).solution.transpose(1, 2)
        st = O.inpaint_loss(w.float(), gt, pred, step=600, grace=500, lambda0=1.0)
        st = {k: st[k] for k in ("err_norm", "err_proj", "w_norms", "reconst_err", "second_moment_mse")}
        return w.float(), st, torch.einsum("bjp,bkp->bjk", V, V), A.float()

    def real_lincomb(x, gt, pred, coef):
        V = torch.cat([x.double().flatten(2), (gt - pred).double().flatten(1)[:, None]], 1)
        return torch.einsum("bik,bkp->bip", coef.double(), V).view_as(x).float()

    setattr_ = monkeypatch.setattr if monkeypatch is not None else setattr     # None: a spawned worker process, nothing to undo
    setattr_(ops, "logmag_normalize", logmag_normalize)
    setattr_(ops, "mask_blend", mask_blend)
    setattr_(ops, "gs_loss_fused_real_with_gram", with_gram)
    setattr_(ops, "gs_loss_fused_real", lambda h, a, b: with_gram(h, a, b)[:2])
    setattr_(ops, "real_lincomb", real_lincomb)
    setattr_(I.UNet, "forward", I.UNet._forward)
    setattr_(torch.Tensor, "cuda", lambda self, *a, **k: self)


def _cpu_product_model():
    import generative_audio_b200 as g
    I = g.inpainting
    rest, p_rest = _params(1, 1, "rest.")
    rest.load_state_dict(p_rest)
    m = I.NPPCModel.__new__(I.NPPCModel)          # NPPCModel.__init__ refuses to build without CUDA: assemble its parts
    torch.nn.Module.__init__(m)
    m.pretrained_restoration_model = I.RestorationWrapper(rest).eval()
    m.pc_wrapper = I.AudioInpaintingPCWrapper(I.AudioInpaintingPCWrapperConfig(
        model_configuration=I.UNetConfig(in_channels=2, out_channels=N_DIRS), n_dirs=N_DIRS))
    m.pc_wrapper.net.load_state_dict(weights.synth_unet_state_dict(_shapes(m.pc_wrapper.net), 0, "head."))
    return m.eval()


@pytest.mark.parametrize("step", [0, 600])
def test_inpainting_train_step_glue_rehearsal_on_cpu(monkeypatch, step):
    """InpaintingNPPCStep.train_step end to end with the kernels replaced by CPU restatements: the product's own autograd
    Functions, mode handling, clipping and optimizer plumbing reproduce the reference's training step (same checks as the GPU
    test in test_zz_inpainting_training_gpu.py)."""
    import generative_audio_b200 as g
    _cpu_kernel_stubs(monkeypatch)
    gd, gg = load_golden("inpaint_model_b2"), load_golden("inpaint_step_b2_grads")
    batch = (gd["masked_spec"], gd["mask"], gd["clean_spec"])
    m = _cpu_product_model()
    m.pc_wrapper.train()
    stepper = g.inpainting.InpaintingNPPCStep(m, 1.0, 500)
    stepper.step = step
    with torch.enable_grad():
        reconst, objective, log = stepper.base_step(batch, requires_grad=True)
        objective.backward()
    assert set(log) == {"w_mat", "err_norm", "err_proj", "w_norms", "reconst_err", "second_moment_mse", "objective"}
    assert not log["w_mat"].requires_grad and rel_err(reconst.detach(), gg[f"s{step}_reconst_err"]) < 1e-4
    assert all(p.grad is None for p in m.pretrained_restoration_model.parameters())
    before = {k: p.grad.detach().clone() for k, p in m.pc_wrapper.net.named_parameters()}
    m2 = _cpu_product_model()
    stepper2 = g.inpainting.InpaintingNPPCStep(m2, 1.0, 500, max_grad_norm=1.0)
    stepper2.step = step
    opt = torch.optim.Adam(m2.parameters(), lr=1e-4, betas=(0.5, 0.999))
    with torch.enable_grad():
        obj2, log2 = stepper2.train_step(batch, opt)
    assert stepper2.step == step + 1 and m2.pc_wrapper.training and not m2.pretrained_restoration_model.training
    # 2e-3: the restoration output comes from the product's BatchNorm-folded convolutions here (1e-5 off the reference's
    # unfolded ones) and the head's train-mode BatchNorm over 2 x 3 bottleneck pixels amplifies that in the gradients
    _check_step_against_reference(m2.pc_wrapper.net, gg, step, obj2.item(), log2["grad_norm"].item(), before, 2e-3)
    with torch.no_grad():
        _, obj3, _ = stepper.base_step(batch)
    assert abs(obj3.item() - objective.item()) < 1e-4 * abs(objective.item())


def test_inpainting_eval_forward_glue_rehearsal_on_cpu(monkeypatch):
    """The eval-mode forward (folded BatchNorm, fp32 path) and the no-grad base_step with the kernels replaced by CPU
    restatements: the Python side of the GPU-verified inference path keeps reproducing the reference golden after the
    train-mode additions."""
    import generative_audio_b200 as g
    _cpu_kernel_stubs(monkeypatch)
    monkeypatch.setattr(g.ops, "gram_schmidt_real", lambda x: O.gram_schmidt_real(x))
    gd = load_golden("inpaint_model_b2")
    m = _cpu_product_model()
    assert not m.training
    clean_n, m4, masked_n = g.inpainting.preprocess_data(gd["clean_spec"], gd["masked_spec"], gd["mask"])
    assert rel_err(m.get_pred_spec_mag_norm(masked_n, m4), gd["pred"]) < 1e-4
    assert rel_err(m(masked_n, m4), gd["w_mat"]) < 1e-3
    stepper = g.inpainting.InpaintingNPPCStep(m, 1.0, 500)
    stepper.step = 600
    _, objective, log = stepper.base_step((gd["masked_spec"], gd["mask"], gd["clean_spec"]))
    assert abs(objective.item() - gd["s600_objective"].item()) < 2e-3 * abs(gd["s600_objective"].item())
    monkeypatch.undo()                                  # without the test-only stubs the product refuses CPU tensors
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        m.pc_wrapper.net(masked_n)


def test_checkpoint_round_trip_in_the_reference_format(monkeypatch, tmp_path):
    """save_checkpoint writes the reference trainers' file ({'model_state_dict', 'optimizer_state_dict', 'step'},
    nppc_trainer.py:604-618 / trainer.py:319-335); load_checkpoint restores model, Adam moments and the step counter, so that a
    resumed run continues bit-identically (CPU rehearsal with stubbed kernels)."""
    import generative_audio_b200 as g
    _cpu_kernel_stubs(monkeypatch)
    gd = load_golden("inpaint_model_b2")
    batch = (gd["masked_spec"], gd["mask"], gd["clean_spec"])

    def fresh():
        m = _cpu_product_model()
        return m, g.inpainting.InpaintingNPPCStep(m, 1.0, 500), torch.optim.Adam(m.parameters(), lr=1e-4, betas=(0.5, 0.999))

    m1, s1, o1 = fresh()
    s1.step = 7
    with torch.enable_grad():
        s1.train_step(batch, o1)
    path = s1.save_checkpoint(str(tmp_path / "ck" / "checkpoint_final.pt"), o1)
    ck = torch.load(path, weights_only=True)
    assert set(ck) == {"model_state_dict", "optimizer_state_dict", "step"} and ck["step"] == 8
    with open(os.path.join(GOLD, "unet_manifest.json")) as f:
        unet_keys = [k for k, _ in json.load(f)["entries"]]
    assert list(ck["model_state_dict"]) == ["pretrained_restoration_model.net." + k for k in unet_keys] + \
        ["pc_wrapper.net." + k for k in unet_keys]                     # the reference NPPCModel's keys: its validator loads them strictly
    m2, s2, o2 = fresh()
    assert s2.load_checkpoint(path, o2, map_location="cpu") == 8
    for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    with torch.enable_grad():
        obj1, _ = s1.train_step(batch, o1)
        obj2, _ = s2.train_step(batch, o2)
    assert obj1.item() == obj2.item() and s1.step == s2.step == 9
    for a, b in zip(m1.pc_wrapper.parameters(), m2.pc_wrapper.parameters()):
        assert torch.equal(a, b)                                       # same Adam moments -> the same second update


def _inpaint_dp_worker(rank, world, port, ret):
    import torch.distributed as dist

    import generative_audio_b200 as g
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    _cpu_kernel_stubs(None)
    gd = load_golden("inpaint_model_b2")
    gen = torch.Generator().manual_seed(5)
    batches = []
    for r in range(world):                                   # rank r trains on its own batch: the golden one, perturbed
        clean = gd["clean_spec"] + 0.3 * r * torch.randn(gd["clean_spec"].shape, generator=gen)
        batches.append((clean * gd["mask"][:, None, None, :], gd["mask"], clean))

    def grads_of(batch):
        m = _cpu_product_model()
        m.pc_wrapper.train()
        st = g.inpainting.InpaintingNPPCStep(m, 1.0, 500)
        st.step = 600
        with torch.enable_grad():
            _, obj, _ = st.base_step(batch, requires_grad=True)
            obj.backward()
        return [p.grad.clone() for p in m.pc_wrapper.parameters()]

    ref = None
    if rank == 0:                                            # single-process reference: mean of the per-rank gradients
        per = [grads_of(b) for b in batches]
        ref = [sum(gs) / world for gs in zip(*per)]
    m = _cpu_product_model()
    st = g.inpainting.InpaintingNPPCStep(m, 1.0, 500, max_grad_norm=1e9)      # no clipping: compare raw averaged gradients
    st.step = 600
    opt = torch.optim.SGD(m.pc_wrapper.parameters(), lr=0.0)
    with torch.enable_grad():
        _, log = st.train_step(batches[rank], opt)
    if rank == 0:
        errs = [((p.grad - r_).abs().max() / r_.abs().max().clamp_min(1e-12)).item() for p, r_ in zip(m.pc_wrapper.parameters(), ref)
                if r_.abs().max() > 1e-6]
        ret["max_err"], ret["buckets"] = max(errs), len(st._reducer.buckets)
        ret["norm_ok"] = abs(log["grad_norm"].item() - torch.sqrt(sum((r_ ** 2).sum() for r_ in ref)).item()) < 1e-4 * log["grad_norm"].item()
    dist.destroy_process_group()


def test_inpainting_train_step_data_parallel_world2():
    """InpaintingNPPCStep.train_step under torch.distributed (gloo, 2 ranks, kernels stubbed on CPU): the head's gradients
    after the overlapped bucket all-reduce are the mean of the per-rank gradients, and the clip norm is taken after averaging."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_inpaint_dp_worker, args=(2, 29671, ret), nprocs=2, join=True)
    assert ret["buckets"] >= 2 and ret["max_err"] < 1e-5 and ret["norm_ok"]


@pytest.mark.parametrize("step", [0, 600])
def test_inpainting_train_step_through_the_real_ops_wrappers_on_an_emulated_cabi(monkeypatch, step):
    """Same training step as the rehearsal above, but through the product's REAL ops.py wrappers: only the shared library
    is replaced, by tests/cabi_emulator.py (numpy on host memory, written from the contract in include/nppc_b200.h).  Covers
    what the rehearsal's stubs skipped — pointer / size marshalling, the SampleScratch decode of gs_loss_fused_real_with_gram,
    real_lincomb riding on nppc_complex_lincomb's two planes, the mask kernel in forward and backward."""
    import cabi_emulator
    import generative_audio_b200 as g
    lib = cabi_emulator.install(monkeypatch)
    gd, gg = load_golden("inpaint_model_b2"), load_golden("inpaint_step_b2_grads")
    batch = (gd["masked_spec"], gd["mask"], gd["clean_spec"])
    m = _cpu_product_model()
    # inference path first (eval head): preprocess, restoration, head, real Gram-Schmidt, fused loss statistics
    clean_n, m4, masked_n = g.inpainting.preprocess_data(gd["clean_spec"], gd["masked_spec"], gd["mask"])
    assert rel_err(clean_n, gd["clean_n"]) < 1e-5 and rel_err(masked_n, gd["masked_n"]) < 1e-5
    assert rel_err(m(masked_n, m4), gd["w_mat"]) < 1e-3
    st0 = g.inpainting.InpaintingNPPCStep(m, 1.0, 500)
    st0.step = step
    _, obj0, log0 = st0.base_step(batch)
    assert abs(obj0.item() - gd[f"s{step}_objective"].item()) < 2e-3 * abs(gd[f"s{step}_objective"].item()) if f"s{step}_objective" in gd else True
    for k in ("err_norm", "err_proj", "w_norms", "reconst_err", "second_moment_mse"):
        if f"s{step}_{k}" in gd:
            assert rel_err(log0[k], gd[f"s{step}_{k}"]) < 2e-3, k
    # training step: gradients before clipping, then the whole iteration on a fresh model
    m.pc_wrapper.train()
    with torch.enable_grad():
        reconst, objective, _ = st0.base_step(batch, requires_grad=True)
        objective.backward()
    before = {k: p.grad.detach().clone() for k, p in m.pc_wrapper.net.named_parameters()}
    m2 = _cpu_product_model()
    st2 = g.inpainting.InpaintingNPPCStep(m2, 1.0, 500, max_grad_norm=1.0)
    st2.step = step
    opt = torch.optim.Adam(m2.parameters(), lr=1e-4, betas=(0.5, 0.999))
    lib.calls.clear()
    with torch.enable_grad():
        obj2, log2 = st2.train_step(batch, opt)
    # 5e-3: fp32 pre-processing, fp32 coefficient matrices out of the scratch and the fp32 linear combination put ~3e-5 of
    # noise on the head's output gradient, which the train-mode BatchNorm over the 2 x 3 bottleneck pixels of this tiny fixture
    # amplifies ~100 x (measured 2.7e-3 on inc.conv.conv.0.weight); the GPU test allows 2e-2
    _check_step_against_reference(m2.pc_wrapper.net, gg, step, obj2.item(), log2["grad_norm"].item(), before, 5e-3)
    assert lib.calls == ["nppc_logmag_stats", "nppc_logmag_apply", "nppc_logmag_apply", "nppc_mask_blend",      # frozen half
                         "nppc_mask_blend", "nppc_gs_loss_fused_real",                                         # head forward
                         "nppc_complex_lincomb", "nppc_mask_blend"]                                            # backward


def test_reference_inpainting_trainer_pattern_through_differentiable_forward(monkeypatch):
    """`model.differentiable_forward = True`: the reference trainer's own base_step pattern — w_mat = model(x, mask), the loss in
    torch (nppc_trainer.py:347-373, restated by the oracle), backward — through the product model on the emulated C ABI,
    against the unmodified reference's gradients (same fixture as the fused training step)."""
    import cabi_emulator
    import generative_audio_b200 as g
    lib = cabi_emulator.install(monkeypatch)
    gd, gg = load_golden("inpaint_model_b2"), load_golden("inpaint_step_b2_grads")
    m = _cpu_product_model()
    m.pc_wrapper.train()
    clean_n, m4, masked_n = g.inpainting.preprocess_data(gd["clean_spec"], gd["masked_spec"], gd["mask"])
    with torch.enable_grad():
        assert not m(masked_n, m4).requires_grad                      # default: the inference path, even with grad mode on
        m.differentiable_forward = True
        lib.calls.clear()
        w = m(masked_n, m4)
        assert w.requires_grad
        pred = m.get_pred_spec_mag_norm(masked_n, m4)
        st = O.inpaint_loss(w, clean_n, pred, step=600, grace=500, lambda0=1.0)
        st["objective"].backward()
    assert lib.calls == ["nppc_mask_blend", "nppc_mask_blend", "nppc_gram_schmidt_real",                    # forward
                         "nppc_mask_blend",                                                                  # get_pred (2nd restoration pass)
                         "nppc_gram_schmidt_real", "nppc_complex_lincomb", "nppc_mask_blend"]                # backward
    assert abs(st["objective"].item() - gg["s600_objective"].item()) < 2e-3 * abs(gg["s600_objective"].item())
    params = dict(m.pc_wrapper.net.named_parameters())
    for k, s in PICKS.items():
        assert rel_err(params[k].grad.flatten()[::s], gg[f"s600_grad_{k}"]) < 5e-3, k
    assert all(p.grad is None for p in m.pretrained_restoration_model.parameters())


def test_inpainting_validate_matches_the_reference_statistics_and_restores_modes(monkeypatch):
    """InpaintingNPPCStep.validate (nppc_trainer.py:689-706) on the emulated C ABI: eval-mode statistics equal the reference's
    base_step fixture; the head returns to train mode, the frozen restoration UNet never leaves eval (the reference's
    `self.nppc_model.train()` would flip it — deliberately not reproduced)."""
    import cabi_emulator
    import generative_audio_b200 as g
    cabi_emulator.install(monkeypatch)
    gd = load_golden("inpaint_model_b2")
    batch = (gd["masked_spec"], gd["mask"], gd["clean_spec"])
    m = _cpu_product_model()
    m.pc_wrapper.train()
    st = g.inpainting.InpaintingNPPCStep(m, 1.0, 500)
    st.step = 300
    loss, err = st.validate([batch, batch])
    assert abs(loss.item() - gd["s300_objective"].item()) < 2e-3 * abs(gd["s300_objective"].item())
    assert abs(err.item() - gd["s300_reconst_err"].mean().item()) < 2e-3
    assert m.pc_wrapper.training and not m.pretrained_restoration_model.training and not any(
        mod.training for mod in m.pretrained_restoration_model.modules())
    with pytest.raises(ValueError):
        st.validate([])
