"""Generates tests/golden/* by running the UNMODIFIED reference (/root/reference, imported through
oracle/ref_loader.py shims) on CPU.  Run in the dev container only:  python oracle/make_golden.py
The fixtures it writes are committed; /root/reference is never read by tests, smoke() or bench.py.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLD = os.path.join(HERE, "..", "tests", "golden")
os.makedirs(GOLD, exist_ok=True)

import ref_loader  # noqa: E402

ns = ref_loader.load_reference(trainer=True)
torch.set_grad_enabled(False)


def wave(B, L, seed, scale=0.05):
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((rng.standard_normal((B, L)) * scale).astype(np.float32))


def write_manifest():
    cfg5 = ns.MultiDirectionConfig(n_directions=5)
    cfg7 = ns.MultiDirectionConfig(n_directions=7)
    bb = ns.FullSubNet_Plus(ns.FullSubNetPlusConfig())
    h5 = ns.MultiDirectionFullSubNet_Plus(cfg5)
    h7 = ns.MultiDirectionFullSubNet_Plus(cfg7)
    entries = []
    for k, v in bb.state_dict().items():
        entries.append(("pretrained_restoration_model." + k, list(v.shape)))
    for (k, v), (k7, v7) in zip(h5.state_dict().items(), h7.state_dict().items()):
        shape = [("2*n_dirs" if a != b else a) for a, b in zip(v.shape, v7.shape)]
        entries.append(("audio_pc_wrapper.net." + k, shape))
    with open(os.path.join(GOLD, "state_dict_manifest.json"), "w") as f:
        json.dump({"source": "reference NPPCModel.state_dict() keys/shapes (nppc_audio/nppc_model.py:25-56)",
                   "entries": entries}, f)
    print("manifest:", len(entries), "tensors")


def build_model(n_dirs, groups, seed=0):
    import weights
    sd = weights.synth_state_dict(n_dirs, seed)
    tmp = tempfile.mkdtemp()
    ck = os.path.join(tmp, "bb.tar")
    torch.save({"model": {k[len("pretrained_restoration_model."):]: v for k, v in sd.items()
                          if k.startswith("pretrained_restoration_model.")}}, ck)
    cfg = ns.NPPCModelConfig(
        pretrained_restoration_model_configuration=ns.FullSubNetPlusConfig(),
        pretrained_restoration_model_path=ck,
        audio_pc_wrapper_configuration=ns.AudioPCWrapperConfig(
            multi_direction_configuration=ns.MultiDirectionConfig(n_directions=n_dirs,
                                                                  num_groups_in_drop_band=groups)),
        stft_configuration=ns.utils.StftConfig(), device="cpu")
    m = ns.NPPCModel(cfg)
    m.load_state_dict(sd, strict=True)
    m.eval()
    return m, cfg


def taps_forward(m, x):
    """Re-trace NPPCModel.forward with the reference's own functions to dump stage boundaries."""
    u = ns.utils
    mag, real, imag = u.prepare_input_from_waveform(x, 512, 256, 512, "cpu")
    crm = m.pretrained_restoration_model(mag, real, imag)
    dec = ns.decompress_cIRM(crm.permute(0, 2, 3, 1))
    emag, ereal, eimag = u.crm_to_stft_components(dec, real, imag)
    head = m.audio_pc_wrapper.net(mag, real, imag, emag.unsqueeze(1), ereal.unsqueeze(1), eimag.unsqueeze(1))
    w = m(x)
    return dict(mag=mag, real=real, imag=imag, pred_crm=crm, emag=emag, ereal=ereal, eimag=eimag, head=head, w_mat=w)


def save(name, **arrs):
    arrs = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()}
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(name, {k: v.shape for k, v in arrs.items()}, f"{os.path.getsize(path) / 1e6:.2f} MB")


def main():
    write_manifest()
    # ---- function-level fixtures -------------------------------------------------------------
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 1, 257, 11, generator=g)
    save("fn_unfold", x=x, n15=ns.BaseModel.unfold(x, 15), n0=ns.BaseModel.unfold(x, 0), n2=ns.BaseModel.unfold(x, 2))
    xd = torch.randn(5, 3, 257, 7, generator=g)
    save("fn_drop_band", x=xd, g1=ns.drop_band(xd, 1), g2=ns.drop_band(xd, 2), g3=ns.drop_band(xd, 3))
    xn = torch.rand(2, 1, 257, 19, generator=g) + 0.1
    xs = torch.randn(2, 1, 257, 19, generator=g)
    save("fn_norm", xpos=xn, xsigned=xs,
         off_pos=ns.BaseModel.offline_laplace_norm(xn), off_signed=ns.BaseModel.offline_laplace_norm(xs),
         cum_pos=ns.BaseModel.cumulative_laplace_norm(xn))
    xg = torch.randn(2, 5, 2, 33, 17, generator=g)
    xg[:, 2] = 0.7 * xg[:, 0] + 0.3 * xg[:, 2]  # some correlation between directions
    save("fn_gram_schmidt", x=xg, out=ns.gram_schmidt_to_crm(xg))
    from nppc_audio.inpainting.nppc.pc_wrapper import gram_schmidt_to_spec_mag
    xr = torch.randn(2, 10, 16, 20, generator=g)
    save("fn_gram_schmidt_real", x=xr, out=gram_schmidt_to_spec_mag(xr))
    mk = torch.randn(2, 2, 257, 9, generator=g) * 6.0
    mk[0, 0, 0, :4] = torch.tensor([9.9, -9.9, 12.0, -15.0])
    wv = wave(2, 2048, 7)
    mag, real, imag = ns.utils.prepare_input_from_waveform(wv, 512, 256, 512, "cpu")
    dec = ns.decompress_cIRM(mk.permute(0, 2, 3, 1))
    emag, ereal, eimag = ns.utils.crm_to_stft_components(dec, real, imag)
    wav_out = ns.utils.model_outputs_to_waveforms(mk, real, imag, 2048)
    wav_out_short = ns.utils.model_outputs_to_waveforms(mk, real, imag, 2000)
    cw = wave(2, 2048, 8)
    cmag, creal, cimag = ns.utils.prepare_input_from_waveform(cw, 512, 256, 512, "cpu")
    gt = ns.build_complex_ideal_ratio_mask(torch.complex(real[:, 0], imag[:, 0]), torch.complex(creal[:, 0], cimag[:, 0]))
    save("fn_stft_crm_istft", wave=wv, mag=mag, real=real.contiguous(), imag=imag.contiguous(), mask=mk, dec=dec,
         emag=emag, ereal=ereal, eimag=eimag, wav_out=wav_out, wav_out_short=wav_out_short,
         clean=cw, gt_cirm=gt, comp=ns.compress_cIRM(mk * 30))

    # ---- model-level fixtures ----------------------------------------------------------------
    m, cfg = build_model(5, 1)
    x = wave(2, 4096, 11)
    t = taps_forward(m, x)
    save("model_small_b2", wave=x, **t)

    x1 = wave(1, 64000, 12)
    t1 = taps_forward(m, x1)
    enh = ns.utils.model_outputs_to_waveforms(t1["pred_crm"], t1["real"], t1["imag"], 64000)
    save("model_full_b1", wave=x1, pred_crm=t1["pred_crm"], head=t1["head"], w_mat=t1["w_mat"], enhanced_wave=enh)

    # training-layout step: groups=2, B=4, loss at three steps (trainer.py:234-317 driven unbound)
    m2, cfg2 = build_model(5, 2)
    clean = wave(4, 4096, 13, 0.03)
    noisy = clean + wave(4, 4096, 14, 0.3 * 0.03 / 0.05 * 0.05 / 0.03)  # 0.3 * randn
    losses = {}
    for step in (0, 250, 600):
        stub = types.SimpleNamespace()
        stub.nppc_model = m2
        stub.device = "cpu"
        stub.step = step
        stub.config = types.SimpleNamespace(nppc_model_configuration=cfg2, second_moment_loss_grace=500,
                                            second_moment_loss_lambda=1.0)
        T = ns.NPPCAudioTrainer
        stub._get_true_and_pred_crm = types.MethodType(T._get_true_and_pred_crm, stub)
        stub._calculate_final_objective = types.MethodType(T._calculate_final_objective, stub)
        reconst, obj, log = T.base_step(stub, (noisy, clean))
        losses[f"objective_{step}"] = obj
        if step == 0:
            losses.update(reconst_err=reconst, w_mat=log["w_mat"], pred_crm=log["pred_crm"],
                          err_norm=log["err_norm"], err_proj_re=log["err_proj"].real, err_proj_im=log["err_proj"].imag,
                          w_norms=log["w_norms"], second_moment_mse=log["second_moment_mse"])
    save("model_step_g2_b4", noisy=noisy, clean=clean, **losses)


if __name__ == "__main__":
    main()
