"""CPU error model of the fp16 tensor-core LSTM (lstm_tc.cu): which rounding points dominate the backbone's pred_crm error on
a speech utterance (tests/golden/speech12.npz).  Analysis tool (imports oracle/; never part of the product path).

    python tools/lstm_fp16_error_model.py [utterance index]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nppc_oracle as O  # noqa: E402
import weights  # noqa: E402

torch.set_grad_enabled(False)
torch.manual_seed(0)


def h16(x):
    return x.half().float()


def tanh_approx(x, on):
    """model of tanh.approx.f32: relative error up to 2^-11 (PTX ISA), here uniform noise of that size."""
    y = torch.tanh(x)
    if on:
        y = y * (1 + (torch.rand_like(y) * 2 - 1) * 2.0 ** -11)
    return y


def make_lstm(cfg):
    def lstm_fc(x, p, pre, fast=False):
        seq = x.transpose(1, 2).float()
        lp = f"{pre}.sequence_model"
        if cfg.get("x16", True):
            seq = h16(seq.clamp(-65504, 65504))
        for layer in (0, 1):
            w_ih, w_hh = p[f"{lp}.weight_ih_l{layer}"].float(), p[f"{lp}.weight_hh_l{layer}"].float()
            b = (p[f"{lp}.bias_ih_l{layer}"] + p[f"{lp}.bias_hh_l{layer}"]).float()
            if cfg.get("w16", True):
                w_ih, w_hh = h16(w_ih), h16(w_hh)
            N, T, _ = seq.shape
            H = w_hh.shape[1]
            zx = seq @ w_ih.T + b
            if layer == 1 and cfg.get("zx16", True):
                zx = h16(zx)
            h = torch.zeros(N, H)
            c = torch.zeros(N, H)
            outs = []
            ta = cfg.get("tanh", True)
            for t in range(T):
                z = zx[:, t] + h @ w_hh.T
                i, f, g, o = z.split(H, dim=1)
                sig = lambda v: 0.5 * tanh_approx(0.5 * v, ta) + 0.5
                c = sig(f) * c + sig(i) * tanh_approx(g, ta)
                hn = sig(o) * tanh_approx(c, ta)
                if cfg.get("c16", True):
                    c = h16(c)
                h = h16(hn) if cfg.get("h16", True) else hn
                outs.append(h)
            seq = torch.stack(outs, dim=1)
        y = seq @ p[f"{pre}.fc_output_layer.weight"].float().T + p[f"{pre}.fc_output_layer.bias"].float()
        return y.transpose(1, 2).contiguous()
    return lstm_fc


def main():
    idx = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    g = np.load(os.path.join(ROOT, "tests", "golden", "speech12.npz"))
    x = torch.from_numpy(g["wave_i16"][idx:idx + 1].astype(np.float32) / 32768.0)
    sd = weights.synth_state_dict(5, 0)
    sd64 = {k: v.double() for k, v in sd.items()}
    ref = O.get_pred_crm(sd64, x.double(), fast=False)
    den = ref.abs().max()
    base = O.get_pred_crm(sd, x, fast=True)
    print(f"utterance {idx}: fp32 oracle vs fp64: {((base - ref).abs().max() / den).item():.2e}")
    orig = O.lstm_fc
    off = dict(x16=False, w16=False, zx16=False, c16=False, h16=False, tanh=False)
    cases = {"all (kernel today)": {}, "none (sanity)": off}
    for k in off:
        cases[f"only {k}"] = dict(off, **{k: True})
    for k in off:
        cases[f"all but {k}"] = {k: False}
    for name, cfg in cases.items():
        O.lstm_fc = make_lstm(cfg)
        y = O.get_pred_crm(sd, x, fast=False)
        e = (y - ref).abs()
        print(f"{name:22s} pred_crm max-rel {((e.max()) / den).item():.2e}  rms-rel {(e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item():.2e}")
    O.lstm_fc = orig


if __name__ == "__main__":
    main()
