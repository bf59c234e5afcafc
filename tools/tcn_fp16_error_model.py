"""CPU error model of the channel-last fp16 TCN path (tcn_cl.cu + gemm_tc.cu), used to decide WHICH quantisation points
matter on ill-conditioned inputs (tests/golden/noise_ill.npz "short_*": wave(2, 4096, 31)).  Test/analysis tool: imports
oracle/ — never part of the product path.

    python tools/tcn_fp16_error_model.py

Every switch below names one place where the GPU path rounds to fp16; the script replaces the oracle's TCN with the emulation,
runs the full NPPC forward (LSTM in fp32) and prints the max-norm relative error of w_mat against the fp64 reference run.
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
from helpers import wave  # noqa: E402

torch.set_grad_enabled(False)


def h(x):
    return x.half().float()


def split2(x):
    hi = x.half().float()
    return hi, (x - hi).half().float()


def make_tcn(cfg):
    def tcn(x, p, pre):
        B, C, T = x.shape
        x = x.float()
        scale = x.reshape(B, -1).abs().amax(dim=1).clamp_min(1e-30)[:, None, None] if cfg.get("scale", True) else torch.ones(B, 1, 1)
        x32 = x.clone()

        def A(v):   # fp16 A operand of a GEMM from the fp32 residual stream
            v = v / scale
            if cfg.get("x_split"):
                hi, lo = split2(v)
                return hi + lo
            return h(v) if cfg.get("xh", True) else v

        def W(w):
            if cfg.get("w_split"):
                hi, lo = split2(w)
                return hi + lo
            return h(w) if cfg.get("w", True) else w

        def out16(v, key):
            return h(v) if cfg.get(key, True) else v

        for i, d in enumerate(O.TCN_DILATIONS):
            q = f"{pre}.sequence_model.{i}"
            w1 = p[f"{q}.conv1x1.weight"][:, :, 0].float()
            y1 = out16(torch.einsum("oc,bct->bot", W(w1), A(x32)), "y1")
            y = y1 * scale + p[f"{q}.conv1x1.bias"].float()[None, :, None]
            y = O._groupnorm1(O._prelu(y, p[f"{q}.prelu1.weight"].float()), p[f"{q}.norm1.weight"].float(), p[f"{q}.norm1.bias"].float())
            y = torch.nn.functional.conv1d(y, p[f"{q}.depthwise_conv.weight"].float(), p[f"{q}.depthwise_conv.bias"].float(),
                                           padding=d, dilation=d, groups=y.shape[1])
            z = O._prelu(y, p[f"{q}.prelu2.weight"].float())
            mu = z.reshape(B, -1).mean(dim=1)[:, None, None]
            var = z.reshape(B, -1).var(dim=1, unbiased=False)[:, None, None]
            rstd = 1.0 / torch.sqrt(var + 1e-8)
            z16 = out16(z, "z")
            W2 = p[f"{q}.sconv.weight"][:, :, 0].double()
            g2, b2 = p[f"{q}.norm2.weight"].double(), p[f"{q}.norm2.bias"].double()
            w2f = (W2 * g2[None, :]).float()
            u = (W2 @ g2).float()
            vb = (W2 @ b2 + p[f"{q}.sconv.bias"].double()).float()
            o = out16(torch.einsum("oc,bct->bot", W(w2f), z16), "o")
            x32 = x32 + o * rstd + (vb[None, :, None] - mu * rstd * u[None, :, None])
        xr = torch.relu(x32)
        wfc = p[f"{pre}.fc_output_layer.weight"].float()
        o = out16(torch.einsum("oc,bct->bot", W(wfc), A(xr)), "ofc")
        return torch.relu(o * scale + p[f"{pre}.fc_output_layer.bias"].float()[None, :, None])
    return tcn


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "noise_ill.npz"))
    ref64 = torch.from_numpy(g["short_w_mat_f64"]).double()
    ref32 = torch.from_numpy(g["short_w_mat"]).double()
    den = ref64.reshape(2, -1).abs().amax(dim=1)
    x = wave(2, 4096, 31)
    sd = weights.synth_state_dict(5, 0)

    def err(w, ref):
        return ((w.double() - ref).reshape(2, -1).abs().amax(dim=1) / den).max().item()

    print(f"reference fp32 vs fp64 gap: {err(ref32, ref64):.3e}")
    orig = O.tcn_sequence_model
    w = O.nppc_forward(sd, x, n_dirs=5, fast=True)
    print(f"oracle fp32 (exact TCN)              vs f64 {err(w, ref64):.3e}   vs ref32 {err(w, ref32):.3e}")
    cases = {
        "all fp16 points (GPU path today)": {},
        "only xh fp16": dict(w=False, y1=False, z=False, o=False, ofc=False),
        "only weights fp16": dict(xh=False, y1=False, z=False, o=False, ofc=False),
        "only y1 fp16": dict(xh=False, w=False, z=False, o=False, ofc=False),
        "only z fp16": dict(xh=False, w=False, y1=False, o=False, ofc=False),
        "only o fp16": dict(xh=False, w=False, y1=False, z=False, ofc=False),
        "only ofc fp16": dict(xh=False, w=False, y1=False, z=False, o=False),
        "x split hi/lo, rest fp16": dict(x_split=True),
        "x split + w split, rest fp16": dict(x_split=True, w_split=True),
        "x split + w split, y1 f32": dict(x_split=True, w_split=True, y1=False),
        "x split + w split, y1+ofc f32": dict(x_split=True, w_split=True, y1=False, ofc=False),
        "x split + w split, y1+ofc+o f32": dict(x_split=True, w_split=True, y1=False, ofc=False, o=False),
    }
    for name, cfg in cases.items():
        O.tcn_sequence_model = make_tcn(cfg)
        w = O.nppc_forward(sd, x, n_dirs=5, fast=True)
        print(f"{name:38s} vs f64 {err(w, ref64):.3e}   vs ref32 {err(w, ref32):.3e}")
    O.tcn_sequence_model = orig


if __name__ == "__main__":
    main()
