#!/usr/bin/env python
"""Error table of the product path on the realistic / ill-conditioned fixtures (GPU box):
    python tools/parity_report.py [--out gpurun_out/parity_report.json] [--cases speech12,noise_ill,noise_ill_short,model_b64]
Per fixture, implementation and output: per-utterance max-norm relative error vs the reference's fp32 run and fp64 run, the
reference's own fp32-vs-fp64 gap, and the VERDICT budget max(base, 2 x gap)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402

import parity_cases as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_report.json"))
    ap.add_argument("--cases", default="speech12,noise_ill,noise_ill_short,model_b64")
    ap.add_argument("--impls", default="f32,tc")
    args = ap.parse_args()
    report = {}
    for name in args.cases.split(","):
        for impl in args.impls.split(","):
            if name == "model_b64" and impl == "f32":   # (tcp, the split-precision tensor-core path, is fast enough)
                continue   # the fp32 SIMT LSTM at B = 64 is minutes of GPU time; tc is the benchmarked path
            r = P.run_case(name, impl)
            if "_auto" in r:
                a = r.pop("_auto")
                report[f"{name}/auto/routing"] = a
                print(f"{name:16s} auto: {len(a['rerouted'])} of {len(a['depth'])} utterances re-run through tcp (depth > {a['threshold']}): {a['rerouted']}", flush=True)
            for k, v in r.items():
                bud = P.budget(impl, v["gap"], k)
                e = np.minimum(v["err32"], v["err64"])
                report[f"{name}/{impl}/{k}"] = dict(err32=v["err32"].tolist(), err64=v["err64"].tolist(), gap=v["gap"].tolist(),
                                                    budget=bud.tolist(), within=bool((e <= bud).all()))
                print(f"{name:16s} {impl:3s} {k:14s} err32 max {v['err32'].max():.2e} med {np.median(v['err32']):.2e} | err64 max "
                      f"{v['err64'].max():.2e} | gap max {v['gap'].max():.2e} | worst err/budget {np.max(e / bud):.2f}", flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(report, f)


if __name__ == "__main__":
    main()
