#!/usr/bin/env python
"""DP training parity on real GPUs (NCCL): the bucketed, backward-overlapped gradient all-reduce of training.GradBucketReducer
gives every rank the MEAN of the per-rank objectives' gradients — checked against rank 0 recomputing every
rank's gradients by itself.  Launch:  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_train_check.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import generative_audio_b200 as G
    from generative_audio_b200 import training
    from helpers import build_model, wave
    m, _ = build_model(5, 2, "tc")
    st = G.NPPCAudioStep(m, 500, 1.0)
    st.step = 600
    params = list(m.audio_pc_wrapper.parameters())
    B, L = 4, 16000
    data = [(wave(B, L, 50 + r, 0.03), wave(B, L, 60 + r, 0.3)) for r in range(world)]
    batches = [((c + n).cuda(), c.cuda()) for c, n in data]

    def grads_of(batch, reducer=None):
        m.zero_grad(set_to_none=True)
        if reducer:
            reducer.reset()
        _, obj, _ = st.base_step(batch, requires_grad=True)
        obj.backward()
        n = reducer.finish() if reducer else 0
        return [p.grad.detach().clone() for p in params], obj.item(), n

    red = training.GradBucketReducer(params)
    g_dp, obj, ncoll = grads_of(batches[rank], red)
    red.remove()
    out = None
    if rank == 0:
        ref = None
        for r in range(world):
            g, _, _ = grads_of(batches[r])
            ref = g if ref is None else [a + b for a, b in zip(ref, g)]
        ref = [g / world for g in ref]
        num = sum(((a - b).double() ** 2).sum() for a, b in zip(g_dp, ref)).sqrt().item()
        den = sum((b.double() ** 2).sum() for b in ref).sqrt().item()
        worst = max(((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item() for a, b in zip(g_dp, ref))
        out = {"world": world, "collectives": ncoll, "buckets": len(red.buckets), "grad_rel_l2": num / den, "worst_tensor_rel_max": worst,
               "ok": bool(num / den < 1e-5)}
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    if out is not None and not out["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
