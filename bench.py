#!/usr/bin/env python
"""bench.py — NPPC-audio inference throughput (audio-seconds / second, 16 kHz, 5 PCs) on N B200s.

One "step" = one pass of the hot path (NPPCModel.forward: STFT -> frozen FullSubNet+ -> cRM -> PC head ->
Gram-Schmidt) over one batch of synthetic 4 s utterances per GPU.  Prints ONE JSON line (contract in the task
statement): value = whole-job audio-s/s with inputs resident in HBM; e2e = same through the public
NPPCModel.forward call with pinned HOST input and the w_mat result copied back to the host inside the timed region;
roofline = tensor-pipe fraction of the dominant kernel (the sub-band LSTM), timed live with CUDA events;
cpu_baseline = the oracle port of the reference's CPU path on this box's host cores (bounded sample).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--lstm-impl tc|f32]
  python bench.py --impl reference ...      # the reference's CPU path (oracle port), rank 0 only
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

SR = 16000
L = 64000          # 4 s
N_DIRS = 5
F, T, TP, H = 257, 251, 253, 384
METRIC = "audio-sec/sec NPPC-audio inference (16 kHz, 5 PCs)"
# 3.645 MFLOP per (sequence, step) for the 2-layer sub-band LSTM + fc
LSTM_FLOP_PER_SEQ_STEP = 2 * 1536 * (34 + 384) + 2 * 1536 * 768


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        if "hbm_gbs" in d and "bf16_tflops" in d:
            # a file without the sustained figure: the recipe's sustained / burst ratio (1400 / 1590) of the measured burst peak
            return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"],
                        tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"] * 1400.0 / 1590.0), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if len(s) >= 6 and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) >= 6 and s[1].isdigit()]
        reasons = set()
        for s in self.samples:
            if len(s) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(steps, warmup, batch_cpu):
    """The reference's CPU path for this metric, restated by oracle/nppc_oracle.py (fast=True -> ATen's fused CPU LSTM,
    i.e. exactly what the reference's nn.LSTM runs), on all host threads."""
    import torch

    import nppc_oracle as O
    import weights
    from helpers import wave
    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = weights.synth_state_dict(N_DIRS, 0)
    x = wave(batch_cpu, L, 100)
    for _ in range(warmup):
        O.nppc_forward(sd, x, n_dirs=N_DIRS, fast=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.nppc_forward(sd, x, n_dirs=N_DIRS, fast=True)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(value=batch_cpu * L / SR / dt, ms_per_step=dt * 1e3, cores=cores,
                sample=f"{steps} x NPPCModel.forward on {batch_cpu} synthetic 4 s utterances, fp32, torch CPU, {cores} threads")


def gpu_eager_run(batch, steps=3, warmup=2):
    """SURVEY.md §0.1 / §8(d) bar: the reference's own Blackwell path = the same torch-eager code on `cuda` (fp32, cuDNN LSTM,
    cuFFT STFT, library conv/GEMM), here the oracle port with its tensors on the device (fast=True -> nn.LSTM-equivalent
    aten::lstm -> cuDNN).  Same workload (B x 4 s, n_dirs = 5), CUDA-event timed, inputs resident."""
    import torch

    import nppc_oracle as O
    import weights
    from helpers import wave
    dev = torch.device("cuda", torch.cuda.current_device())
    prev_tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False   # the reference never enables TF32: plain fp32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            sd = {k: v.to(dev) for k, v in weights.synth_state_dict(N_DIRS, 0).items()}
            x = wave(batch, L, 1000).to(dev)
            for _ in range(warmup):
                O.nppc_forward(sd, x, n_dirs=N_DIRS, fast=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                O.nppc_forward(sd, x, n_dirs=N_DIRS, fast=True)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / steps
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev_tf32
        torch.cuda.empty_cache()
    return {"value": batch * L / SR / (ms * 1e-3), "unit": "audio-s/s", "ms_per_step": ms, "batch": batch, "steps": steps,
            "warmup": warmup, "what": "reference's torch-eager path on the same B200 (oracle port on cuda: fp32, TF32 off, "
                                      "cuDNN LSTM, cuFFT, library convolutions), device-resident input"}


def sweep_bench(args, rank, world, local_rank, dev, G, build_model, wave, ClockSampler, dist):
    """BASELINE config 5 as written: a job of --utterances synthetic 4 s utterances, round-robin sharded over the ranks
    (utterance i -> rank i % N, sharding.shard_utterances), each rank walking its shard in micro-batches of --batch with a
    PARTIAL last micro-batch, pinned host waveforms in, pinned host w_mat out (forward_host).  value = all utterances' audio
    seconds / the slowest rank's device time for its whole shard; `steps` = passes over the job."""
    import torch
    from generative_audio_b200.sharding import aggregate_throughput, shard_utterances
    model, _ = build_model(N_DIRS, 1, args.lstm_impl)
    mine = shard_utterances(args.utterances, rank, world)
    B = args.batch
    batches = [mine[i:i + B] for i in range(0, len(mine), B)]
    # utterance i is wave(1, L, seed = 10_000 + i): the same job whatever the number of ranks
    hosts = [torch.cat([wave(1, L, 10_000 + i) for i in idx]).pin_memory() for idx in batches]
    outs = [torch.empty(len(idx), N_DIRS, 2, F, T, dtype=torch.float32).pin_memory() for idx in batches]

    def one_pass():
        for h, o in zip(hosts, outs):
            model.forward_host(h, o)

    def timed(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            one_pass()
        if getattr(model, "_copy_stream", None) is not None:
            torch.cuda.current_stream().wait_stream(model._copy_stream)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    if hosts:
        model.forward_host(hosts[0], outs[0])     # warm-up: plans, tables, allocator
        if len(hosts[-1]) != len(hosts[0]):
            model.forward_host(hosts[-1], outs[-1])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(args.steps) / args.steps
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if world > 1:
        dist.barrier()
    _, ms_job = aggregate_throughput(len(mine) * L / SR, ms)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    audio_s = args.utterances * L / SR
    line = {"metric": "audio-sec/sec NPPC-audio inference, 1000-utterance sweep (16 kHz, 5 PCs)", "value": audio_s / (ms_job * 1e-3),
            "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": 1, "ms_per_step": ms_job, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16" if args.lstm_impl == "tc" else "f32", "data": "synthetic", "mode": "sweep",
            "config": {"workload": f"BASELINE config 5: {args.utterances} synthetic 4 s utterances, round-robin over {world} rank(s), micro-batch "
                                   f"{B} with a partial last batch ({len(batches)} micro-batches on rank 0, last = {len(batches[-1]) if batches else 0}), "
                                   "host waveforms in / host w_mat out", "utterances": args.utterances, "batch_per_gpu": B, "n_dirs": N_DIRS,
                       "parallelism": f"utterance-sharded x{world} (no data-path collective)"},
            "e2e": {"value": audio_s / (ms_job * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_job,
                    "h2d_bytes_per_step": sum(h.numel() for h in hosts) * 4, "d2h_bytes_per_step": sum(o.numel() for o in outs) * 4},
            "clocks": sampler.summary()}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def train_bench(args, rank, world, local_rank, dev, G, build_model, wave, ClockSampler, dist):
    """BASELINE config 3 / 5: one training step of the PC head per GPU per step (frozen backbone on the inference kernels, head
    forward + hand-written backward, Adam), gradients averaged over the ranks by the bucketed all-reduce that overlaps the
    backward.  Seeds differ per rank (base + rank) and the weights are identical, as DP requires."""
    import torch
    ops = G.ops
    B = args.train_batch
    model, _ = build_model(N_DIRS, 2, "tc")
    stepper = G.NPPCAudioStep(model, 500, 1.0)
    stepper.step = 600
    # whole step as ONE CUDA graph replay (trainer.train_step_graphed).  Default: on for one GPU; with N > 1 the step stays eager
    # unless NPPC_TRAIN_GRAPH=1: the graph with the NCCL bucket all-reduces captured inside it was measured at N = 2
    # (profiles/r02_bench_train_n2.json, 99.8 ms/step vs 97.9 at N = 1) but an N = 4 run did not finish inside its time limit and
    # could not be diagnosed before the round's GPU budget ran out, so it is opt-in there.
    graphed = os.environ.get("NPPC_TRAIN_GRAPH", "1" if world == 1 else "0") != "0"
    opt = torch.optim.Adam(model.audio_pc_wrapper.parameters(), lr=1e-5, capturable=graphed)
    do_step = stepper.train_step_graphed if graphed else stepper.train_step
    clean_h = wave(B, L, 3000 + rank, 0.03).pin_memory()
    noisy_h = (clean_h + 0.3 * wave(B, L, 4000 + rank, 1.0)).pin_memory()
    clean_d, noisy_d = clean_h.to(dev), noisy_h.to(dev)
    lstm_ev = []
    of, ob = ops.lstm_step_forward, ops.lstm_step_backward

    def tf(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = of(*a, **k); e1.record()
        lstm_ev.append((e0, e1))
        return r

    def tb(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = ob(*a, **k); e1.record()
        lstm_ev.append((e0, e1))
        return r

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(n, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    obj_host = torch.zeros(1).pin_memory()

    def step_dev():
        do_step((noisy_d, clean_d), opt)

    def step_e2e():
        obj, _ = do_step((noisy_h.to(dev, non_blocking=True), clean_h.to(dev, non_blocking=True)), opt)
        obj_host.copy_(obj.reshape(1), non_blocking=True)

    run(args.warmup, step_dev)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ops.reset_launch_count()
    ms = run(args.steps, step_dev) / args.steps
    launches = ops.launch_count()
    barrier()
    # the LSTM's share: one EAGER step with CUDA events around nppc_lstm_step_forward / _backward (events cannot sit inside a
    # captured graph); the same kernels with the same arguments as in the timed steps
    ops.lstm_step_forward, ops.lstm_step_backward = tf, tb
    ops.reset_launch_count()
    stepper._graph_saved, stepper._graph = stepper._graph, None
    eager_opt = torch.optim.Adam(model.audio_pc_wrapper.parameters(), lr=1e-5)
    stepper._reducer_saved, stepper._reducer = stepper._reducer, None
    stepper._step_body((noisy_d, clean_d), eager_opt)
    torch.cuda.synchronize()
    if graphed:
        launches = ops.launch_count()     # kernels of one step (a replayed graph launches the same kernels without the counter)
    stepper._graph, stepper._reducer = stepper._graph_saved, stepper._reducer_saved
    ops.lstm_step_forward, ops.lstm_step_backward = of, ob
    barrier()
    run(1, step_e2e)
    barrier()
    ms_e2e = run(args.steps, step_e2e) / args.steps
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    from generative_audio_b200.sharding import aggregate_throughput
    _, ms_step = aggregate_throughput(B * L / SR, ms)
    _, ms_e2e_step = aggregate_throughput(B * L / SR, ms_e2e)
    per_rank = None
    if world > 1:
        mine = torch.tensor([ms, ms_e2e, float(sampler.summary()["sm_mhz"] or 0)], device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        allr = torch.stack(allr).cpu()
        per_rank = {"ms_per_step": [round(v, 3) for v in allr[:, 0].tolist()], "sm_mhz": [int(v) for v in allr[:, 2].tolist()]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    R, Tp_ = B * (F // 2), TP
    lstm_ms = sum(a.elapsed_time(b) for a, b in lstm_ev)                                  # forward + backward of ONE (eager) step
    flops = 3 * LSTM_FLOP_PER_SEQ_STEP * R * Tp_                                          # forward + 2 x for BPTT (dX and dW)
    pk = peaks()
    achieved = flops / (max(lstm_ms, 1e-9) * 1e-3) / 1e12
    audio_s = B * world * L / SR
    nparams = sum(p.numel() for p in model.audio_pc_wrapper.parameters())
    line = {"metric": "audio-sec/sec NPPC-audio PC-head training step (16 kHz, 5 PCs)", "value": audio_s / (ms_step * 1e-3), "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic", "mode": "train",
            "config": {"workload": f"BASELINE config 3: PC-head training step, {B} x 4 s utterances per GPU, n_dirs={N_DIRS}, drop_band groups=2, "
                                   "frozen FullSubNet+ (inference kernels) + head fwd/bwd (hand-written) + Adam; DP gradient all-reduce "
                                   f"({nparams} head parameters, fp32) bucketed and overlapped with the backward over NCCL when N > 1",
                       "batch_per_gpu": B, "n_dirs": N_DIRS, "parallelism": f"dp{world}"},
            "e2e": {"value": audio_s / (ms_e2e_step * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_e2e_step,
                    "h2d_bytes_per_step": 2 * noisy_h.numel() * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches) * (args.steps if graphed else 1), "cuda_graph": graphed,
            "roofline": {"bound": "tensor", "kernel": "stepwise tcgen05 LSTM forward + BPTT + weight-gradient GEMMs per training step",
                         "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
                         "traffic": None, "ms_per_step": lstm_ms, "share_of_step": lstm_ms / ms_step, "peak_source": pk["src"]},
            "clocks": sampler.summary()}
    if per_rank:
        line["per_rank"] = per_rank
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="utterances per GPU per step")
    ap.add_argument("--lstm-impl", default=os.environ.get("NPPC_LSTM_IMPL", "tc"), choices=["tc", "f32", "tcp"])
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the torch-eager GPU baseline leg (N = 1 only)")
    ap.add_argument("--utterances", type=int, default=1000, help="--mode sweep: utterances in the whole job (BASELINE config 5)")
    ap.add_argument("--mode", default="infer", choices=["infer", "train", "sweep"],
                    help="infer: NPPCModel.forward (the headline metric); train: BASELINE config 3, one PC-head training step per "
                         "GPU (B = 32 x 4 s, drop_band groups = 2) with the DP gradient all-reduce over NCCL when N > 1")
    ap.add_argument("--train-batch", type=int, default=32)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"NPPC-audio inference, {args.batch} x 4 s 16 kHz synthetic utterances per GPU per step, "
                          f"n_dirs={N_DIRS}, random-init FullSubNet+ + PC head (BASELINE configs[4] per-GPU micro-batch)",
              "batch_per_gpu": args.batch, "n_dirs": N_DIRS, "seconds_per_utterance": L / SR,
              "parallelism": f"utterance-sharded x{world} (no data-path collective)",
              "l2": "inputs larger than L2: every step streams > 30 GB of activations through the 126 MB L2; 256 MiB flush before the timed region"}

    if args.impl == "reference":
        if rank != 0:
            return
        # EXACTLY K timed + W warm-up steps like our arm; each step is a bounded SAMPLE of the workload: --cpu-batch of the
        # 64 utterances a step holds (audio-s/s is per audio-second, so the sample's throughput is the workload's; B = 64
        # on these host cores would be ~25 s per step).  `config` is our arm's; what actually ran is in `sample` below.
        r = cpu_reference_run(args.steps, args.warmup, args.cpu_batch)
        line = {"metric": METRIC, "value": r["value"], "unit": "audio-s/s", "impl": "reference", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "sample": {"batch_per_step": args.cpu_batch, "of_batch": args.batch,
                           "note": f"each reference step runs {args.cpu_batch} of the {args.batch} utterances of a step (bounded sample); "
                                   "reference CPU path via the oracle port, fast=True -> ATen fused CPU LSTM (the reference is Python "
                                   "and /root/reference does not travel to the GPU box)"},
                "cpu_baseline": {"value": r["value"], "unit": "audio-s/s", "cores": r["cores"], "kind": "port",
                                 "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # watchdog: a rank that stops making progress (a wedged collective, a kernel that never returns) dumps every thread's
    # Python stack to stderr and exits instead of holding the box until the caller's limit (an N = 4 training run did exactly
    # that in round 2); torchrun then tears the other ranks down.  NPPC_BENCH_WATCHDOG_S=0 disables it.
    import faulthandler
    watchdog_s = int(os.environ.get("NPPC_BENCH_WATCHDOG_S", "1500"))
    if watchdog_s > 0:
        faulthandler.dump_traceback_later(watchdog_s, exit=True)
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation (NCCL_DEBUG=VERSION in some environments);
        # stdout must carry exactly ONE JSON line, so fd 1 points at stderr while the communicator comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    import generative_audio_b200 as G
    from helpers import build_model, wave
    ops = G.ops
    dev = torch.device("cuda", local_rank)
    if args.mode == "train":
        return train_bench(args, rank, world, local_rank, dev, G, build_model, wave, ClockSampler, dist)
    if args.mode == "sweep":
        return sweep_bench(args, rank, world, local_rank, dev, G, build_model, wave, ClockSampler, dist)
    model, _ = build_model(N_DIRS, 1, args.lstm_impl)
    B = args.batch
    x_host = wave(B, L, 1000 + rank).pin_memory()
    x_dev = x_host.to(dev)
    out_host = torch.empty(B, N_DIRS, 2, F, T, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    # live CUDA-event timing of the dominant kernel (the LSTM launches) inside the timed region
    lstm_events = []
    orig_forward = ops.LstmPlan.forward

    def timed_forward(self, xs, impl, R=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = orig_forward(self, xs, impl, R)
        e1.record()
        lstm_events.append((e0, e1, xs.shape[1] if R is None else R, xs.shape[0]))
        return y

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n, fn, timed):
        """EXACTLY n steps back to back between one pair of CUDA events (the host runs ahead of the device as in a serving
        loop; no per-step synchronisation).  Every step streams > 30 GB of activations (12.8 GB of fp16 gate pre-activations
        per LSTM alone) through the 126 MB L2, so no step can reuse the previous step's cache contents; an explicit 256 MiB
        flush still runs once before the timed region."""
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        if getattr(model, "_copy_stream", None) is not None:
            torch.cuda.current_stream().wait_stream(model._copy_stream)   # pending host downloads end inside the region
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    def step_device():
        return model(x_dev)

    def step_e2e():
        # public serving call: pinned host waveform in, pinned host w_mat out; the download rides a side stream and overlaps
        # the next step's kernels (run_steps joins that stream before the closing event, so every byte is inside the region)
        model.forward_host(x_host, out_host)

    # ---- device-resident throughput -------------------------------------------------------------------
    run_steps(args.warmup, step_device, False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ops.LstmPlan.forward = timed_forward
    ops.reset_launch_count()
    ms_total = run_steps(args.steps, step_device, True)
    launches = ops.launch_count()
    ops.LstmPlan.forward = orig_forward
    barrier()
    # ---- end-to-end (host buffers) --------------------------------------------------------------------
    run_steps(1, step_e2e, False)
    barrier()
    ms_e2e = run_steps(args.steps, step_e2e, True)
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # whole-job aggregation: all ranks' audio-seconds / the slowest rank's device time (max over ranks)
    from generative_audio_b200.sharding import aggregate_throughput
    _, ms_step = aggregate_throughput(B * L / SR, ms_total / args.steps)
    _, ms_e2e_step = aggregate_throughput(B * L / SR, ms_e2e / args.steps)
    audio_s = B * world * L / SR
    per_rank = None
    if world > 1:   # which GPU is the slow one, and at what clock (the line's ms_per_step is the MAX over ranks)
        mine = torch.tensor([ms_total / args.steps, ms_e2e / args.steps, float(sampler.summary()["sm_mhz"] or 0)],
                            device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        allr = torch.stack(allr).cpu()
        ms_r = sorted(allr[:, 0].tolist())
        per_rank = {"ms_per_step": [round(v, 3) for v in allr[:, 0].tolist()], "ms_per_step_e2e": [round(v, 3) for v in allr[:, 1].tolist()],
                    "sm_mhz": [int(v) for v in allr[:, 2].tolist()], "ms_min": ms_r[0], "ms_median": ms_r[len(ms_r) // 2],
                    "ms_max": ms_r[-1]}
    lstm_ms = sum(e0.elapsed_time(e1) for e0, e1, _, _ in lstm_events) / max(len(lstm_events), 1)
    R, Tp = (lstm_events[0][2], lstm_events[0][3]) if lstm_events else (B * F, TP)
    flops = LSTM_FLOP_PER_SEQ_STEP * R * Tp
    pk = peaks()
    achieved = flops / (max(lstm_ms, 1e-9) * 1e-3) / 1e12 if lstm_events else 0.0
    roofline = {"bound": "tensor", "kernel": f"sub-band LSTM (2 layers + fc), impl={args.lstm_impl}, per nppc_lstm_forward call",
                "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
                # DRAM bytes of one nppc_lstm_forward call (rec layer 0 + Zx GEMM + rec layer 1) from the ncu --set full capture
                # profiles/r02_lstm_ncu_full_summary.csv (B=64, O=2): 3.69 + 16.00 + 17.19 GB; scaled to this R x T'
                "traffic": (36.88e9 * (R * Tp) / (64 * 257 * 253)) if args.lstm_impl == "tc" else None,
                "traffic_note": "dram__bytes_read+write per call from profiles/r02_lstm_ncu_full_summary.csv (ncu --set full, this round); 25.6 GB of it is the "
                                "fp16 gate pre-activation (Zx) round trip of layer 1",
                "peak_source": f"{pk['src']} (sustained dense 16-bit tensor peak, kernel timed inside a long step)",
                "ms_per_call": lstm_ms, "calls_per_step": len(lstm_events) / max(args.steps, 1),
                "algorithmic_flops_per_call": flops, "share_of_step": lstm_ms * len(lstm_events) / max(args.steps, 1) / ms_step}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {"metric": METRIC, "value": audio_s / (ms_step * 1e-3), "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if args.lstm_impl == "tc" else "f32", "data": "synthetic", "config": config,
            "e2e": {"value": audio_s / (ms_e2e_step * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_e2e_step,
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": sampler.summary()}
    if per_rank is not None:
        line["per_rank"] = per_rank
    if not args.no_gpu_eager and world == 1:
        try:
            line["gpu_eager_baseline"] = gpu_eager_run(B)
        except Exception as e:   # reported, never fatal: it is a side-by-side bar, not the product
            line["gpu_eager_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    if not args.no_cpu_baseline and world == 1:
        # bounded sample, ~10 s of host work: 12 timed + 2 warm-up passes over --cpu-batch utterances (3 + 1 in round 1 moved by
        # +-25 % from box to box: too few passes for a 16-thread oneDNN LSTM to settle)
        r = cpu_reference_run(12, 2, args.cpu_batch)
        line["cpu_baseline"] = {"value": r["value"], "unit": "audio-s/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
