"""N>1 host logic on CPU (gloo, world_size 2): utterance sharding of the inference sweep (BASELINE config 5) and the
max-over-ranks timing / whole-job aggregation bench.py uses.  The data path has NO collective (independent utterances);
the only communication is the timing all-reduce."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from generative_audio_b200.sharding import shard_utterances, aggregate_throughput
    n_utt, micro = 1000, 64
    mine = shard_utterances(n_utt, rank, world)
    # every utterance is owned by exactly one rank, round-robin
    owned = torch.zeros(n_utt, dtype=torch.int64)
    owned[mine] = 1
    dist.all_reduce(owned)
    assert bool((owned == 1).all())
    assert abs(len(mine) - n_utt / world) <= 1
    batches = [mine[i:i + micro] for i in range(0, len(mine), micro)]
    assert sum(len(b) for b in batches) == len(mine) and all(len(b) <= micro for b in batches)
    # timing aggregation: value = all units / max-over-ranks time
    my_ms = 100.0 + 50.0 * rank
    value, ms = aggregate_throughput(len(mine) * 4.0, my_ms, backend_device="cpu")
    if rank == 0:
        ret["value"], ret["ms"] = value, ms
    dist.destroy_process_group()


def test_sharding_and_aggregation_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29631, ret), nprocs=2, join=True)
    assert abs(ret["ms"] - 150.0) < 1e-9                     # max over ranks
    assert abs(ret["value"] - 1000 * 4.0 / 0.150) < 1e-6     # whole-job audio-seconds / slowest rank's time


def test_shard_edge_cases():
    from generative_audio_b200.sharding import shard_utterances
    assert shard_utterances(0, 0, 4) == []
    assert shard_utterances(3, 3, 4) == []            # more ranks than utterances: empty shard
    assert shard_utterances(5, 1, 2) == [1, 3]
    all_idx = sorted(i for r in range(8) for i in shard_utterances(1000, r, 8))
    assert all_idx == list(range(1000))


def _grad_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from generative_audio_b200.training import allreduce_gradients
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 1000, 3, 70000)]
    params.append(torch.nn.Parameter(torch.zeros(4)))  # no grad: must be skipped
    for i, p in enumerate(params[:4]):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    calls = allreduce_gradients(params, bucket_bytes=2048)
    expect = [(1 + 2) / 2 * (i + 1) for i in range(4)]   # mean over 2 ranks
    ok = all(torch.allclose(p.grad, torch.full_like(p, e)) for p, e in zip(params[:4], expect)) and params[4].grad is None
    if rank == 0:
        ret["ok"], ret["calls"] = ok, calls
    dist.destroy_process_group()


def test_dp_gradient_allreduce_world2():
    """DP training exchange: flat-bucket mean all-reduce of the PC-head gradients."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_grad_worker, args=(2, 29641, ret), nprocs=2, join=True)
    assert ret["ok"] and ret["calls"] >= 2


def _reducer_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from generative_audio_b200.training import GradBucketReducer
    torch.manual_seed(0)                                   # identical weights on both ranks
    net = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(), torch.nn.Linear(64, 3))
    red = GradBucketReducer(net.parameters(), bucket_bytes=4096)
    xs = [torch.randn(8, 16, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    # single-process reference: mean over ranks of the per-rank objective's gradients
    ref = None
    for r in range(world):
        net.zero_grad(set_to_none=True)
        red.reset()
        # (hooks fire here too; finish() below is only called for the real DP step)
        net(xs[r]).pow(2).mean().backward()
        g = [p.grad.clone() for p in net.parameters()]
        ref = g if ref is None else [a + b for a, b in zip(ref, g)]
        for _, _, _, w in red._work:
            w.wait()
    ref = [g / world for g in ref]
    dist.barrier()
    net.zero_grad(set_to_none=True)
    red.reset()
    net(xs[rank]).pow(2).mean().backward()                 # buckets are exchanged from the hooks while backward runs
    launched_in_backward = len(red._work)
    n = red.finish()
    ok = all(torch.allclose(p.grad, g, atol=1e-6) for p, g in zip(net.parameters(), ref))
    # a parameter the graph never reaches shares a bucket with live ones: its bucket's count-down never finishes, and finish()
    # must still average the members that have a gradient (it used to skip such a bucket silently)
    red.remove()
    unused = torch.nn.Parameter(torch.zeros(7))
    red2 = GradBucketReducer(list(net.parameters()) + [unused], bucket_bytes=1 << 30)       # ONE bucket
    net.zero_grad(set_to_none=True)
    red2.reset()
    net(xs[rank]).pow(2).mean().backward()
    early2 = len(red2._work)
    n2 = red2.finish()
    ok2 = early2 == 0 and n2 == 1 and unused.grad is None and \
        all(torch.allclose(p.grad, g, atol=1e-6) for p, g in zip(net.parameters(), ref))
    if rank == 0:
        ret["ok"], ret["n"], ret["early"], ret["buckets"], ret["ok_unused"] = ok, n, launched_in_backward, len(red.buckets), ok2
    dist.destroy_process_group()


def test_dp_bucket_reducer_overlaps_backward_world2():
    """GradBucketReducer: buckets are all-reduced from post-accumulate hooks during backward; the averaged gradients equal
    the mean of the per-rank objectives' gradients (SURVEY §8e DP parity)."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_reducer_worker, args=(2, 29651, ret), nprocs=2, join=True)
    assert ret["ok"] and ret["buckets"] >= 2 and ret["n"] == ret["buckets"] and ret["early"] == ret["buckets"]
    assert ret["ok_unused"]
