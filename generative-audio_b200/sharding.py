"""Multi-GPU plumbing for the hot path: utterances are independent, so N GPUs = N processes that
each run the single-GPU path on a round-robin shard — no data-path collective.  torch.distributed (NCCL on GPUs,
gloo in the CPU tests) is used only to agree on the timing (max over ranks) and to count the work."""
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_utterances(n_utterances: int, rank: int, world: int) -> List[int]:
    """Round-robin shard (BASELINE config 5): utterance i belongs to rank i % world."""
    return list(range(rank, n_utterances, world))


def aggregate_throughput(local_units: float, local_ms: float, backend_device: str = "cuda") -> Tuple[float, float]:
    """(whole-job units per second, max-over-ranks milliseconds). Collective: every rank must call it."""
    t = torch.tensor([local_units, local_ms], dtype=torch.float64, device=backend_device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        units, ms = t[0:1].clone(), t[1:2].clone()
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        t = torch.cat([units, ms])
    units, ms = t.tolist()
    return units / (ms * 1e-3), ms
