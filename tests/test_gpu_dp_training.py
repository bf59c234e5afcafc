"""DP training on >= 2 GPUs over NCCL (row J2): runs tools/dp_train_check.py under torch.distributed.run and checks that the
overlapped bucket all-reduce produced the mean of the per-rank gradients.  Skipped on a single-GPU box."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp_gradients_equal_mean_of_per_rank_gradients():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29713", os.path.join(ROOT, "tools", "dp_train_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["ok"] and out["collectives"] == out["buckets"] >= 2, out
