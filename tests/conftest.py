import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    import torch
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files if z[k].dtype.kind in "fiub"}   # (string arrays: names only)


def rel_err(a, b):
    """‖a−b‖∞ / ‖b‖∞ per tensor (SURVEY.md §8d tolerance definition)."""
    import torch
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    d = (a - b).abs().max().item()
    return d / max(b.abs().max().item(), 1e-30)


@pytest.fixture(scope="session")
def golden():
    return load_golden
