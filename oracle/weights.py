"""TEST/BENCH INFRASTRUCTURE — deterministic synthetic weights for the NPPC model (no checkpoints exist:
`*.tar` is git-ignored in the reference).  numpy PCG64 keyed by (seed, crc32(name)) so the
same tensors are produced in the dev container (for the reference run that makes tests/golden) and on the
GPU box, independent of torch's RNG or module construction order.  Scales mimic torch's default inits.
"""
import json
import os
import zlib

import numpy as np
import torch

MANIFEST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "state_dict_manifest.json")


def load_manifest(n_dirs: int = 5):
    with open(MANIFEST) as f:
        man = json.load(f)
    out = []
    for name, shape in man["entries"]:
        shape = [2 * n_dirs if s == "2*n_dirs" else s for s in shape]
        out.append((name, tuple(shape)))
    return out


def synth_tensor(name: str, shape, seed: int) -> torch.Tensor:
    rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))
    leaf = name.split(".")[-1]
    parent = name.split(".")[-2] if "." in name else ""
    if parent.startswith("norm"):  # GroupNorm affine
        a = rng.uniform(-0.1, 0.1, size=shape)
        return torch.from_numpy((a + (1.0 if leaf == "weight" else 0.0)).astype(np.float32))
    if parent.startswith("prelu"):
        return torch.from_numpy((0.25 + rng.uniform(-0.05, 0.05, size=shape)).astype(np.float32))
    if "sequence_model.weight_" in name or "sequence_model.bias_" in name:  # nn.LSTM: U(-1/sqrt(H), 1/sqrt(H))
        H = 384
        bound = 1.0 / np.sqrt(H)
    else:
        # conv / linear: fan_in from the matching weight shape
        if leaf == "weight":
            fan_in = int(np.prod(shape[1:]))
        else:
            fan_in = None
        bound = None if fan_in is None else 1.0 / np.sqrt(fan_in)
    if bound is None:
        bound = 0.05  # biases of conv/linear: fixed small bound (fan_in not recoverable from the bias shape)
    return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))


def synth_state_dict(n_dirs: int = 5, seed: int = 0, prefix: str = ""):
    """Full NPPCModel state_dict (680 tensors) or a sub-tree selected by prefix (prefix stripped)."""
    sd = {}
    for name, shape in load_manifest(n_dirs):
        if name.startswith(prefix):
            sd[name[len(prefix):]] = synth_tensor(name, shape, seed)
    return sd


# ---- inpainting variant (a16): UNet tensors by name; scales keep 18 conv layers of activations O(1) ----------------
def synth_unet_tensor(name: str, shape, seed: int) -> torch.Tensor:
    rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))
    leaf = name.split(".")[-1]
    shape = tuple(shape)
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_mean":
        return torch.from_numpy(rng.uniform(-0.1, 0.1, size=shape).astype(np.float32))
    if leaf == "running_var":
        return torch.from_numpy((1.0 + rng.uniform(-0.2, 0.2, size=shape)).astype(np.float32))
    if len(shape) == 1 and leaf == "weight":   # BatchNorm gamma
        return torch.from_numpy((1.0 + rng.uniform(-0.1, 0.1, size=shape)).astype(np.float32))
    if leaf == "bias":
        return torch.from_numpy(rng.uniform(-0.05, 0.05, size=shape).astype(np.float32))
    fan_in = int(np.prod(shape[1:]))
    bound = float(np.sqrt(6.0 / (1.04 * fan_in)))   # kaiming-uniform for LeakyReLU(0.2)
    return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))


def synth_unet_state_dict(shapes, seed: int = 0, tag: str = ""):
    """shapes: iterable of (name, shape) (e.g. from module.state_dict()); `tag` separates the two UNets of the model."""
    return {name: synth_unet_tensor(tag + name, shape, seed) for name, shape in shapes}
