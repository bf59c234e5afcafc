"""Builds libnppc_b200.so (all CUDA sources under csrc/) for sm_100a, in-tree, with plain nvcc.

    python generative-audio_b200/build.py [--force]

The shared library has a pure C ABI (include/nppc_b200.h) and no dependency on torch or libcuda at link time
(driver entry points are resolved at run time), so the same file loads in the CPU-only dev container
(symbol check) and on the B200 box.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libnppc_b200.so")
CUTLASS_INC = "/opt/prime-rl/.venv/lib/python3.12/site-packages/flashinfer/data/cutlass/include"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = (["-DNPPC_REC_TRACE"] if os.environ.get("NPPC_REC_TRACE") else []) + \
        (["-DNPPC_LSTM_ABLATE"] if os.environ.get("NPPC_LSTM_ABLATE") else []) + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/nppc_b200.h"]:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = os.path.join(OBJ, "stamp")
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == _digest()


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    # one builder at a time: the N ranks of a torchrun launch all come through here on a box that received no prebuilt
    # library; the first one compiles, the others wait on the lock and then find the stamp current
    import fcntl
    with open(os.path.join(OBJ, "lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
                return LIB
            return _build_locked(dig, stamp, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(dig: str, stamp: str, verbose: bool) -> str:

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, _sources()))
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs, "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)      # atomic: a process that already mapped the old file keeps it, nobody sees a half-written one
    with open(stamp, "w") as f:
        f.write(dig)
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
