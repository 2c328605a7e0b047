"""Build ``libxtrack_b200.so`` in-tree with nvcc for sm_100a (no JIT cache, no torch dependency).

    python -m extrack_b200.build
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "xt_engine.cu")
OUT = os.path.join(HERE, "libxtrack_b200.so")


def _deps():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in os.listdir(d)] + [os.path.join(ROOT, "include", "xtrack.h")]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build the CUDA engine (there is no CPU fallback)")


def up_to_date() -> bool:
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(f) <= t for f in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "csrc"),
        "-shared", "-Xcompiler", "-fPIC", "-o", OUT, SRC,
    ]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
