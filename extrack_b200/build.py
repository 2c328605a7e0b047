"""Build ``libxtrack_b200.so`` in-tree with nvcc for sm_100a (no JIT cache, no torch dependency).

    python -m extrack_b200.build [--force] [-v] [-DMACRO ...] [--out=path]

One translation unit per kernel family (``csrc/xt_k*.cu``) plus the host driver (``csrc/xt_engine.cu``),
compiled in parallel and linked into one shared library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
UNITS = ["xt_engine.cu", "xt_k1.cu", "xt_k1w.cu", "xt_k1x.cu", "xt_k1v.cu", "xt_k2f.cu", "xt_k2f32.cu", "xt_k2old.cu", "xt_k3.cu", "xt_k4.cu"]
OUT = os.path.join(HERE, "libxtrack_b200.so")
OBJ_DIR = os.path.join(HERE, "build")


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "xtrack.h")]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build the CUDA engine (there is no CPU fallback)")


def up_to_date(out: str = OUT) -> bool:
    if not os.path.isfile(out):
        return False
    t = os.path.getmtime(out)
    return all(os.path.getmtime(f) <= t for f in _deps())


def build(force: bool = False, verbose: bool = False, defines=(), out: str = OUT) -> str:
    """Compile every CUDA translation unit for sm_100a and link ``out``.  ``defines``: extra -D macros
    (profiling variants of the library, e.g. ``("XT_K1_PROF",)``, are linked under another name)."""
    if not force and not defines and up_to_date(out):
        return out
    nvcc = nvcc_path()
    tag = "_".join(defines) if defines else "default"
    obj_dir = os.path.join(OBJ_DIR, tag)
    os.makedirs(obj_dir, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
             "-I" + os.path.join(ROOT, "include"), "-I" + CSRC] + ["-D" + d for d in defines]
    if verbose:
        flags.insert(0, "-Xptxas=-v")

    def includes(path, seen):
        """Files `path` includes (quoted includes, recursively), for per-unit rebuild decisions."""
        if path in seen or not os.path.isfile(path):
            return seen
        seen.add(path)
        for line in open(path):
            if line.startswith('#include "'):
                inc = line.split('"')[1]
                for base in (CSRC, os.path.join(ROOT, "include")):
                    includes(os.path.join(base, inc), seen)
        return seen

    def compile_unit(name):
        src = os.path.join(CSRC, name)
        obj = os.path.join(obj_dir, name[:-3] + ".o")
        newest = max(os.path.getmtime(f) for f in includes(src, set()))
        if not force and not verbose and os.path.isfile(obj) and os.path.getmtime(obj) >= newest:
            return obj, ""
        r = subprocess.run([nvcc] + flags + ["-c", "-o", obj, src], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {name}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr if verbose else ""

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as ex:
        res = list(ex.map(compile_unit, UNITS))
    if verbose:
        for _, log in res:
            if log:
                print(log)
    subprocess.run([nvcc, "-shared", "-o", out] + [o for o, _ in res], check=True)
    return out


if __name__ == "__main__":
    defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    outp = OUT
    for a in sys.argv[1:]:
        if a.startswith("--out="):
            outp = a[6:]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outp))
