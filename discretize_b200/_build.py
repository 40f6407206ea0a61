"""In-tree build of libdzb.so (hand-written sm_100a CUDA behind a C ABI).

``python -m discretize_b200._build`` or ``__graft_entry__.build()``.  nvcc
cross-compiles without a GPU; the resulting ``discretize_b200/libdzb.so`` is
git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdzb.so")
OBJ = os.path.join(HERE, "csrc", "_obj")

NVCC_FLAGS = [
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3",
    "--extended-lambda", "--expt-relaxed-constexpr",
    "-diag-suppress", "549",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libdzb.so")
    return nvcc


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = True) -> str:
    """Compile every .cu under csrc/ (only the stale ones) and link libdzb.so."""
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "dzb.h")]
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    extra = os.environ.get("DZB_EXTRA_NVCC_FLAGS", "").split()   # e.g. -DDZB_APPLY_SWEEP for tools/sweep_elementary.py
    objs, procs = [], []
    for src in srcs:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
            if verbose:
                print("[dzb build]", " ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
        if verbose and out.strip():
            print(out)
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
        if verbose:
            print("[dzb build]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB)
