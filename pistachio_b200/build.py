"""Builds libpistachio_b200.so in-tree with nvcc for sm_100a (the only target).

    python -m pistachio_b200.build [--force] [--verbose]

The shared library carries the C ABI declared in include/pistachio_b200.h; nothing in it
depends on torch or Python.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libpistachio_b200.so")
SOURCES = [os.path.join(CSRC, "pst_api.cu")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("pst_kernels.cuh", "pst_chain.cuh", "pst_math.cuh", "pst_bulk.cuh", "pst_f32.cuh", "pst_fit.cuh")] + [
    os.path.join(ROOT, "include", "pistachio_b200.h")
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--fmad=true",            # FMA contraction on; the phase path opts out with __dmul_rn (pst_math.cuh)
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-O3",
]
# fast-math must stay off: the phase path (mul_rn / div_rn, pst_math.cuh) and the table lookup rely on IEEE rounding
EXTRA_FLAGS = os.environ.get("PST_NVCC_EXTRA", "").split()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + EXTRA_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES + ["-lcudart"]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
