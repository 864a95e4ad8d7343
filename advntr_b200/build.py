"""Build the CUDA engine (``libadvhmm.so``) in-tree with nvcc for sm_100a.

``python -m advntr_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU.  The library is the only thing that can decode: nothing in the package
falls back to the CPU when it is missing.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(PKG, "csrc", "advhmm.cu")
DEPS = [SRC] + [os.path.join(PKG, "csrc", f) for f in
                ("model_compile.hpp", "engine_types.cuh", "kernels_common.cuh", "kernels_banded.cuh",
                 "kernels_generic.cuh", "kernels_kfilter.cuh")] + [os.path.join(os.path.dirname(PKG), "include", "advhmm.h")]
LIB = os.path.join(PKG, "libadvhmm.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "1886"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libadvhmm.so")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
