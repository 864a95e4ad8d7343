"""Build the CUDA engine (``libadvhmm.so``, nvcc, sm_100a) and the host-side BAM reader
(``libadvbam.so``, g++ + zlib) in-tree.

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
                ("model_compile.hpp", "locus_compile.hpp", "locus_calls.hpp", "engine_types.cuh", "kernels_common.cuh", "kernels_banded.cuh",
                 "kernels_generic.cuh", "kernels_kfilter.cuh")] + [os.path.join(os.path.dirname(PKG), "include", "advhmm.h")]
LIB = os.path.join(PKG, "libadvhmm.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared", "-diag-suppress", "1886"]


BAM_SRC = os.path.join(PKG, "csrc", "bam_ingest.cpp")
BAM_DEPS = [BAM_SRC, os.path.join(os.path.dirname(PKG), "include", "advbam.h")]
BAM_LIB = os.path.join(PKG, "libadvbam.so")


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


def build_bam_library(force: bool = False) -> str:
    """The read-ingest library (BGZF / BAM / BAI reader, include/advbam.h): host C++ only."""
    if not force and os.path.exists(BAM_LIB) and \
            all(os.path.getmtime(d) <= os.path.getmtime(BAM_LIB) for d in BAM_DEPS if os.path.exists(d)):
        return BAM_LIB
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", BAM_LIB, BAM_SRC,
                           "-lz", "-lpthread"])
    return BAM_LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_bam_library(force="--force" in sys.argv))
