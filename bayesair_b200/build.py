"""Builds ``bayesair_b200/_lib/libbayesair_b200.so`` in-tree with nvcc for sm_100a.

The library is the product: there is no fallback if it is missing (``_capi`` raises).
nvcc cross-compiles without a GPU, so this also runs in the build container; the
built ``.so`` travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "_lib")
# BA_LIB_PATH selects an alternative build of the same library (tuning experiments)
LIB_PATH = os.environ.get("BA_LIB_PATH") or os.path.join(LIB_DIR, "libbayesair_b200.so")
SOURCES = ["ba_kernels.cu", "ba_guide.cu", "ba_capi.cu", "ba_plan_build.cpp"]
HEADERS = ["ba_sim.cuh", "ba_scan2.cuh", "ba_kernels.cuh", "ba_plan.h", "ba_plan_build.h",
           os.path.join("..", "..", "include", "bayesair_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libbayesair_b200.so")


def needs_build() -> bool:
    if os.environ.get("BA_LIB_PATH") and os.path.exists(LIB_PATH):
        return False
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    extra = os.environ.get("BA_NVCC_EXTRA", "").split()
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH] + [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
