"""Builds the in-tree sm_100a shared library `csrc/libqimg_b200.so` with nvcc.

nvcc cross-compiles for sm_100a without a GPU, so this runs in the CPU build container;
the built .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(CSRC, "libqimg_b200.so")
SOURCES = ["qimg_api.cu", "qimg_engine.cu", "qimg_tp_p2p.cu"]
HEADERS = ["qimg_common.cuh", "qimg_elementwise.cuh", "qimg_gemm.cuh", "qimg_gemm2.cuh", "qimg_fmha.cuh", "qimg_fmha2.cuh", "qimg_fmha3.cuh", "qimg_fmha4.cuh", "qimg_fmha5.cuh", "qimg_fmha6.cuh", "qimg_host.cuh",
           os.path.join("..", "..", "include", "qimg_b200.h")]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found (needed to build libqimg_b200.so)")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "--use_fast_math" if os.environ.get("QIMG_FAST_MATH") else "-DQIMG_NO_FAST_MATH",
           *(["-DQIMG_FMHA_TRACE"] if os.environ.get("QIMG_FMHA_TRACE") else []),
           *(["-DQIMG_FMHA_NOCLAMP"] if os.environ.get("QIMG_FMHA_NOCLAMP") else []),  # experiment only
           "-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
