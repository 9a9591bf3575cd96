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
SOURCES = ["qimg_api.cu", "qimg_engine.cu", "qimg_tp_p2p.cu", "qimg_vae.cu"]
HEADERS = ["qimg_common.cuh", "qimg_elementwise.cuh", "qimg_gemm.cuh", "qimg_gemm2.cuh", "qimg_fmha.cuh", "qimg_fmha4.cuh",
           "qimg_fmha6.cuh", "qimg_host.cuh", "qimg_tp.h", os.path.join("..", "..", "include", "qimg_b200.h")]
FLAGS_STAMP = LIB_PATH + ".flags"  # the optional build flags the existing .so was compiled with


def _opt_flags() -> list[str]:
    return ["--use_fast_math" if os.environ.get("QIMG_FAST_MATH") else "-DQIMG_NO_FAST_MATH",
            *(["-DQIMG_FMHA_TRACE"] if os.environ.get("QIMG_FMHA_TRACE") else [])]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found (needed to build libqimg_b200.so)")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    try:  # a .so built with other optional flags (trace counters, fast math) is stale for this environment
        with open(FLAGS_STAMP) as f:
            if f.read().split() != _opt_flags():
                return True
    except OSError:
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           *_opt_flags(), "-shared", "-Xcompiler", "-fPIC", "-ldl", "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    with open(FLAGS_STAMP, "w") as f:
        f.write(" ".join(_opt_flags()))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
