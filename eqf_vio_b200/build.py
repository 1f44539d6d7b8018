"""Builds eqf_vio_b200/csrc/libeqvio_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libeqvio_b200.so")
SOURCES = ["dgemm_sm100.cu", "ozaki_sm100.cu", "filter_kernels.cu", "eqvio_capi.cu"]
HEADERS = ["dgemm_sm100.cuh", "ozaki_sm100.cuh", "filter_kernels.cuh", "kernels_api.cuh", "eqvio_math.cuh", os.path.join("..", "..", "include", "eqvio.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2",
    "--fmad=false",  # fp64 contraction off in the O(N) kernels: keeps their round-off closer to the reference's un-fused Eigen code
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return exe


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), lib: str = LIB) -> str:
    """`extra_flags` / `lib` build an instrumented variant next to the product library (tools/ only)."""
    if lib == LIB and not force and not needs_build():
        return LIB
    tag = "" if lib == LIB else "." + os.path.basename(lib).replace(".so", "")
    objs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", tag + ".o"))
        cmd = [nvcc(), *NVCC_FLAGS, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        objs.append(obj)
    cmd = [nvcc(), "-shared", "-o", lib, *objs, "-lcudart"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
