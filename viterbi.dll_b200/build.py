"""Build libviterbi_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

The shared library links the CUDA runtime statically, so it loads on machines without a GPU
(only calling into it needs one) and travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libviterbi_b200.so")
SOURCES = ["viterbi_kernels.cu", "rs_kernels.cu", "fec_api.cu"]
HEADERS = [os.path.join(CSRC, "fec_internal.h"), os.path.join(os.path.dirname(HERE), "include", "viterbi_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-cudart", "static"]


def nvcc() -> str | None:
    return shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else None)


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cc = nvcc()
    if cc is None:
        raise RuntimeError("nvcc not found: cannot build libviterbi_b200.so (there is no CPU fallback)")
    cmd = [cc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
