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
# per-source extra flags.  viterbi_kernels.cu is compiled with ptxas -O1: at -O2/-O3 the ptxas
# scheduler hoists the packed-min instructions far ahead of the predicated IMADs that consume their
# predicate outputs, runs out of the 7 predicate registers and spills predicates through LOP3 pairs
# (2642 vs 607 LOP3 in the loop body); -O1 keeps program order, which is already interleaved.
SOURCES = {"viterbi_kernels.cu": ["-Xptxas", "-O1"], "viterbi_warp_kernel.cu": [], "rs_kernels.cu": [], "fec_api.cu": []}
HEADERS = [os.path.join(CSRC, "fec_internal.h"), os.path.join(CSRC, "rs_chien_bitsliced.h"), os.path.join(CSRC, "rs_decode.h"), os.path.join(CSRC, "viterbi_pair_core.h"),
           os.path.join(CSRC, "rs_bitslice_tables.h"), os.path.join(os.path.dirname(HERE), "include", "viterbi_b200.h")]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden"]
OBJ_DIR = os.path.join(HERE, "build")


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
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs = []
    for src, extra in SOURCES.items():
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [cc] + ARCH_FLAGS + extra + ["-c", "-o", obj, os.path.join(CSRC, src)]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        objs.append(obj)
    cmd = [cc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
