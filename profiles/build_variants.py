"""Build A/B variants of libviterbi_b200.so into gpurun_variants/ (git-ignored, travels with gpurun).

usage: python profiles/build_variants.py name:"-DVIT_SYM_PREFETCH=0" name2:"-DX=1 -DY=2" ...
       a name of the form  git:<rev>  builds viterbi_kernels.cu of that revision (baseline for the A/B).
Then  bash profiles/variant_bench.sh  under gpurun benches every variant on the FIC and MSC shapes.
"""
import os
import shlex
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "viterbi.dll_b200", "csrc")
OUT = os.path.join(ROOT, "gpurun_variants")
BASE = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
        "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I", CSRC]


def build(name: str, flags: str) -> None:
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        vit_src = os.path.join(CSRC, "viterbi_kernels.cu")
        if name.startswith("git:"):
            rev = name[4:]
            vit_src = os.path.join(tmp, "viterbi_kernels_rev.cu")
            with open(vit_src, "w") as f:
                f.write(subprocess.run(["git", "-C", ROOT, "show", rev + ":viterbi.dll_b200/csrc/viterbi_kernels.cu"],
                                       check=True, capture_output=True, text=True).stdout)
            name = "rev_" + rev.replace("/", "_")
        objs = []
        for src, extra in ((vit_src, ["-Xptxas", "-O1"] + shlex.split(flags)),
                           (os.path.join(CSRC, "viterbi_warp_kernel.cu"), shlex.split(flags)),
                           (os.path.join(CSRC, "rs_kernels.cu"), shlex.split(flags)),
                           (os.path.join(CSRC, "fec_api.cu"), shlex.split(flags))):
            obj = os.path.join(tmp, os.path.basename(src) + ".o")
            subprocess.run(BASE + extra + ["-c", "-o", obj, src], check=True)
            objs.append(obj)
        lib = os.path.join(OUT, name + ".so")
        subprocess.run(["nvcc", "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-o", lib] + objs, check=True)
        print("built", lib)


if __name__ == "__main__":
    for arg in sys.argv[1:]:
        n, _, fl = arg.partition(":") if not arg.startswith("git:") else (arg, "", "")
        build(n, fl)
