"""In-tree build of libyasph_gpu.so for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "yasph_gpu.cu")
OUT = os.path.join(_HERE, "libyasph_gpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: the reference computes f32 without FMA contraction; IEEE div/sqrt are nvcc's defaults.
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", "-ccbin", "/usr/bin/g++",
]


def sources():
    d = os.path.join(_HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + [os.path.join(_HERE, "..", "include", "yasph_gpu.h")]


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in sources()):
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
