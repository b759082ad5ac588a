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


def build(force=False, verbose=False, defines=(), out=OUT):
    """defines / out: build a variant of the library (e.g. -DYASPH_SWEEP_MIN_CTAS=5) next to the default one for A/B timing."""
    if not force and not defines and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in sources()):
        return out
    cmd = [NVCC] + FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, SRC]
    subprocess.check_call(cmd)
    # Bit-exactness guard: the path computes in strict f32 without contraction.  ptxas is known to contract
    # packed mul.rn.f32x2 + add/sub.rn.f32x2 into FFMA2 despite -fmad=false; no kernel of this library may contain one.
    sass = subprocess.run(["cuobjdump", "-sass", out], capture_output=True, text=True).stdout
    if "FFMA2" in sass:
        os.remove(out)
        raise RuntimeError("libyasph_gpu.so: ptxas emitted FFMA2 (fused packed multiply-add): results would not match the reference bit for bit")
    return out


if __name__ == "__main__":
    import sys

    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith("-o")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=os.path.abspath(outs[0]) if outs else OUT))
