"""Build deepdish_b200/libdeepdish_b200.so for sm_100a with nvcc (in-tree, no JIT cache).

    python -m deepdish_b200.build [--force]

-fmad=false: f64/f32 arithmetic whose rounding decides a bit-exact output (gating, IoU, count-line,
LSAP duals) must not be contracted into FMAs; the dot products that want FMAs call fmaf explicitly.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdeepdish_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "deepdish_b200.h")]
    if not force and os.path.exists(OUT) and all(
            os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("DD_NVCC_EXTRA", "").split()       # A/B builds: e.g. DD_NVCC_EXTRA="-DDD_GS_CW=8"
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + sources() + ["-o", OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
