"""Build tests/hostemu/libdd_hostemu.so (g++, host only).  TEST SCAFFOLDING -- see dd_hostemu.cpp."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libdd_hostemu.so")
SRC = os.path.join(HERE, "dd_hostemu.cpp")
CSRC = os.path.join(HERE, "..", "..", "deepdish_b200", "csrc")


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "..", "include", "deepdish_b200.h"))
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", SRC,
           "-o", SO]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force=True))
