"""Build recipe of the product library: nvcc, sm_100a only, in-tree output (ttcross_b200/libttcross_b200.so)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("TTC_BUILD_OUT", os.path.join(HERE, "libttcross_b200.so"))
SOURCES = ["ttc_engine.cu"]
NVCC = os.environ.get("TTC_NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                 # reference arithmetic has no FMA contraction
    "-Xcompiler", "-fPIC", "-shared",
    "-ccbin", "/usr/bin/g++",
    "-cudart", "shared", "-ldl", "-Xcompiler", "-pthread",
] + os.environ.get("TTC_NVCC_EXTRA", "").split()


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ttcross_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if a source is newer than it.  Several processes may arrive at once (one per GPU under torchrun):
    an exclusive file lock serialises them, the compiler writes to a temporary name and the result is renamed into place, so
    nobody ever loads a half-written library and only the first arrival compiles."""
    if not (force or needs_build()):
        return LIB
    import fcntl
    with open(LIB + ".lock", "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            if force or needs_build():
                tmp = f"{LIB}.tmp{os.getpid()}"
                cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
                if verbose:
                    print(" ".join(cmd))
                subprocess.check_call(cmd)
                os.replace(tmp, LIB)
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
