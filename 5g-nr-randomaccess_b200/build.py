"""Build librach_gpu.so (in-tree, sm_100a) and the host CLI.  nvcc cross-compiles without a GPU."""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "librach_gpu.so")
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("rach_engine.cu", "rach_host.cpp")]
HEADERS = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + glob.glob(os.path.join(HERE, "csrc", "*.h")) +
                 glob.glob(os.path.join(ROOT, "include", "*.h")))
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
              "-I", os.path.join(HERE, "csrc")]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: librach_gpu cannot be built (there is no CPU path)")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    if force or _stale(LIB, SOURCES + HEADERS):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + SOURCES + ["-o", LIB]
        subprocess.check_call(cmd)
    return LIB


def build_host(force=False):
    """The C host program with the reference's CLI (RandomAccessWithNOMA.c:90-206)."""
    src = os.path.join(HERE, "host", "rach_sim.c")
    exe = os.path.join(HERE, "host", "rach_sim")
    if not os.path.exists(src):
        return None
    if force or _stale(exe, [src, LIB] + HEADERS):
        subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-I", os.path.join(ROOT, "include"), src,
                               "-o", exe, "-L", HERE, "-lrach_gpu", "-Wl,-rpath,$ORIGIN/..", "-lm"])
    return exe


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
