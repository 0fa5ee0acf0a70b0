"""In-tree build of libusrt_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m unitysimpleraytracing_b200.build [--force] [--verbose]

The .so lands next to this file so that it travels with the repo snapshot to the GPU box.
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OBJ_DIR = os.path.join(_HERE, "csrc", "build")
LIB_PATH = os.path.join(_HERE, "libusrt_b200.so")
SOURCES = ["api.cu", "morton.cu", "radix_sort.cu", "lbvh.cu", "trace.cu", "shade.cu"]
HEADERS = [os.path.join(CSRC, "usrt_internal.cuh"), os.path.join(_HERE, "..", "include", "usrt.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: no FMA contraction anywhere (the fp32 kernels also spell every op with _rn
# intrinsics; this is belt and braces). Division and sqrt are IEEE by default (no --use_fast_math).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + HEADERS):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
    if force or _stale(LIB_PATH, objs):
        cmd = [NVCC, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                          "-Xcompiler", "-fPIC"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB_PATH)
