"""Builds libbroadphase_b200.so (hand-written CUDA for sm_100a + the C ABI of include/bp.h) in-tree.

nvcc cross-compiles without a GPU.  -fmad=false keeps the f32 quantiser bit-identical to the
reference (one IEEE rounding per operation); -lineinfo lets ncu's source page map to the .cu files.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libbroadphase_b200.so")
SOURCES = ["bp_layer.cu", "bp_dist.cu"]
HEADERS = ["bp_common.cuh", "bp_encode.cuh", "bp_radix.cuh", "bp_exchange.cuh", "bp_scan.cuh", "bp_merge.cuh", "bp_query.cuh", "../../include/bp.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
    "--shared", "-cudart", "shared",
    "--threads", "2",   # the two translation units side by side
]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, s) for s in SOURCES] + ["-o", SO]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libbroadphase_b200.so")
    return SO


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(SO)
