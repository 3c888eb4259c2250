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
]
LINK_FLAGS = ["--shared", "-cudart", "shared"]
OBJ_DIR = os.path.join(HERE, "_build")   # one object per translation unit (git- and gpurun-ignored): an edit of bp_dist.cu
                                         # does not recompile the four-minute bp_layer.cu


def _deps(src):
    return [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _obj(src):
    return os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")


def needs_build():
    return _stale(SO, [d for s in SOURCES for d in _deps(s)])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs = []
    for src in SOURCES:   # the translation units side by side
        if force or _stale(_obj(src), _deps(src)):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", _obj(src)]
            jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in jobs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed.append(src)
    if failed:
        raise RuntimeError("nvcc failed compiling " + ", ".join(failed))
    r = subprocess.run([nvcc] + NVCC_FLAGS + LINK_FLAGS + [_obj(s) for s in SOURCES] + ["-o", SO], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed linking libbroadphase_b200.so")
    return SO


if __name__ == "__main__":
    build(force="-f" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
