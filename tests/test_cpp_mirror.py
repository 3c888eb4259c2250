"""The C++ host mirror (cpp/broadphase/layer.hpp) compiles against include/bp.h and, on a GPU box,
runs the reference's crate-level doc example (src/lib.rs:24-47) through the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "broadphase-rs_b200")
SRC = os.path.join(PKG, "cpp", "tests", "doc_example.cpp")
EXE = os.path.join(PKG, "cpp", "tests", "doc_example.bin")


def _build(bp):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", SRC, "-o", EXE, "-L" + PKG,
                    "-lbroadphase_b200", "-Wl,-rpath," + PKG, "-L/usr/local/cuda/lib64"], check=True)


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(bp):
    _build(bp)
    r = subprocess.run([EXE], capture_output=True, text=True)
    if bp.device_count() == 0:
        assert r.returncode == 77 and "CUDA error" in r.stdout   # no CPU fallback
    else:
        assert r.returncode == 0, r.stdout


@pytest.mark.gpu
def test_cpp_mirror_doc_example(bp):
    _build(bp)
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "pairs=1" in r.stdout and "(9, 7)" in r.stdout
