"""The C++ host mirror (include/seekstorm_b200.hpp: ssb::Index::search with the reference's Search::search signature)
compiled with g++ against the C-ABI library and run on the reference's own fixtures (tests/cpp/test_reference_fixtures.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_reference_fixtures.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "cpp", "_build")
EXE = os.path.join(OUT_DIR, "test_reference_fixtures")


def _build():
    os.makedirs(OUT_DIR, exist_ok=True)
    lib_dir = os.path.join(ROOT, "seekstorm_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE, "-L", lib_dir,
           "-lseekstorm_b200", f"-Wl,-rpath,{lib_dir}", "-L/usr/local/cuda/lib64", "-lcudart"]
    subprocess.check_call(cmd)


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu():
    import torch
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stdout, r.stdout + r.stderr     # no CPU fallback


@pytest.mark.gpu
def test_cpp_mirror_reference_fixtures():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
