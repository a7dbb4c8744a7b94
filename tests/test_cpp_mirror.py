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


CSRC = os.path.join(ROOT, "tests", "cpp", "test_concurrent.c")
CEXE = os.path.join(OUT_DIR, "test_concurrent")


def _build_c():
    os.makedirs(OUT_DIR, exist_ok=True)
    lib_dir = os.path.join(ROOT, "seekstorm_b200")
    subprocess.check_call(["gcc", "-std=c11", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), CSRC, "-o", CEXE, "-L", lib_dir,
                           "-lseekstorm_b200", f"-Wl,-rpath,{lib_dir}", "-L/usr/local/cuda/lib64", "-lcudart", "-lpthread", "-lm"])


def test_c_header_compiles_as_plain_c():
    """include/seekstorm_b200.h is a C header: a plain-C caller compiles and links against the library; without a GPU it exits 3."""
    import torch
    _build_c()
    if not torch.cuda.is_available():
        r = subprocess.run([CEXE], capture_output=True, text=True, timeout=120)
        assert r.returncode == 3 and "no CUDA device" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_concurrent_searches_from_8_threads():
    """8 threads x (vector, lexical, hybrid) on one handle == the serial results (search-context pool)."""
    _build_c()
    r = subprocess.run([CEXE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
