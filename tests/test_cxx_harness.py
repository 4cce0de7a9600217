"""Builds tests/cxx_harness.cpp with g++ against the C++ host mirror (include/massiv_b200.hpp) and runs it:
the reference's constructors and calls by name, the reference's expected values — in a compiled language,
since the reference's own (Haskell) cannot be compiled in this image."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "cxx_harness")


def _build():
    from massiv_b200 import _abi
    _abi.lib()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.join(ROOT, "massiv_b200")
    cmd = ["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "cxx_harness.cpp"),
           "-o", EXE, "-L", libdir, "-l:libmassiv_b200.so", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cxx_mirror_geometry():
    _build()
    r = subprocess.run([EXE, "geometry"], capture_output=True, text=True, timeout=60)
    assert "CXX_HARNESS_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cxx_mirror_reference_vectors():
    _build()
    r = subprocess.run([EXE, "gpu"], capture_output=True, text=True, timeout=120)
    assert "CXX_HARNESS_OK" in r.stdout, r.stdout + r.stderr
