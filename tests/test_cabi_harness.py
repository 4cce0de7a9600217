"""Builds tests/cabi_harness.c with plain gcc against include/massiv_b200.h + libmassiv_b200.so and
runs it: the C ABI is usable with no Python, torch or C++ in the process (what the Haskell FFI does)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "cabi_harness")


def _build():
    from massiv_b200 import _abi
    _abi.lib()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.join(ROOT, "massiv_b200")
    cmd = ["gcc", "-O1", "-std=c11", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cabi_harness.c"),
           "-o", EXE, "-L", libdir, "-l:libmassiv_b200.so", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True, capture_output=True)


def test_cabi_geometry_from_plain_c():
    _build()
    r = subprocess.run([EXE, "geometry"], capture_output=True, text=True, timeout=60)
    assert "CABI_HARNESS_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cabi_compute_from_plain_c():
    _build()
    r = subprocess.run([EXE, "gpu"], capture_output=True, text=True, timeout=120)
    assert "CABI_HARNESS_OK" in r.stdout, r.stdout + r.stderr
