"""The C++ port of the reference's own constraint-API test file (/root/reference/tests/dleq_using_constraint_api.rs),
written against the host mirror's classes the way the Rust test is written against zkp's: compiled with g++ and linked
with the built library.  CPU: it compiles and links.  GPU: it runs and every check passes."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "dleq_using_constraint_api.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "dleq_using_constraint_api.bin")


def _build():
    from zkp_b200 import build as b
    lib = b.build(force=False)
    libdir = os.path.dirname(lib)
    deps = [SRC, lib, os.path.join(ROOT, "zkp_b200", "csrc", "host", "toolbox.hpp")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", SRC, "-o", EXE, "-L" + libdir, "-lzkp_b200",
                               "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lcudart"])
    return EXE


def test_cpp_port_compiles_and_links():
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_cpp_port_of_reference_constraint_api_tests_passes():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference tests passed" in r.stdout
