"""C++ ports of the reference's own test files -- /root/reference/tests/dleq_using_constraint_api.rs (constraint-system API)
and /root/reference/tests/zkp.rs (the define_proof! surface, with its bincode round trips) -- written against the host
mirror the way the Rust tests are written against zkp: compiled with g++ and linked with the built library.
CPU: they compile and link.  GPU: they run and every check passes."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORTS = ["dleq_using_constraint_api", "zkp_macro_tests"]


def _build(name):
    from zkp_b200 import build as b
    lib = b.build(force=False)
    libdir = os.path.dirname(lib)
    src = os.path.join(ROOT, "tests", "cpp", name + ".cpp")
    exe = os.path.join(ROOT, "tests", "cpp", name + ".bin")
    deps = [src, lib, os.path.join(ROOT, "zkp_b200", "csrc", "host", "toolbox.hpp"),
            os.path.join(ROOT, "include", "zkp_b200_host.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", src, "-o", exe, "-L" + libdir, "-lzkp_b200",
                               "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lcudart"])
    return exe


@pytest.mark.parametrize("name", PORTS)
def test_cpp_port_compiles_and_links(name):
    assert os.path.exists(_build(name))


@pytest.mark.gpu
@pytest.mark.parametrize("name", PORTS)
def test_cpp_port_of_reference_tests_passes(name):
    exe = _build(name)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference tests passed" in r.stdout
