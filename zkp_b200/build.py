"""Build the CUDA library in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libzkp_b200.so")
SOURCES = ["api.cu", "host/merlin.cpp", "host/scalar.cpp", "host/toolbox.cpp", "host/host_api.cpp", "host/wire.cpp"]


def _headers():
    """Every header the library is built from (any change triggers a rebuild)."""
    out = []
    for d in (CSRC, os.path.join(CSRC, "host"), os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cuh", ".hpp", ".h")):
                out.append(os.path.join(d, f))
    return out


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in [os.path.join(CSRC, s) for s in SOURCES] + _headers())


def build(force=False, verbose=False, ablations=False):
    """Compile zkp_b200/csrc/*.cu -> zkp_b200/lib/libzkp_b200.so (sm_100a, -lineinfo).  ablations=True also compiles the
    measured-negative paths (-DZKP_ABLATIONS: second-stream sort, ramped chunk schedule, byte-wise STROBE front end, FP64
    field arithmetic): for re-measuring them, not for the product."""
    if not force and not _stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC,-pthread", "-shared", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if ablations:
        cmd.insert(1, "-DZKP_ABLATIONS")
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv or "--ablations" in sys.argv, verbose="-v" in sys.argv,
                ablations="--ablations" in sys.argv))
