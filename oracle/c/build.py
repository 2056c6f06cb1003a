"""Build oracle/c/ref_u64.c -> oracle/_ref/libref_u64.so (git-ignored, travels to the GPU box).
The reference itself is Rust and cannot be compiled here (no rustc/cargo; crate sources absent), so
oracle/_ref holds the C *port* of its CPU algorithms, not reference code: bench.py labels it kind="port"."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUTDIR = os.path.join(os.path.dirname(HERE), "_ref")
LIB = os.path.join(OUTDIR, "libref_u64.so")
SRC = os.path.join(HERE, "ref_u64.c")


def build(force=False):
    newest = max(os.path.getmtime(SRC), os.path.getmtime(os.path.join(HERE, "ref_ifma.h")))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    os.makedirs(OUTDIR, exist_ok=True)
    # -march=native is deliberately NOT used: the .so is built here and runs on the GPU box's host CPU
    subprocess.check_call(["gcc", "-O3", "-std=gnu11", "-fPIC", "-shared", "-pthread", "-mbmi2", "-madx",
                           "-o", LIB, SRC])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
