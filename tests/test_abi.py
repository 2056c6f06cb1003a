"""The C-ABI library loads and exports every symbol include/zkp_b200.h declares (no compute without a GPU)."""
import os
import re

from zkp_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = native.load()
    hdr = open(os.path.join(ROOT, "include", "zkp_b200.h")).read()
    declared = set(re.findall(r"\b(zkp_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("zkp_ctx")
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export: " + name
    assert declared == set(native.SYMBOLS)


def test_fails_loudly_without_gpu_or_library(monkeypatch):
    import pytest
    from zkp_b200 import Engine, EngineError
    lib = native.load()
    if lib.zkp_device_count() == 0:
        with pytest.raises(EngineError):
            Engine(0)
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "LIB_PATH", "/nonexistent/libzkp_b200.so")
    with pytest.raises(native.NativeLibraryMissing):
        native.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "zkp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "libref_u64" not in src and "oracle/_ref" not in src and "oracle." not in src, f
