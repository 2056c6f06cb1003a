"""The Rust side of the drop-in (rust/): no Rust toolchain in this image, so what can be checked is that the patch applies
cleanly to the reference tree (when /root/reference is present: this container only), that it touches exactly the files
it claims, and that every C symbol the Rust `extern "C"` block binds is declared by include/zkp_b200.h and exported."""
import os
import re
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_extern_block_matches_the_c_header():
    rs = open(os.path.join(ROOT, "rust", "src", "toolbox", "cuda_backend.rs")).read()
    block = rs[rs.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    bound = set(re.findall(r"fn (zkp_[a-z0-9_]+)\s*\(", block))
    hdr = open(os.path.join(ROOT, "include", "zkp_b200.h")).read()
    declared = set(re.findall(r"\b(zkp_[a-z0-9_]+)\s*\(", hdr))
    assert bound and bound <= declared, bound - declared
    from zkp_b200 import native
    assert bound <= set(native.SYMBOLS)


@pytest.mark.skipif(not os.path.isdir(REF) or shutil.which("patch") is None, reason="needs the reference tree and patch(1)")
def test_patch_applies_to_the_reference_tree():
    patch = os.path.join(ROOT, "rust", "zkp-cuda-backend.patch")
    touched = sorted(set(re.findall(r"^\+\+\+ b/(\S+)", open(patch).read(), flags=re.M)))
    assert touched == ["Cargo.toml", "src/toolbox/batch_verifier.rs", "src/toolbox/mod.rs", "src/toolbox/prover.rs",
                       "src/toolbox/verifier.rs"]
    with tempfile.TemporaryDirectory() as tmp:
        shutil.copytree(os.path.join(REF, "src"), os.path.join(tmp, "src"))
        shutil.copy(os.path.join(REF, "Cargo.toml"), tmp)
        r = subprocess.run(["patch", "-p1", "--dry-run", "-i", patch], cwd=tmp, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        subprocess.check_call(["patch", "-p1", "-s", "-i", patch], cwd=tmp)
        for f in touched[1:]:
            body = open(os.path.join(tmp, f)).read()
            assert 'feature = "cuda_backend"' in body, f
        # the stock build is untouched: every added line of the call sites sits behind the feature gate or is a cfg(not)
        prover = open(os.path.join(tmp, "src/toolbox/prover.rs")).read()
        assert "cuda_backend::multiscalar_mul_compressed" in prover and '#[cfg(not(feature = "cuda_backend"))]' in prover
