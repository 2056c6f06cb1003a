"""CPU tests of the C++ host mirror's primitives (Merlin, scalar arithmetic mod l, the SHAKE rng) against the
oracle.  The flows that touch the GPU are in tests/test_gpu_toolbox.py."""
import ctypes
import os
import random
import re

from oracle import merlin as OM, ristretto as R, toolbox as OT
from tests import util_data as U
from zkp_b200 import native, toolbox as PT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_header_symbols_exported():
    lib = native.load()
    hdr = open(os.path.join(ROOT, "include", "zkp_b200_host.h")).read()
    declared = set(re.findall(r"\b(zkph_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), name


def test_merlin_scalar_rng():
    lib = PT._lib()
    o = ctypes.create_string_buffer(32)
    lib.zkph_merlin_test_vector(o)
    assert o.raw.hex() == U.golden("merlin.json")["complex"]
    rnd = random.Random(4)
    L = R.L
    for _ in range(300):
        a, b = rnd.getrandbits(256), rnd.getrandbits(256)
        lib.zkph_scalar_mul(o, a.to_bytes(32, "little"), b.to_bytes(32, "little"))
        assert int.from_bytes(o.raw, "little") == (a % L) * (b % L) % L
        w = rnd.getrandbits(512)
        lib.zkph_scalar_from_wide(o, w.to_bytes(64, "little"))
        assert int.from_bytes(o.raw, "little") == w % L
    for w in (2**512 - 1, (L << 256) - 1, L << 200, 0, L, L - 1):
        lib.zkph_scalar_from_wide(o, w.to_bytes(64, "little"))
        assert int.from_bytes(o.raw, "little") == w % L
    buf = ctypes.create_string_buffer(500)
    lib.zkph_rng_bytes(b"seed-x", 6, buf, 500)
    assert buf.raw == OT.SeededRng(b"seed-x").bytes(500)


def test_statement_descriptor_matches_oracle():
    st = PT.cmz10_statement()
    assert (st.m, st.p, st.k) == (21, 25, 11)
    assert st.secrets == OT.CMZ10.secrets and st.instance == OT.CMZ10.instance and st.common == OT.CMZ10.common
    assert st.constraints == OT.CMZ10.constraints and st.label.encode() == OT.CMZ10.label
    d = PT.dleq_statement()
    assert d.constraints == OT.DLEQ.constraints and d.label.encode() == OT.DLEQ.label
