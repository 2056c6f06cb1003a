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


def test_wire_format_matches_oracle_and_rejects_what_bincode_rejects():
    """bincode layout of src/proofs.rs (tests/zkp.rs:53-54, :96-97): serialize / parse round trips against the oracle's
    restatement, sizes 32 + 8 + 32m and 8 + 32k + 8 + 32m, and the refusals: truncated input, trailing bytes,
    non-canonical scalars (MalformedProof), wrong counts (VerificationFailure / BatchSizeMismatch like the verifiers)."""
    import numpy as np
    import pytest
    rnd = random.Random(12)
    L = R.L
    k, m, N = 11, 21, 37
    proofs = [OT.BatchableProof([rnd.randbytes(32) for _ in range(k)], [rnd.randrange(L) for _ in range(m)]) for _ in range(N)]
    blob = b"".join(OT.serialize_batchable(p) for p in proofs)
    assert len(blob) == N * (8 + 32 * k + 8 + 32 * m)
    from oracle import scalar as S
    for p in proofs[:3]:
        mine = PT.serialize_batchable(np.frombuffer(b"".join(p.commitments), np.uint8),
                                      np.frombuffer(b"".join(S.to_bytes(r) for r in p.responses), np.uint8))
        assert mine == OT.serialize_batchable(p)
        q, off = OT.parse_batchable(mine)
        assert off == len(mine) and q.commitments == p.commitments and q.responses == p.responses
    for threads in (1, 3):
        com, resp = PT.parse_batchable_many(blob, N, k, m, threads=threads)
        for j, p in enumerate(proofs):
            assert [bytes(c) for c in com[j]] == p.commitments
            assert [int.from_bytes(bytes(r), "little") for r in resp[j]] == p.responses
    com0, resp0 = PT.parse_batchable_many(b"", 0, k, m)
    assert com0.shape == (0, k, 32) and resp0.shape == (0, m, 32)
    with pytest.raises(PT.MalformedProof):
        PT.parse_batchable_many(blob[:-1], N, k, m)                         # truncated
    with pytest.raises(PT.MalformedProof):
        PT.parse_batchable_many(blob + b"\0", N, k, m)                      # trailing byte
    bad = bytearray(blob)
    off = 5 * (8 + 32 * k + 8 + 32 * m) + 8 + 32 * k + 8 + 32 * 3           # response 3 of proof 5 := l (non-canonical)
    bad[off:off + 32] = L.to_bytes(32, "little")
    with pytest.raises(PT.MalformedProof):
        PT.parse_batchable_many(bytes(bad), N, k, m, threads=2)
    with pytest.raises(ValueError):
        OT.parse_batchable(bytes(bad), 5 * (8 + 32 * k + 8 + 32 * m))
    # a proof with another shape inside the batch: BatchSizeMismatch (batch_verifier.rs:138-148), also when the total
    # length happens to fit
    odd = OT.BatchableProof(proofs[0].commitments[:-1], proofs[0].responses + [5])
    mixed = OT.serialize_batchable(odd) + b"".join(OT.serialize_batchable(p) for p in proofs[1:])
    assert len(mixed) == len(blob)
    with pytest.raises(PT.BatchSizeMismatch):
        PT.parse_batchable_many(mixed, N, k, m, threads=2)
    with pytest.raises(PT.BatchSizeMismatch):
        PT.parse_batchable_many(blob, N, k - 1, m + 1)
    # length prefixes that exceed the input
    with pytest.raises(PT.MalformedProof):
        PT.parse_batchable_many((2**40).to_bytes(8, "little") + bytes(100), 1, k, m)
    # CompactProof
    c, rs = rnd.randrange(L), [rnd.randrange(L) for _ in range(m)]
    ser = PT.serialize_compact(np.frombuffer(S.to_bytes(c), np.uint8), np.frombuffer(b"".join(S.to_bytes(r) for r in rs), np.uint8))
    assert ser == OT.serialize_compact(c, rs) and len(ser) == 32 + 8 + 32 * m
    pc, pr = PT.parse_compact(ser, m)
    assert bytes(pc) == S.to_bytes(c) and [int.from_bytes(bytes(r), "little") for r in pr] == rs
    with pytest.raises(PT.VerificationFailure):
        PT.parse_compact(ser, m - 1)                                         # verifier.rs:82-84
    with pytest.raises(PT.VerificationFailure):
        PT.parse_compact(ser, m + 1)
    with pytest.raises(PT.MalformedProof):
        PT.parse_compact(ser[:-5], m)
    with pytest.raises(PT.MalformedProof):
        PT.parse_compact(L.to_bytes(32, "little") + ser[32:], m)             # non-canonical challenge
