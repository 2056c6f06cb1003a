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
    non-canonical scalars (MalformedProof), wrong counts (VerificationFailure like the verifiers)."""
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
    # a proof with another shape inside the batch: VerificationFailure (batch_verifier.rs:142-148; BatchSizeMismatch is
    # only proofs.len() != batch_size), also when the total length happens to fit
    odd = OT.BatchableProof(proofs[0].commitments[:-1], proofs[0].responses + [5])
    mixed = OT.serialize_batchable(odd) + b"".join(OT.serialize_batchable(p) for p in proofs[1:])
    assert len(mixed) == len(blob)
    with pytest.raises(PT.VerificationFailure):
        PT.parse_batchable_many(mixed, N, k, m, threads=2)
    with pytest.raises(PT.VerificationFailure):
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


def test_compiled_transcript_script_on_the_host():
    """The symbolic STROBE that compiles a batch's transcript script for the device (api.cu bv_script, executed here on
    the host by zkp_selftest_bv_script) gives the challenge bytes of the byte-wise Merlin (host mirror AND oracle) for
    statements whose label lengths walk the per-proof values across every alignment of the 166-byte rate, including
    values split over two blocks and headers ending exactly on a block boundary."""
    import numpy as np
    lib = native.load()
    hl = PT._lib()
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    hl.zkph_transcript_export_state.argtypes = [vp, vp]

    class Desc(ctypes.Structure):
        _fields_ = [("m", ctypes.c_int32), ("ni", ctypes.c_int32), ("nc", ctypes.c_int32), ("k", ctypes.c_int32),
                    ("labels", ctypes.c_char_p), ("lhs", vp), ("cons_off", vp), ("term_scalar", vp), ("term_point", vp)]
    lib.zkp_selftest_bv_script.argtypes = [ctypes.POINTER(Desc), vp, vp, vp, vp, vp, ctypes.POINTER(ctypes.c_int32)]
    lib.zkp_selftest_bv_script.restype = ctypes.c_int32
    rnd = random.Random(77)
    seen_blocks = set()
    for trial in range(120):
        ni, nc, k = rnd.randrange(0, 5), rnd.randrange(0, 4), rnd.randrange(0, 6)
        if ni + nc == 0:
            nc = 1
        names = [("v%d" % i) + "x" * rnd.choice([0, 1, 2, 5, 17, 40, 100, 160, 170]) for i in range(ni + nc)]
        lhs = [rnd.randrange(ni + nc) for _ in range(k)]
        tl = b"T" * rnd.randrange(1, 40)
        plabel = b"proof" + b"p" * rnd.randrange(0, 30)
        secrets = ["s%d" % i for i in range(rnd.randrange(0, 4))]
        inst = [rnd.randbytes(32) for _ in range(ni)]
        comm = [rnd.randbytes(32) for _ in range(nc)]
        coms = [rnd.randbytes(32) for _ in range(k)]
        # reference: byte-wise Merlin (oracle), in the order of batch_verifier.rs / macros.rs:348-365
        t = OM.Transcript(tl)
        OT.domain_sep(t, plabel)
        for s_ in secrets:
            OT.append_scalar_var(t, s_.encode())
        # prefix state from the host mirror's transcript fed the same bytes
        h = hl.zkph_transcript_new(tl, len(tl))
        hl.zkph_transcript_append_message(h, b"dom-sep", 7, b"schnorrzkp/1.0/ristretto255", 27)
        hl.zkph_transcript_append_message(h, b"dom-sep", 7, plabel, len(plabel))
        for s_ in secrets:
            hl.zkph_transcript_append_message(h, b"scvar", 5, s_.encode(), len(s_))
        prefix = (ctypes.c_uint32 * 53)()
        hl.zkph_transcript_export_state(h, prefix)
        for i in range(ni):
            t.append_message(b"ptvar", names[i].encode())
            t.append_message(b"val", inst[i])
        for i in range(nc):
            t.append_message(b"ptvar", names[ni + i].encode())
            t.append_message(b"val", comm[i])
        for c in range(k):
            t.append_message(b"blindcom", names[lhs[c]].encode())
            t.append_message(b"val", coms[c])
        expected = t.challenge_bytes(b"chal", 64)
        hl.zkph_transcript_free(h)
        lab = b"".join(n.encode() + b"\0" for n in names)
        lhs_a = np.array(lhs, dtype=np.int32)
        off_a = np.zeros(k + 1, dtype=np.int32)
        empty = np.zeros(1, dtype=np.int32)
        d = Desc(len(secrets), ni, nc, k, lab, lhs_a.ctypes.data, off_a.ctypes.data, empty.ctypes.data, empty.ctypes.data)
        out = ctypes.create_string_buffer(64)
        nb = ctypes.c_int32(0)
        ib, cb, kb = b"".join(inst) or b"\0", b"".join(comm) or b"\0", b"".join(coms) or b"\0"
        rc = lib.zkp_selftest_bv_script(ctypes.byref(d), prefix, ib, cb, kb, out, ctypes.byref(nb))
        assert rc == 0
        assert out.raw == expected, (trial, ni, nc, k, [len(n) for n in names])
        seen_blocks.add(nb.value)
    assert len(seen_blocks) >= 4        # scripts of very different lengths were exercised


def test_inconsistent_statement_descriptions_are_refused():
    """A flattened statement is used as sizes and indices by the engine: descriptions whose ranges do not add up
    (negative counts, decreasing or non-zero-based term offsets, out-of-range lhs, missing arrays) come back as
    ZKP_ERR_SIZE -- no crash, no exception through the C ABI (api.cu statement_ok / guarded)."""
    import numpy as np
    lib = native.load()
    vp = ctypes.c_void_p

    class Desc(ctypes.Structure):
        _fields_ = [("m", ctypes.c_int32), ("ni", ctypes.c_int32), ("nc", ctypes.c_int32), ("k", ctypes.c_int32),
                    ("labels", ctypes.c_char_p), ("lhs", vp), ("cons_off", vp), ("term_scalar", vp), ("term_point", vp)]
    lib.zkp_selftest_bv_script.argtypes = [ctypes.POINTER(Desc), vp, vp, vp, vp, vp, ctypes.POINTER(ctypes.c_int32)]
    lib.zkp_selftest_bv_script.restype = ctypes.c_int32
    prefix = (ctypes.c_uint32 * 53)()
    out = ctypes.create_string_buffer(64)
    buf = b"\1" * 256
    lab = b"A\0B\0"
    lhs = np.array([0, 1], dtype=np.int32)
    ts = np.array([0, 0], dtype=np.int32)
    tp = np.array([1, 0], dtype=np.int32)

    def call(m, ni, nc, k, lhs_a, off, ts_a=ts, tp_a=tp, labels=lab):
        off_a = np.array(off, dtype=np.int32)
        d = Desc(m, ni, nc, k, labels, lhs_a.ctypes.data if lhs_a is not None else None, off_a.ctypes.data,
                 ts_a.ctypes.data if ts_a is not None else None, tp_a.ctypes.data if tp_a is not None else None)
        return lib.zkp_selftest_bv_script(ctypes.byref(d), prefix, buf, buf, buf, out, None)

    assert call(1, 1, 1, 2, lhs, [0, 1, 2]) == 0                           # the well-formed statement
    assert call(1, 1, 1, 2, lhs, [0, 2, 1]) == native.ZKP_ERR_SIZE         # decreasing term offsets
    assert call(1, 1, 1, 2, lhs, [1, 1, 2]) == native.ZKP_ERR_SIZE         # offsets not starting at 0
    assert call(1, 1, 1, 2, lhs, [0, 1, -5]) == native.ZKP_ERR_SIZE        # negative term count
    assert call(1, 1, 1, 2, lhs, [0, 1, 1 << 30]) == native.ZKP_ERR_SIZE   # absurd term count
    assert call(1, 1, 1, 2, np.array([0, 2], dtype=np.int32), [0, 1, 2]) == native.ZKP_ERR_SIZE   # lhs out of range
    assert call(1, 1, 1, 2, np.array([-1, 0], dtype=np.int32), [0, 1, 2]) == native.ZKP_ERR_SIZE
    assert call(1, 1, 1, 2, lhs, [0, 1, 2], tp_a=np.array([2, 0], dtype=np.int32)) == native.ZKP_ERR_SIZE   # term point
    assert call(1, 1, 1, 2, lhs, [0, 1, 2], ts_a=np.array([1, 0], dtype=np.int32)) == native.ZKP_ERR_SIZE   # term scalar
    assert call(1, 1, 1, 2, lhs, [0, 1, 2], ts_a=None) == native.ZKP_ERR_SIZE                               # missing array
    assert call(1, 1, 1, 2, None, [0, 1, 2]) == native.ZKP_ERR_SIZE
    assert call(1, -1, 1, 0, lhs, [0]) == native.ZKP_ERR_SIZE              # negative counts
    assert call(1, 1, 1, -2, lhs, [0]) == native.ZKP_ERR_SIZE
    assert call(1, 1, 1, 0, lhs, [0], labels=None) == native.ZKP_ERR_SIZE  # labels missing
    assert call(1, 1000, 1, 0, lhs, [0]) == native.ZKP_ERR_SIZE            # more variables than the front end holds


def test_host_statement_constructor_refuses_bad_indices():
    """zkph_statement_new checks every index of the flattened statement (host_api.cpp) and returns NULL instead of
    building a Statement the classes behind it would index out of range with."""
    import numpy as np
    hl = PT._lib()
    vp, i32 = ctypes.c_void_p, ctypes.c_int32
    labels = b"x\0A\0G\0"

    def new(ns, ni, nc, k, lhs, off, ts, tp):
        a = [np.array(v, dtype=np.int32) for v in (lhs, off, ts, tp)]
        h = hl.zkph_statement_new(b"n", b"l", labels, ns, ni, nc, k, *[x.ctypes.data for x in a])
        if h:
            hl.zkph_statement_free(h)
        return bool(h)

    assert new(1, 1, 1, 1, [0], [0, 1], [0], [1])
    assert not new(1, 1, 1, 1, [2], [0, 1], [0], [1])        # lhs beyond the points
    assert not new(1, 1, 1, 1, [0], [0, 1], [1], [1])        # secret index beyond the secrets
    assert not new(1, 1, 1, 1, [0], [0, 1], [0], [-1])       # negative point index
    assert not new(1, 1, 1, 1, [0], [1, 1], [0], [1])        # offsets not starting at 0
    assert not new(1, 1, 1, 2, [0, 0], [0, 1, 0], [0], [1])  # decreasing offsets
    assert not new(-1, 1, 1, 0, [0], [0], [0], [0])          # negative count


def test_host_input_chunk_schedule_properties():
    """The chunk schedule of the host-input pipeline (api.cu chunk_schedule, read through zkp_selftest_chunk_schedule):
    boundaries strictly increase from 0 to n, every chunk is non-empty and at most 2 * chunk_terms long, the phase
    boundary sits where the phase-1 share says, the ramped schedule starts (and ends its first phase) with chunks of
    chunk_terms / 8, and the chunk count stays within a small factor of the uniform schedule's."""
    import numpy as np
    lib = native.load()
    sz = ctypes.c_size_t
    lib.zkp_selftest_chunk_schedule.argtypes = [sz, sz, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, sz,
                                                ctypes.POINTER(sz)]
    lib.zkp_selftest_chunk_schedule.restype = ctypes.c_int64
    rnd = random.Random(11)
    cases = [(50331660, 1 << 21, 50), (1, 1024, 50), (1025, 1024, 50), (2049, 1024, 1), (4097, 1024, 100),
             (20000, 4096, 50), (20000, 5000, 50), ((1 << 31) - 2, 1 << 21, 50)]
    cases += [(rnd.randrange(1, 1 << 24), rnd.choice([1024, 4096, 5000, 1 << 16, 1 << 21]), rnd.randrange(1, 101))
              for _ in range(300)]
    for n, chunk, pct in cases:
        for ramp in (1, 0):
            cap = n // 1024 + 64
            b = np.zeros(cap, dtype=np.uint64)
            k1 = sz(0)
            cnt = lib.zkp_selftest_chunk_schedule(n, chunk, pct, ramp, b.ctypes.data, cap, ctypes.byref(k1))
            assert cnt >= 1, (n, chunk, pct, ramp)
            bb = [int(x) for x in b[:cnt + 1]]
            sizes = [bb[i + 1] - bb[i] for i in range(cnt)]
            assert bb[0] == 0 and bb[-1] == n and min(sizes) >= 1, (n, chunk, pct, ramp)
            assert 1 <= k1.value <= cnt
            uniform = (n + chunk - 1) // chunk
            if ramp and n > chunk:
                small = max(chunk // 8, 1024)
                p1 = max(1, n // 100 * pct + (n % 100) * pct // 100)
                assert bb[k1.value] == p1 and max(sizes) <= 2 * chunk
                assert sizes[0] == min(small, (p1 + 1) // 2)
                assert cnt <= uniform + 24                       # ramps add O(log) chunks, not a multiple
                if p1 >= 8 * small:
                    assert sizes[k1.value - 1] <= small          # phase 1 ends on a small chunk
                if n - p1 >= 4 * small:
                    assert sizes[k1.value] == small              # phase 2 starts on a small chunk
            else:
                assert cnt == uniform and max(sizes) <= chunk
