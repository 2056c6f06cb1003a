"""GPU parity of the host mirror (C++ toolbox over the CUDA engine) against the oracle's restatement of the
reference's Prover / Verifier / BatchVerifier: identical injected randomness => identical bytes."""
import os

import numpy as np
import pytest

from oracle import merlin as OM, msm as M, ristretto as R, scalar as S, toolbox as OT
from tests import util_data as U
from zkp_b200 import toolbox as PT

pytestmark = pytest.mark.gpu


def limbs(pt):
    return [((c % R.P) >> (51 * j)) & ((1 << 51) - 1) for c in pt for j in range(5)]


def sbytes(xs):
    return np.frombuffer(b"".join(S.to_bytes(x) for x in xs), dtype=np.uint8).reshape(-1, 32)


def dleq_instance(j=0):
    G = R.BASEPOINT
    H = R.hash_from_bytes_sha512(R.compress(G))
    x = 89327492234 + j
    return x, dict(A=R.pt_mul(x, G), B=R.pt_mul(x, H), H=H, G=G)


def test_dleq_prove_verify_matches_oracle(engine):
    """BASELINE configs[0] (DLEQ 1 prove + 1 verify) through the product path, byte-equal to the oracle."""
    st, ost = PT.dleq_statement(), OT.DLEQ
    x, pts = dleq_instance()
    order = st.points
    plimbs = np.array([limbs(pts[n]) for n in order], dtype=np.uint64)
    for batchable in (False, True):
        seed = b"seed-%d" % batchable
        if batchable:
            (com, resp), enc = st.prove_batchable(engine, b"DLEQTest", sbytes([x]), plimbs, seed)
            op, oenc = ost.prove_batchable(OM.Transcript(b"DLEQTest"), dict(x=x), pts, OT.SeededRng(seed))
            assert [bytes(c) for c in com] == op.commitments
        else:
            (chal, resp), enc = st.prove_compact(engine, b"DLEQTest", sbytes([x]), plimbs, seed)
            op, oenc = ost.prove_compact(OM.Transcript(b"DLEQTest"), dict(x=x), pts, OT.SeededRng(seed))
            assert chal == S.to_bytes(op.challenge)
        assert [bytes(r) for r in resp] == [S.to_bytes(r) for r in op.responses]
        assert [bytes(e) for e in enc] == [oenc[n] for n in order]
        if batchable:
            st.verify_batchable(engine, (com, resp), b"DLEQTest", enc, b"w")
            bad = resp.copy()
            bad[0, 0] ^= 1
            with pytest.raises(PT.VerificationFailure):
                st.verify_batchable(engine, (com, bad), b"DLEQTest", enc, b"w")
            with pytest.raises(PT.VerificationFailure):      # wrong domain separator
                st.verify_batchable(engine, (com, resp), b"DLEQTesu", enc, b"w")
        else:
            st.verify_compact(engine, (chal, resp), b"DLEQTest", enc)
            bad = resp.copy()
            bad[0, 0] ^= 1
            with pytest.raises(PT.VerificationFailure):
                st.verify_compact(engine, (chal, bad), b"DLEQTest", enc)
            wrong = enc.copy()
            wrong[0] = enc[1]                                   # wrong public point
            with pytest.raises(PT.VerificationFailure):
                st.verify_compact(engine, (chal, resp), b"DLEQTest", wrong)
            with pytest.raises(PT.VerificationFailure):      # identity encoding rejected at allocation (mod.rs:191)
                wrong[0] = 0
                st.verify_compact(engine, (chal, resp), b"DLEQTest", wrong)
            with pytest.raises(PT.VerificationFailure):      # undecodable point (verifier.rs:92)
                wrong[0] = np.frombuffer(bytes.fromhex(U.golden("rfc9496.json")["bad_encodings"][6]), dtype=np.uint8)
                st.verify_compact(engine, (chal, resp), b"DLEQTest", wrong)


def _cmz_instances(N, seed, fresh_common=False):
    """fresh_common is accepted for readability at the call site: the common points are derived from `seed` anyway, so
    a different seed gives different common points."""
    rng = OT.SeededRng(seed)
    ost = OT.CMZ10
    common = {n: R.from_uniform_bytes(rng.bytes(64)) for n in ost.common}
    secs, ptss = [], []
    for _ in range(N):
        sec = {n: int.from_bytes(rng.bytes(64), "little") % R.L for n in ost.secrets}
        P, Q = R.from_uniform_bytes(rng.bytes(64)), R.from_uniform_bytes(rng.bytes(64))
        pts = dict(common)
        pts["P"], pts["Q"] = P, Q
        for i in range(1, 11):
            pts["C_%d" % i] = M.naive_msm([sec["m_%d" % i], sec["z_%d" % i]], [P, pts["A"]])
        pts["V"] = M.naive_msm([sec["m_%d" % i] for i in range(1, 11)] + [sec["minus_z_Q"]],
                               [pts["X_%d" % i] for i in range(1, 11)] + [Q])
        secs.append(sec)
        ptss.append(pts)
    return secs, ptss


def test_serialized_proofs_round_trip_like_the_reference_tests(engine):
    """/root/reference/tests/zkp.rs:31-112 and :115-175 with the wire step they contain: prove, `bincode::serialize`,
    `deserialize`, verify -- compact, batchable, and a batch of serialized proofs parsed straight into the SoA arrays of
    the batch verifier (host path and device front end)."""
    st = PT.dleq_statement()
    x, pts = dleq_instance()
    plimbs = np.array([limbs(pts[n]) for n in st.points], dtype=np.uint64)
    (chal, resp), enc = st.prove_compact(engine, b"DLEQTest", sbytes([x]), plimbs, b"s1")
    wire = PT.serialize_compact(np.frombuffer(chal, np.uint8), resp)
    assert len(wire) == 32 + 8 + 32                                     # m + 1 32-byte elements + the Vec length
    c2, r2 = PT.parse_compact(wire, st.m)
    st.verify_compact(engine, (bytes(c2), r2), b"DLEQTest", enc)
    (com, resp), enc = st.prove_batchable(engine, b"DLEQTest", sbytes([x]), plimbs, b"s2")
    wire = PT.serialize_batchable(com, resp)
    assert len(wire) == 8 + 32 * 2 + 8 + 32
    cm, rs = PT.parse_batchable_many(wire, 1, st.k, st.m)
    st.verify_batchable(engine, (cm[0], rs[0]), b"DLEQTest", enc, b"w")
    # a stream of serialized CMZ proofs -> batch verification
    cst = PT.cmz10_statement()
    N = 6
    secs, ptss = _cmz_instances(N, b"cmz-wire")
    sec_arr = np.stack([sbytes([s_[n] for n in cst.secrets]) for s_ in secs])
    pts_arr = np.array([[limbs(p[n]) for n in cst.points] for p in ptss], dtype=np.uint64)
    entropy = np.frombuffer(OT.SeededRng(b"entropy-wire").bytes(32 * N), dtype=np.uint8).reshape(N, 32)
    enc, com, resp = cst.prove_many_device(engine, b"CMZ", sec_arr, pts_arr, entropy)
    stream = b"".join(PT.serialize_batchable(com[j], resp[j]) for j in range(N))
    assert len(stream) == N * (8 + 32 * 11 + 8 + 32 * 21)
    com2, resp2 = PT.parse_batchable_many(stream, N, cst.k, cst.m, threads=2)
    assert (com2 == com).all() and (resp2 == resp).all()
    ni = len(cst.instance)
    inst = np.ascontiguousarray(enc[:, :ni].transpose(1, 0, 2))
    cst.batch_verify(engine, com2, resp2, b"CMZ", inst, enc[0, ni:], b"rho", threads=2)
    cst.batch_verify_device(engine, com2, resp2, b"CMZ", inst, enc[0, ni:], bytes(range(32)))
    tampered = bytearray(stream)
    tampered[3 * (8 + 32 * 11 + 8 + 32 * 21) + 8 + 5] ^= 1                  # a commitment byte of proof 3
    com3, resp3 = PT.parse_batchable_many(bytes(tampered), N, cst.k, cst.m)
    with pytest.raises(PT.VerificationFailure):
        cst.batch_verify_device(engine, com3, resp3, b"CMZ", inst, enc[0, ni:], bytes(range(32)))


def test_cmz_prove_many_and_batch_verify_match_oracle(engine):
    st, ost = PT.cmz10_statement(), OT.CMZ10
    N = 5
    secs, ptss = _cmz_instances(N, b"cmz-gpu")
    sec_arr = np.stack([sbytes([s[n] for n in st.secrets]) for s in secs])
    pts_arr = np.array([[limbs(p[n]) for n in st.points] for p in ptss], dtype=np.uint64)
    entropy = np.frombuffer(OT.SeededRng(b"entropy").bytes(32 * N), dtype=np.uint8).reshape(N, 32)
    enc, com, resp = st.prove_many(engine, b"CMZ", sec_arr, pts_arr, entropy, threads=3)
    oproofs = []
    for j in range(N):
        class OneShot:            # feeds the oracle prover the same 32 entropy bytes
            def __init__(self, b): self.b = b
            def bytes(self, n): return self.b
        op, oenc = ost.prove_batchable(OM.Transcript(b"CMZ"), secs[j], ptss[j], OneShot(entropy[j].tobytes()))
        assert [bytes(c) for c in com[j]] == op.commitments
        assert [bytes(r) for r in resp[j]] == [S.to_bytes(r) for r in op.responses]
        assert [bytes(e) for e in enc[j]] == [oenc[n] for n in st.points]
        oproofs.append(op)
    # single-proof paths on proof 0
    st.verify_batchable(engine, (com[0], resp[0]), b"CMZ", enc[0], b"rho")
    # batch verification: the MSM inputs equal the oracle's (same rho stream), and the verdicts agree
    ni = len(st.instance)
    inst = np.ascontiguousarray(enc[:, :ni].transpose(1, 0, 2))
    comm = enc[0, ni:]
    co, po, _ = st.batch_verify(engine, com, resp, b"CMZ", inst, comm, b"batch-rho", threads=2, want_msm_inputs=True)
    oencs = {n: [bytes(enc[j, i]) for j in range(N)] for i, n in enumerate(st.instance)}
    for i, n in enumerate(st.common):
        oencs[n] = bytes(comm[i])
    bv = ost.build_batch_verifier(N, [OM.Transcript(b"CMZ") for _ in range(N)], oencs)
    oscal, opts = bv.batch_coeffs(oproofs, OT.SeededRng(b"batch-rho"))
    assert [bytes(c) for c in co] == [S.to_bytes(s) for s in oscal]
    assert [bytes(p) for p in po] == opts
    bad = resp.copy()
    bad[3, 7, 0] ^= 1
    with pytest.raises(PT.VerificationFailure):
        st.batch_verify(engine, com, bad, b"CMZ", inst, comm, b"batch-rho")
    with pytest.raises(PT.BatchSizeMismatch):
        st.batch_verify(engine, com, resp[:4], b"CMZ", inst, comm, b"batch-rho")
    badc = com.copy()
    badc[2, 4] = 0                                                  # identity commitment (mod.rs:215)
    with pytest.raises(PT.VerificationFailure):
        st.batch_verify(engine, badc, resp, b"CMZ", inst, comm, b"batch-rho")


def test_device_merlin_selftest(engine):
    """Keccak-f / STROBE / Merlin on the device reproduce Merlin's published conformance vector."""
    import ctypes
    buf = ctypes.create_string_buffer(32)
    assert engine._lib.zkp_selftest_hash(engine._ctx, buf) == 0
    assert buf.raw.hex() == U.golden("merlin.json")["complex"]


def test_batch_verify_device_front_end_matches_oracle(engine):
    """zkp_batch_verify_proofs (transcripts, challenges, weights, coefficient fold on the GPU): the MSM inputs it
    builds are byte-equal to the oracle's BatchVerifier with the same per-proof weight derivation."""
    st, ost = PT.cmz10_statement(), OT.CMZ10
    N = 6
    secs, ptss = _cmz_instances(N, b"cmz-dev")
    sec_arr = np.stack([sbytes([s[n] for n in st.secrets]) for s in secs])
    pts_arr = np.array([[limbs(p[n]) for n in st.points] for p in ptss], dtype=np.uint64)
    entropy = np.frombuffer(OT.SeededRng(b"entropy2").bytes(32 * N), dtype=np.uint8).reshape(N, 32)
    enc, com, resp = st.prove_many(engine, b"CMZ", sec_arr, pts_arr, entropy, threads=2)
    ni = len(st.instance)
    inst = np.ascontiguousarray(enc[:, :ni].transpose(1, 0, 2))
    comm = enc[0, ni:]
    seed = bytes(range(32))
    co, po = st.batch_verify_device(engine, com, resp, b"CMZ", inst, comm, seed, want_msm_inputs=True)
    oproofs = [OT.BatchableProof([bytes(c) for c in com[j]], [int.from_bytes(bytes(r), "little") for r in resp[j]])
               for j in range(N)]
    oencs = {n: [bytes(enc[j, i]) for j in range(N)] for i, n in enumerate(st.instance)}
    for i, n in enumerate(st.common):
        oencs[n] = bytes(comm[i])
    bv = ost.build_batch_verifier(N, [OM.Transcript(b"CMZ") for _ in range(N)], oencs)
    oscal, opts = bv.batch_coeffs(oproofs, OT.PerProofRng(seed))
    assert [bytes(p) for p in po] == opts
    assert [bytes(c) for c in co] == [S.to_bytes(s) for s in oscal]
    # verdicts
    bad = resp.copy()
    bad[2, 5, 0] ^= 1
    with pytest.raises(PT.VerificationFailure):
        st.batch_verify_device(engine, com, bad, b"CMZ", inst, comm, seed)
    with pytest.raises(PT.VerificationFailure):           # wrong transcript label
        st.batch_verify_device(engine, com, resp, b"CMY", inst, comm, seed)
    badc = com.copy()
    badc[4, 3] = 0                                          # identity commitment (mod.rs:215)
    with pytest.raises(PT.VerificationFailure):
        st.batch_verify_device(engine, badc, resp, b"CMZ", inst, comm, seed)
    noncanon = resp.copy()
    noncanon[1, 0] = 0xFF                                   # response >= l
    with pytest.raises(PT.VerificationFailure):
        st.batch_verify_device(engine, com, noncanon, b"CMZ", inst, comm, seed)
    # DLEQ (static lhs-free statement with a static rhs point) through the same path
    dst, odst = PT.dleq_statement(), OT.DLEQ
    G = R.BASEPOINT
    H = R.hash_from_bytes_sha512(R.compress(G))
    Nd = 9
    xs = [89327492234 + j for j in range(Nd)]
    dl = np.array([[limbs(R.pt_mul(x, G)), limbs(R.pt_mul(x, H)), limbs(H), limbs(G)] for x in xs], dtype=np.uint64)
    ent = np.frombuffer(OT.SeededRng(b"e3").bytes(32 * Nd), dtype=np.uint8).reshape(Nd, 32)
    denc, dcom, dresp = dst.prove_many(engine, b"DLEQBatchTest", np.stack([sbytes([x]) for x in xs]), dl, ent, threads=2)
    dinst = np.ascontiguousarray(denc[:, :3].transpose(1, 0, 2))
    co, po = dst.batch_verify_device(engine, dcom, dresp, b"DLEQBatchTest", dinst, denc[0, 3:], seed, want_msm_inputs=True)
    oproofs = [OT.BatchableProof([bytes(c) for c in dcom[j]], [int.from_bytes(bytes(r), "little") for r in dresp[j]])
               for j in range(Nd)]
    oencs = {"A": [bytes(e) for e in denc[:, 0]], "B": [bytes(e) for e in denc[:, 1]], "H": [bytes(e) for e in denc[:, 2]],
             "G": bytes(denc[0, 3])}
    bv = odst.build_batch_verifier(Nd, [OM.Transcript(b"DLEQBatchTest") for _ in range(Nd)], oencs)
    oscal, opts = bv.batch_coeffs(oproofs, OT.PerProofRng(seed))
    assert [bytes(p) for p in po] == opts and [bytes(c) for c in co] == [S.to_bytes(s) for s in oscal]


def test_batch_verify_device_slab_pipeline_and_phase_splits(engine):
    """zkp_batch_verify_proofs over several slabs: whatever the split of the rows between the two ingestion phases (three
    term ranges per thread, digit-only launches for rows nothing is left to hide under) and wherever the front-end kernel
    runs (its own high-priority stream next to the previous slab's decompression, or the compute stream), the MSM inputs are
    those of the single-slab call, the batch is accepted, and a tampered proof in the LAST slab is caught."""
    from tools import workloads as WL
    rng = np.random.default_rng(20261017)
    N = 2500
    st, sec, lim, enc = WL.cmz_instances(engine, N, rng)
    ent = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    _, com, resp = st.prove_many_device(engine, b"CMZ", sec, lim, ent)
    ni = len(st.instance)
    inst = np.ascontiguousarray(enc[:, :ni].transpose(1, 0, 2))
    comm = np.ascontiguousarray(enc[0, ni:])
    seed = bytes(range(7, 39))
    bad = resp.copy()
    bad[N - 3, 20, 2] ^= 0x10
    try:
        engine.set_option("bv_chunk_terms", 1 << 30)              # one slab
        co0, po0 = st.batch_verify_device(engine, com, resp, b"CMZ", inst, comm, seed, want_msm_inputs=True)
        co0, po0 = co0.copy(), po0.copy()
        engine.set_option("bv_chunk_terms", 1024)                 # 1024-proof slabs: 1024 + 1024 + 452
        # rows1: 0 = the split chosen per slab from the arrival of the copies, -1 = alternating half / all, else fixed
        for rows1, prep, cap in ((0, 1, 64), (0, 0, 64), (-1, 1, 64), (-1, 0, 64), (1, 1, 0), (7, 1, 64), (8, 0, 64), (12, 1, 32),
                                 (16, 1, 64), (20, 1, 64), (24, 1, 64)):
            engine.set_option("bv_phase1_rows", rows1)
            engine.set_option("bv_prep_stream", prep)
            engine.set_option("bv_prep_smem_kb", cap)
            co, po = st.batch_verify_device(engine, com, resp, b"CMZ", inst, comm, seed, want_msm_inputs=True)
            assert (co == co0).all() and (po == po0).all(), (rows1, prep)
            with pytest.raises(PT.VerificationFailure):
                st.batch_verify_device(engine, com, bad, b"CMZ", inst, comm, seed)
    finally:
        for key, v in (("bv_chunk_terms", 3 << 17), ("bv_phase1_rows", 0), ("bv_prep_stream", 1), ("bv_prep_smem_kb", 64)):
            engine.set_option(key, v)


def test_prove_many_device_front_end_matches_host_and_oracle(engine):
    """zkp_prove_batch (allocate_point compressions, transcript replay, TranscriptRng blindings, constant-time MSMs,
    challenge and responses on the GPU) is byte-identical to the host mirror and to the oracle prover given the same
    entropy: CMZ (11 constraints, 31 terms), DLEQ, a statement with a static lhs, N = 1, and an empty batch."""
    st, ost = PT.cmz10_statement(), OT.CMZ10
    N = 7
    secs, ptss = _cmz_instances(N, b"cmz-pv")
    sec_arr = np.stack([sbytes([s[n] for n in st.secrets]) for s in secs])
    pts_arr = np.array([[limbs(p[n]) for n in st.points] for p in ptss], dtype=np.uint64)
    entropy = np.frombuffer(OT.SeededRng(b"entropy-pv").bytes(32 * N), dtype=np.uint8).reshape(N, 32)
    enc_h, com_h, resp_h = st.prove_many(engine, b"CMZ", sec_arr, pts_arr, entropy, threads=2)
    enc_d, com_d, resp_d = st.prove_many_device(engine, b"CMZ", sec_arr, pts_arr, entropy)
    assert (enc_d == enc_h).all() and (com_d == com_h).all() and (resp_d == resp_h).all()

    class OneShot:
        def __init__(self, b): self.b = b
        def bytes(self, n): return self.b
    for j in (0, N - 1):
        op, oenc = ost.prove_batchable(OM.Transcript(b"CMZ"), secs[j], ptss[j], OneShot(entropy[j].tobytes()))
        assert [bytes(c) for c in com_d[j]] == op.commitments
        assert [bytes(r) for r in resp_d[j]] == [S.to_bytes(r) for r in op.responses]
        assert [bytes(e) for e in enc_d[j]] == [oenc[n] for n in st.points]
    # the device-made proofs verify (single and batch), a different transcript label gives different proofs
    st.verify_batchable(engine, (com_d[3], resp_d[3]), b"CMZ", enc_d[3], b"rho")
    ni = len(st.instance)
    st.batch_verify_device(engine, com_d, resp_d, b"CMZ", np.ascontiguousarray(enc_d[:, :ni].transpose(1, 0, 2)), enc_d[0, ni:],
                           bytes(range(32)))
    _, com_x, _ = st.prove_many_device(engine, b"CMY", sec_arr, pts_arr, entropy)
    assert (com_x != com_d).any()
    # shared constant-time tables of the batch-static points on / off give the same bytes; a batch whose "common" points
    # differ between proofs (legal: every proof carries its own assignments) falls back to per-proof tables
    engine.set_option("share_static_tables", 0)
    try:
        e_n, c_n, r_n = st.prove_many_device(engine, b"CMZ", sec_arr, pts_arr, entropy)
    finally:
        engine.set_option("share_static_tables", 1)
    assert (e_n == enc_d).all() and (c_n == com_d).all() and (r_n == resp_d).all()
    secs2, ptss2 = _cmz_instances(2, b"cmz-pv-other-common", fresh_common=True)
    mixed_pts = pts_arr.copy()
    mixed_pts[N - 1] = np.array([limbs(ptss2[0][n]) for n in st.points], dtype=np.uint64)
    mixed_sec = sec_arr.copy()
    mixed_sec[N - 1] = sbytes([secs2[0][n] for n in st.secrets])
    e_m, c_m, r_m = st.prove_many_device(engine, b"CMZ", mixed_sec, mixed_pts, entropy)
    e_h, c_h, r_h = st.prove_many(engine, b"CMZ", mixed_sec, mixed_pts, entropy, threads=2)
    assert (e_m == e_h).all() and (c_m == c_h).all() and (r_m == r_h).all()
    st.verify_batchable(engine, (c_m[N - 1], r_m[N - 1]), b"CMZ", e_m[N - 1], b"rho")
    # large batches run as slices (here: 3 proofs per slice, a ragged last slice): same bytes
    engine.set_option("prove_chunk", 3)
    try:
        e_s, c_s, r_s = st.prove_many_device(engine, b"CMZ", sec_arr, pts_arr, entropy)
    finally:
        engine.set_option("prove_chunk", 1 << 17)
    assert (e_s == enc_d).all() and (c_s == com_d).all() and (r_s == resp_d).all()
    # N = 1 and the empty batch
    e1, c1, r1 = st.prove_many_device(engine, b"CMZ", sec_arr[:1], pts_arr[:1], entropy[:1])
    assert (e1 == enc_h[:1]).all() and (c1 == com_h[:1]).all() and (r1 == resp_h[:1]).all()
    e0, c0, r0 = st.prove_many_device(engine, b"CMZ", sec_arr[:0], pts_arr[:0], entropy[:0])
    assert e0.shape[0] == 0 and c0.shape[0] == 0 and r0.shape[0] == 0
    # a non-canonical secret is refused
    bad = sec_arr.copy()
    bad[2, 4] = 0xFF
    with pytest.raises(Exception):
        st.prove_many_device(engine, b"CMZ", bad, pts_arr, entropy)
    # DLEQ and a statement with a static left-hand side
    dst = PT.dleq_statement()
    G = R.BASEPOINT
    H = R.hash_from_bytes_sha512(R.compress(G))
    xs = [89327492234 + j for j in range(9)]
    dl = np.array([[limbs(R.pt_mul(x, G)), limbs(R.pt_mul(x, H)), limbs(H), limbs(G)] for x in xs], dtype=np.uint64)
    ent = np.frombuffer(OT.SeededRng(b"e-pv").bytes(32 * 9), dtype=np.uint8).reshape(9, 32)
    dsec = np.stack([sbytes([x]) for x in xs])
    a = dst.prove_many(engine, b"DLEQBatchTest", dsec, dl, ent, threads=2)
    b = dst.prove_many_device(engine, b"DLEQBatchTest", dsec, dl, ent)
    assert all((x == y).all() for x, y in zip(a, b))
    pk = PT.Statement("pk", "PK proof", ["x"], ["A"], ["G", "Q"], [("A", [("x", "G")]), ("Q", [("x", "G")])])
    Q = R.pt_mul(123456789, G)
    pts = np.array([[limbs(Q), limbs(G), limbs(Q)] for _ in range(4)], dtype=np.uint64)
    ent = np.frombuffer(OT.SeededRng(b"pk-pv").bytes(32 * 4), dtype=np.uint8).reshape(4, 32)
    a = pk.prove_many(engine, b"PK", np.stack([sbytes([123456789])] * 4), pts, ent, threads=1)
    b = pk.prove_many_device(engine, b"PK", np.stack([sbytes([123456789])] * 4), pts, ent)
    assert all((x == y).all() for x, y in zip(a, b))


def test_prove_many_device_comb_path_matches_straus_path(engine):
    """Option prove_comb (signed four-tooth combs, comb.cuh) gives the bytes of the default Straus path of
    zkp_prove_batch: CMZ with 70 proofs (three interleave groups, the last one partial), shared and per-proof combs for
    the batch-static points, a batch whose common points differ (fallback), slices, N = 1, DLEQ and a static-lhs statement."""
    st = PT.cmz10_statement()
    N = 70
    secs, ptss = _cmz_instances(N, b"cmz-comb")
    sec_arr = np.stack([sbytes([s[n] for n in st.secrets]) for s in secs])
    pts_arr = np.array([[limbs(p[n]) for n in st.points] for p in ptss], dtype=np.uint64)
    entropy = np.frombuffer(OT.SeededRng(b"entropy-comb").bytes(32 * N), dtype=np.uint8).reshape(N, 32)
    engine.set_option("prove_comb", 0)       # the reference bytes come from the Straus path (k_small_msm_ct)
    want = st.prove_many_device(engine, b"CMZ", sec_arr, pts_arr, entropy)
    secs2, ptss2 = _cmz_instances(1, b"cmz-comb-other-common", fresh_common=True)
    mixed_pts, mixed_sec = pts_arr.copy(), sec_arr.copy()
    mixed_pts[N - 1] = np.array([limbs(ptss2[0][n]) for n in st.points], dtype=np.uint64)
    mixed_sec[N - 1] = sbytes([secs2[0][n] for n in st.secrets])
    want_mixed = st.prove_many_device(engine, b"CMZ", mixed_sec, mixed_pts, entropy)
    dst = PT.dleq_statement()
    G = R.BASEPOINT
    H = R.hash_from_bytes_sha512(R.compress(G))
    xs = [89327492234 + j for j in range(9)]
    dl = np.array([[limbs(R.pt_mul(x, G)), limbs(R.pt_mul(x, H)), limbs(H), limbs(G)] for x in xs], dtype=np.uint64)
    ent = np.frombuffer(OT.SeededRng(b"e-comb").bytes(32 * 9), dtype=np.uint8).reshape(9, 32)
    dsec = np.stack([sbytes([x]) for x in xs])
    want_dleq = dst.prove_many_device(engine, b"DLEQBatchTest", dsec, dl, ent)
    pk = PT.Statement("pk", "PK proof", ["x"], ["A"], ["G", "Q"], [("A", [("x", "G")]), ("Q", [("x", "G")])])
    Q = R.pt_mul(123456789, G)
    pkp = np.array([[limbs(Q), limbs(G), limbs(Q)] for _ in range(4)], dtype=np.uint64)
    pke = np.frombuffer(OT.SeededRng(b"pk-comb").bytes(32 * 4), dtype=np.uint8).reshape(4, 32)
    want_pk = pk.prove_many_device(engine, b"PK", np.stack([sbytes([123456789])] * 4), pkp, pke)
    same = lambda a, b: all((x == y).all() for x, y in zip(a, b))
    try:
        # 1 = combs scanned from global memory (k_small_msm_comb); 2 = combs staged in shared memory, one CTA per 32 proofs,
        # constraints cut into units of <= prove_piece terms (k_comb_msm_cta: the default)
        for comb, piece in ((1, 2), (2, 2), (2, 1), (2, 3), (2, 11)):
            engine.set_option("prove_comb", comb)
            engine.set_option("prove_piece", piece)
            for share in (1, 0):
                engine.set_option("share_static_tables", share)
                tag = (comb, piece, share)
                assert same(st.prove_many_device(engine, b"CMZ", sec_arr, pts_arr, entropy), want), tag
                assert same(st.prove_many_device(engine, b"CMZ", mixed_sec, mixed_pts, entropy), want_mixed), tag
                assert same(st.prove_many_device(engine, b"CMZ", sec_arr[:1], pts_arr[:1], entropy[:1]), [w[:1] for w in want]), tag
                assert same(dst.prove_many_device(engine, b"DLEQBatchTest", dsec, dl, ent), want_dleq), tag
                assert same(pk.prove_many_device(engine, b"PK", np.stack([sbytes([123456789])] * 4), pkp, pke), want_pk), tag
            engine.set_option("share_static_tables", 1)
            engine.set_option("prove_chunk", 33)
            assert same(st.prove_many_device(engine, b"CMZ", sec_arr, pts_arr, entropy), want), (comb, piece)
            engine.set_option("prove_chunk", 1 << 17)
            # the pipelined prover: slices of 16 proofs (the last one ragged) alternating between two workspaces and
            # streams; a batch whose common points differ in the LAST slice is redone without sharing
            engine.set_option("prove_pipe_chunk", 16)
            assert same(st.prove_many_device(engine, b"CMZ", sec_arr, pts_arr, entropy), want), (comb, piece)
            assert same(st.prove_many_device(engine, b"CMZ", mixed_sec, mixed_pts, entropy), want_mixed), (comb, piece)
            engine.set_option("prove_pipe_chunk", 1 << 14)
    finally:
        engine.set_option("prove_comb", 2)
        engine.set_option("prove_piece", 2)
        engine.set_option("share_static_tables", 1)
        engine.set_option("prove_chunk", 1 << 17)
        engine.set_option("prove_pipe_chunk", 1 << 14)


def test_compiled_transcript_script_equals_bytewise_strobe(engine):
    """k_bv_prepare2 (host-compiled STROBE framing, block-wise absorb) against k_bv_prepare (byte-wise STROBE on the
    device) and the oracle: identical MSM inputs for CMZ, DLEQ and for statements whose label lengths walk the per-proof
    values across every alignment and block boundary of the 166-byte rate (values split over two blocks, headers ending
    exactly on a boundary)."""
    G = R.BASEPOINT
    seed = bytes(range(7, 39))
    from zkp_b200 import EngineError
    try:   # the byte-wise kernel is an ablation: in the library only when built with -DZKP_ABLATIONS
        engine.set_option("bv_compiled", 0)
        bytewise = True
    except EngineError:
        bytewise = False
    engine.set_option("bv_compiled", 1)

    def both(st, N, tl, com, resp, inst, comm):
        outs = []
        for compiled in ((1, 0) if bytewise else (1,)):
            engine.set_option("bv_compiled", compiled)
            try:
                outs.append(st.batch_verify_device(engine, com, resp, tl, inst, comm, seed, want_msm_inputs=True))
            finally:
                engine.set_option("bv_compiled", 1)
        if bytewise:
            assert (outs[0][0] == outs[1][0]).all() and (outs[0][1] == outs[1][1]).all()
        return outs[0]
    # CMZ
    st, ost = PT.cmz10_statement(), OT.CMZ10
    N = 5
    secs, ptss = _cmz_instances(N, b"cmz-script")
    sec_arr = np.stack([sbytes([s_[n] for n in st.secrets]) for s_ in secs])
    pts_arr = np.array([[limbs(p[n]) for n in st.points] for p in ptss], dtype=np.uint64)
    entropy = np.frombuffer(OT.SeededRng(b"entropy-script").bytes(32 * N), dtype=np.uint8).reshape(N, 32)
    enc, com, resp = st.prove_many_device(engine, b"CMZ", sec_arr, pts_arr, entropy)
    ni = len(st.instance)
    co, po = both(st, N, b"CMZ", com, resp, np.ascontiguousarray(enc[:, :ni].transpose(1, 0, 2)), enc[0, ni:])
    oproofs = [OT.BatchableProof([bytes(c) for c in com[j]], [int.from_bytes(bytes(r), "little") for r in resp[j]])
               for j in range(N)]
    oencs = {n: [bytes(enc[j, i]) for j in range(N)] for i, n in enumerate(st.instance)}
    for i, n in enumerate(st.common):
        oencs[n] = bytes(enc[0, ni + i])
    bv = ost.build_batch_verifier(N, [OM.Transcript(b"CMZ") for _ in range(N)], oencs)
    oscal, opts = bv.batch_coeffs(oproofs, OT.PerProofRng(seed))
    assert [bytes(p) for p in po] == opts and [bytes(c) for c in co] == [S.to_bytes(s_) for s_ in oscal]
    # label lengths 1 .. 60 and transcript labels of several lengths: every alignment of the values in the rate block
    x = 987654321
    A = R.pt_mul(x, G)
    for L in list(range(1, 61, 3)) + [166, 167, 200]:
        names = ["A" * L, "G" + "g" * (L // 2), "Q" + "q" * (L % 7)]
        pk = PT.Statement("pk", "PK proof " + "p" * (L % 11), ["x"], [names[0]], [names[1], names[2]],
                          [(names[0], [("x", names[1])]), (names[2], [("x", names[1])])])
        opk = OT.Statement("pk", "PK proof " + "p" * (L % 11), ["x"], [names[0]], [names[1], names[2]],
                           [(names[0], [("x", names[1])]), (names[2], [("x", names[1])])])
        Np = 3
        pts = np.array([[limbs(A), limbs(G), limbs(A)] for _ in range(Np)], dtype=np.uint64)
        ent = np.frombuffer(OT.SeededRng(b"pk-script%d" % L).bytes(32 * Np), dtype=np.uint8).reshape(Np, 32)
        tl = b"T" * (1 + L % 5)
        e2, c2, r2 = pk.prove_many_device(engine, tl, np.stack([sbytes([x])] * Np), pts, ent)
        inst = np.ascontiguousarray(e2[:, :1].transpose(1, 0, 2))
        co, po = both(pk, Np, tl, c2, r2, inst, e2[0, 1:])
        oproofs = [OT.BatchableProof([bytes(c) for c in c2[j]], [int.from_bytes(bytes(r), "little") for r in r2[j]])
                   for j in range(Np)]
        oencs = {names[0]: [bytes(e) for e in e2[:, 0]], names[1]: bytes(e2[0, 1]), names[2]: bytes(e2[0, 2])}
        bv = opk.build_batch_verifier(Np, [OM.Transcript(tl) for _ in range(Np)], oencs)
        oscal, opts = bv.batch_coeffs(oproofs, OT.PerProofRng(seed))
        assert [bytes(p) for p in po] == opts and [bytes(c) for c in co] == [S.to_bytes(s_) for s_ in oscal], L


def test_single_verdict_mode_over_shards(engine):
    """SURVEY 8e, single-verdict mode: ONE BatchVerifier batch (one weight stream, one static coefficient vector) cut by
    proofs into shards; the shards' partial sums (zkp_batch_verify_partial) are individually NOT the identity when the
    static coefficients all travel with shard 0, their sum is (zkp_partials_verdict) -- exactly the reference's verdict."""
    st = PT.cmz10_statement()
    N = 9
    secs, ptss = _cmz_instances(N, b"cmz-shards")
    sec_arr = np.stack([sbytes([s_[n] for n in st.secrets]) for s_ in secs])
    pts_arr = np.array([[limbs(p[n]) for n in st.points] for p in ptss], dtype=np.uint64)
    entropy = np.frombuffer(OT.SeededRng(b"entropy-shards").bytes(32 * N), dtype=np.uint8).reshape(N, 32)
    enc, com, resp = st.prove_many_device(engine, b"CMZ", sec_arr, pts_arr, entropy)
    ni, nc = len(st.instance), len(st.common)
    inst = np.ascontiguousarray(enc[:, :ni].transpose(1, 0, 2))
    co, po, _ = st.batch_verify(engine, com, resp, b"CMZ", inst, enc[0, ni:], b"one-rho-stream", threads=2, want_msm_inputs=True)
    rows = ni + st.k
    sc_static, sc_inst = co[:nc], co[nc:].reshape(rows, N, 32)
    pt_static, pt_inst = po[:nc], po[nc:].reshape(rows, N, 32)
    from zkp_b200 import parallel
    for world in (1, 2, 3):
        partials = []
        for rank in range(world):
            lo, hi = parallel.shard_range(N, rank, world)
            s_sc = sc_static if rank == 0 else sc_static[:0]
            s_pt = pt_static if rank == 0 else pt_static[:0]
            part = engine.batch_verify_partial(s_sc, s_pt, np.ascontiguousarray(sc_inst[:, lo:hi]).reshape(-1, 32),
                                               np.ascontiguousarray(pt_inst[:, lo:hi]).reshape(-1, 32), rows, hi - lo)
            assert part is not None and part.shape == (4, 5)
            partials.append(part)
        ok, enc_sum = engine.partials_verdict(np.stack(partials))
        assert ok and enc_sum == bytes(32)
        if world > 1:
            alone, e0 = engine.partials_verdict(partials[0])
            assert not alone and e0 != bytes(32)        # shard 0 carries all static coefficients: not the identity by itself
    # a tampered coefficient in one shard voids the single verdict; an invalid encoding is reported by its shard
    bad = sc_inst.copy()
    bad[3, N - 1, 0] ^= 1
    lo, hi = parallel.shard_range(N, 1, 2)
    p0 = engine.batch_verify_partial(sc_static, pt_static, np.ascontiguousarray(sc_inst[:, :lo]).reshape(-1, 32),
                                     np.ascontiguousarray(pt_inst[:, :lo]).reshape(-1, 32), rows, lo)
    p1 = engine.batch_verify_partial(sc_static[:0], pt_static[:0], np.ascontiguousarray(bad[:, lo:hi]).reshape(-1, 32),
                                     np.ascontiguousarray(pt_inst[:, lo:hi]).reshape(-1, 32), rows, hi - lo)
    ok, _ = engine.partials_verdict(np.stack([p0, p1]))
    assert not ok
    badp = pt_inst.copy()
    badp[2, lo] = 0xFF
    assert engine.batch_verify_partial(sc_static[:0], pt_static[:0], np.ascontiguousarray(sc_inst[:, lo:hi]).reshape(-1, 32),
                                       np.ascontiguousarray(badp[:, lo:hi]).reshape(-1, 32), rows, hi - lo) is None
    # the empty shard contributes the identity
    pe = engine.batch_verify_partial(sc_static[:0], pt_static[:0], sc_static[:0], pt_static[:0], rows, 0)
    ok, _ = engine.partials_verdict(np.stack([p0, pe]))
    assert not ok                                        # p0 alone is only part of the batch
    ok, _ = engine.partials_verdict(pe)
    assert ok


def test_device_front_end_edge_shapes(engine):
    """N = 1, an empty batch, and a statement whose constraint has a STATIC lhs (static_coeffs path of
    batch_verifier.rs:187-189) through both the host mirror and the device front end."""
    st = PT.Statement("pk", "PK proof", ["x"], ["A"], ["G", "Q"], [("A", [("x", "G")]), ("Q", [("x", "G")])])
    ost = OT.Statement("pk", "PK proof", ["x"], ["A"], ["G", "Q"], [("A", [("x", "G")]), ("Q", [("x", "G")])])
    G = R.BASEPOINT
    x = 123456789
    Q = R.pt_mul(x, G)                  # the same secret for every proof, so the static lhs Q is consistent
    seed = bytes(range(1, 33))
    for N in (1, 4):
        pts = np.array([[limbs(Q), limbs(G), limbs(Q)] for _ in range(N)], dtype=np.uint64)
        ent = np.frombuffer(OT.SeededRng(b"pk%d" % N).bytes(32 * N), dtype=np.uint8).reshape(N, 32)
        enc, com, resp = st.prove_many(engine, b"PK", np.stack([sbytes([x])] * N), pts, ent, threads=1)
        inst = np.ascontiguousarray(enc[:, :1].transpose(1, 0, 2))
        st.batch_verify(engine, com, resp, b"PK", inst, enc[0, 1:], b"r")
        co, po = st.batch_verify_device(engine, com, resp, b"PK", inst, enc[0, 1:], seed, want_msm_inputs=True)
        oproofs = [OT.BatchableProof([bytes(c) for c in com[j]], [int.from_bytes(bytes(r), "little") for r in resp[j]])
                   for j in range(N)]
        oencs = {"A": [bytes(e) for e in enc[:, 0]], "G": bytes(enc[0, 1]), "Q": bytes(enc[0, 2])}
        bv = ost.build_batch_verifier(N, [OM.Transcript(b"PK") for _ in range(N)], oencs)
        oscal, opts = bv.batch_coeffs(oproofs, OT.PerProofRng(seed))
        assert [bytes(p) for p in po] == opts and [bytes(c) for c in co] == [S.to_bytes(s) for s in oscal]
        bad = resp.copy()
        bad[0, 0, 1] ^= 4
        with pytest.raises(PT.VerificationFailure):
            st.batch_verify_device(engine, com, bad, b"PK", inst, enc[0, 1:], seed)
    # empty batch: the sum over no proofs is the identity -> accepted, like the reference's empty MSM
    st.batch_verify_device(engine, np.zeros((0, 2, 32), np.uint8), np.zeros((0, 1, 32), np.uint8), b"PK",
                           np.zeros((1, 0, 32), np.uint8), enc[0, 1:], seed)


def test_dleq_batch_golden(engine):
    """tests/zkp.rs:115-175 shape: 4 macro-form DLEQ proofs, golden fixture produced by the oracle."""
    kb = U.golden("toolbox_kat.json")["dleq_batch"]
    st = PT.dleq_statement()
    h = lambda xs: np.frombuffer(b"".join(bytes.fromhex(x) for x in xs), dtype=np.uint8).reshape(-1, 32)
    com = np.stack([h(p["commitments"]) for p in kb["proofs"]])
    resp = np.stack([h(p["responses"]) for p in kb["proofs"]])
    inst = np.stack([h(kb["A"]), h(kb["B"]), h(kb["H"])])
    comm = h([kb["G"]])
    co, po, _ = st.batch_verify(engine, com, resp, b"DLEQBatchTest", inst, comm, kb["verify_rng_seed"].encode(),
                                want_msm_inputs=True)
    assert [bytes(c).hex() for c in co] == kb["msm_scalars"]
    assert [bytes(p).hex() for p in po] == kb["msm_points"]
