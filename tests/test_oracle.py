"""CPU tests: the oracle (Python big-int and C port) against the golden vectors and libsodium."""
import hashlib
import random

import numpy as np
import pytest

from oracle import cref, merlin, msm as M, ristretto as R, scalar as S, sodium, toolbox as T
from tests import util_data as U


def _h(xs):
    return np.frombuffer(b"".join(bytes.fromhex(x) for x in xs), dtype=np.uint8).reshape(-1, 32)


def test_rfc9496_vectors():
    g = U.golden("rfc9496.json")
    for k, exp in enumerate(g["multiples_of_generator"]):
        assert R.compress(R.pt_mul(k, R.BASEPOINT)).hex() == exp
        if k:
            assert R.compress(R.decompress(bytes.fromhex(exp))).hex() == exp
    for bad in g["bad_encodings"]:
        assert R.decompress(bytes.fromhex(bad)) is None
    for v in g["hash_to_group_sha512"]:
        assert R.compress(R.hash_from_bytes_sha512(v["msg"].encode())).hex() == v["enc"]


def test_merlin_conformance():
    g = U.golden("merlin.json")
    t = merlin.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == g["simple"]
    t = merlin.Transcript(b"test protocol")
    t.append_message(b"step1", b"some data")
    for _ in range(32):
        ch = t.challenge_bytes(b"challenge", 32)
        t.append_message(b"bigdata", b"\x63" * 1024)
        t.append_message(b"challengedata", ch)
    assert ch.hex() == g["complex"]
    # keccak-f against hashlib: sha3-256 of the empty string through our permutation
    st = bytearray(200)
    st[0] ^= 0x06
    st[135] ^= 0x80
    merlin.keccak_f1600(st)
    assert bytes(st[:32]) == hashlib.sha3_256(b"").digest()


def test_libsodium_cross_check():
    na = sodium.load()
    if na is None:
        pytest.skip("libsodium with ristretto255 not present")
    rnd = random.Random(5)
    pts = U.base_points(12)
    for e in pts:
        assert sodium.is_valid_point(na, e)
    for _ in range(25):
        k = rnd.randrange(R.L)
        e = rnd.choice(pts)
        assert sodium.scalarmult(na, k, e) == R.compress(R.pt_mul(k, R.decompress(e)))
    for _ in range(300):
        b = rnd.randbytes(32)
        b = b[:31] + bytes([b[31] & 0x7F])
        assert sodium.is_valid_point(na, b) == (R.decompress(b) is not None)
    ks = [rnd.randrange(R.L) for _ in pts]
    assert sodium.msm(na, ks, pts) == R.compress(M.naive_msm(ks, [R.decompress(e) for e in pts]))


def test_dalek_algorithms_agree():
    """Straus CT, Straus NAF-5 and Pippenger (every window) compute the same group element."""
    rnd = random.Random(11)
    pts = [R.decompress(e) for e in U.base_points(20)]
    ks = [rnd.randrange(R.L) for _ in pts]
    ks[0], ks[1], ks[2] = 0, 1, R.L - 1
    exp = R.compress(M.naive_msm(ks, pts))
    assert R.compress(M.straus_ct(ks, pts)) == exp
    assert R.compress(M.straus_vartime(ks, pts)) == exp
    for w in (6, 7, 8):
        assert R.compress(M.pippenger(ks, pts, w)) == exp
    for k in ks[:6]:
        assert sum(d * 16**i for i, d in enumerate(S.to_radix_16(k))) == k
        assert sum(d << i for i, d in enumerate(S.non_adjacent_form(k, 5))) == k
        for w in (6, 7, 8):
            assert sum(d << (w * i) for i, d in enumerate(S.to_radix_2w(k, w))) == k


def test_msm_kats_python_and_c():
    for kat in U.golden("msm_kat.json")["kats"]:
        sc, pt = _h(kat["scalars"]), _h(kat["points"])
        assert cref.msm_vartime(sc, pt).hex() == kat["expected"], kat["n"]
        assert cref.msm_vartime(sc, pt, threads=3).hex() == kat["expected"], kat["n"]
        if kat["n"] <= 36:
            assert M.msm_bytes(list(sc), list(pt)).hex() == kat["expected"]
    for case in U.golden("msm_seeded.json")["cases"]:
        base = U.base_points(case["K"])
        sc = U.random_scalars(case["n"], seed=case["seed"])
        pts = np.frombuffer(b"".join(base[i % case["K"]] for i in range(case["n"])), dtype=np.uint8).reshape(-1, 32)
        assert cref.msm_vartime(sc, pts, threads=4).hex() == case["expected"]


def _needs_vector_cpu():
    if not cref.simd_available():
        pytest.skip("this CPU lacks avx512ifma/avx512vl: the vector restatement of simd_backend cannot run here")


def test_vector_backend_field_against_big_integers():
    """oracle/c/ref_ifma.h: four-lane radix-2^51 multiplication and squaring (vpmadd52) against Python integers,
    including the extreme operands p-1, p-2, 0, 1 and outputs inside the bound the next multiplication needs."""
    _needs_vector_cpu()
    P = 2**255 - 19
    rnd = random.Random(5)
    limbs = lambda x: [(x >> (51 * i)) & ((1 << 51) - 1) for i in range(5)]
    val = lambda l: sum(int(v) << (51 * i) for i, v in enumerate(l)) % P
    for t in range(600):
        a = [rnd.randrange(P) for _ in range(4)]
        b = [rnd.randrange(P) for _ in range(4)]
        if t == 0:
            a, b = [0, 1, P - 1, P - 2], [P - 1, P - 1, P - 1, 0]
        if t == 1:   # unreduced but legal inputs: every limb at 2^51 - 1 (the value 2^255 - 1 = p + 18)
            a = b = [2**255 - 1] * 4
        m, q = cref.simd_field_selftest([limbs(x) for x in a], [limbs(x) for x in b])
        for j in range(4):
            assert val(m[j]) == a[j] * b[j] % P, (t, j)
            assert val(q[j]) == a[j] * a[j] % P, (t, j)
            assert max(int(v) for v in m[j]) < 2**52 and max(int(v) for v in q[j]) < 2**52


def test_vector_backend_msm_equals_serial_port():
    """The vector restatement of the reference's simd_backend gives the bytes of the serial u64 port -- on the golden
    KATs and on seeded MSMs at every dispatch boundary (Straus < 190 <= Pippenger; window 6 / 7 / 8 at 500 / 800),
    single- and multi-threaded; an undecodable point fails the same way."""
    _needs_vector_cpu()
    for kat in U.golden("msm_kat.json")["kats"]:
        sc, pt = _h(kat["scalars"]), _h(kat["points"])
        assert cref.msm_vartime(sc, pt, simd=True).hex() == kat["expected"], kat["n"]
    base = U.base_points(64)
    for n in (0, 1, 2, 5, 64, 189, 190, 300, 499, 500, 799, 800, 3000):
        sc = U.random_scalars(n, seed=n + 1)
        pts = np.frombuffer(b"".join(base[(7 * i + n) % 64] for i in range(n)), dtype=np.uint8).reshape(-1, 32)
        want = cref.msm_vartime(sc, pts)
        assert cref.msm_vartime(sc, pts, simd=True) == want, n
        assert cref.msm_vartime(sc, pts, threads=3, simd=True) == want, n
    for case in U.golden("msm_seeded.json")["cases"]:
        sc = U.random_scalars(case["n"], seed=case["seed"])
        bp = U.base_points(case["K"])
        pts = np.frombuffer(b"".join(bp[i % case["K"]] for i in range(case["n"])), dtype=np.uint8).reshape(-1, 32)
        assert cref.msm_vartime(sc, pts, threads=4, simd=True).hex() == case["expected"]
    bad = np.array(pts[:300])
    bad[123] = 0xFF
    assert cref.msm_vartime(sc[:300], bad, simd=True) is None


def test_c_port_codec_and_batched():
    g = U.golden("rfc9496.json")
    good = [bytes.fromhex(x) for x in g["multiples_of_generator"]] + U.base_points(30)
    bad = [bytes.fromhex(x) for x in g["bad_encodings"]]
    limbs, valid = cref.decompress(good + bad)
    assert valid[:len(good)].all() and not valid[len(good):].any()
    assert [bytes(b) for b in cref.compress(limbs[:len(good)])] == good
    assert cref.msm_vartime(U.random_scalars(3, 1), [good[3], bad[0], good[4]]) is None
    kats = [k for k in U.golden("msm_kat.json")["kats"] if k["n"] <= 36]
    sc = np.concatenate([_h(k["scalars"]) for k in kats])
    pt = np.concatenate([_h(k["points"]) for k in kats])
    off = np.cumsum([0] + [k["n"] for k in kats]).astype(np.uint64)
    assert [bytes(o).hex() for o in cref.msm_ct_batched(sc, pt, off)] == [k["expected"] for k in kats]
    out, valid = cref.msm_vartime_batched(sc, pt, off)
    assert valid.all() and [bytes(o).hex() for o in out] == [k["expected"] for k in kats]


def test_toolbox_golden_dleq():
    """The oracle's Prover reproduces the committed DLEQ proof bytes; verification accepts / rejects."""
    k = U.golden("toolbox_kat.json")["dleq_compact"]
    G = R.BASEPOINT
    H = R.hash_from_bytes_sha512(R.compress(G))
    x = k["x"]
    A, B = R.pt_mul(x, G), R.pt_mul(x, H)
    tr = merlin.Transcript(b"DLEQTest")
    pr = T.Prover(b"DLEQProof", tr)
    vx = pr.allocate_scalar(b"x", x)
    vG, _ = pr.allocate_point(b"G", G)
    vH, _ = pr.allocate_point(b"H", H)
    vA, _ = pr.allocate_point(b"A", A)
    vB, _ = pr.allocate_point(b"B", B)
    T.dleq_statement(pr, vx, vA, vB, vG, vH)
    chal, resp, coms, _ = pr._prove_impl(T.SeededRng(k["rng_seed"].encode()))
    assert S.to_bytes(chal).hex() == k["challenge"]
    assert [S.to_bytes(r).hex() for r in resp] == k["responses"]
    assert [c.hex() for c in coms] == k["commitments"]

    def verifier():
        v = T.Verifier(b"DLEQProof", merlin.Transcript(b"DLEQTest"))
        sx = v.allocate_scalar(b"x")
        pG = v.allocate_point(b"G", bytes.fromhex(k["G"]))
        pH = v.allocate_point(b"H", bytes.fromhex(k["H"]))
        pA = v.allocate_point(b"A", bytes.fromhex(k["A"]))
        pB = v.allocate_point(b"B", bytes.fromhex(k["B"]))
        T.dleq_statement(v, sx, pA, pB, pG, pH)
        return v
    verifier().verify_compact(T.CompactProof(chal, resp))
    verifier().verify_batchable(T.BatchableProof(coms, resp), T.SeededRng(b"w"))
    with pytest.raises(T.VerificationFailure):
        verifier().verify_compact(T.CompactProof(chal, [(resp[0] + 1) % R.L]))
    with pytest.raises(T.VerificationFailure):
        verifier().verify_batchable(T.BatchableProof(coms, [(resp[0] + 1) % R.L]), T.SeededRng(b"w"))


def test_toolbox_cmz_batch_and_negative_controls():
    rng = T.SeededRng(b"cmz")
    st = T.CMZ10
    pts_common = {n: R.from_uniform_bytes(rng.bytes(64)) for n in st.common}
    proofs, encs = [], {n: [] for n in st.instance}
    N = 3
    for j in range(N):
        sec = {n: int.from_bytes(rng.bytes(64), "little") % R.L for n in st.secrets}
        P = R.from_uniform_bytes(rng.bytes(64))
        Q = R.from_uniform_bytes(rng.bytes(64))
        pts = dict(pts_common)
        pts["P"], pts["Q"] = P, Q
        for i in range(1, 11):
            pts["C_%d" % i] = M.naive_msm([sec["m_%d" % i], sec["z_%d" % i]], [P, pts["A"]])
        pts["V"] = M.naive_msm([sec["m_%d" % i] for i in range(1, 11)] + [sec["minus_z_Q"]],
                               [pts["X_%d" % i] for i in range(1, 11)] + [Q])
        pr, e = st.prove_batchable(merlin.Transcript(b"CMZ"), sec, pts, rng)
        st.verify_batchable(pr, merlin.Transcript(b"CMZ"), e, rng)
        proofs.append(pr)
        for n in st.instance:
            encs[n].append(e[n])
    for n in st.common:
        encs[n] = e[n]
    st.batch_verify(proofs, [merlin.Transcript(b"CMZ") for _ in range(N)], encs, rng)
    # layout of the combined MSM: 12 + 24*N terms (batch_verifier.rs:219-228)
    bv = st.build_batch_verifier(N, [merlin.Transcript(b"CMZ") for _ in range(N)], encs)
    scal, pe = bv.batch_coeffs(proofs, rng)
    assert len(scal) == len(pe) == 12 + 24 * N
    # one bad response anywhere rejects the batch
    proofs[1].responses[7] = (proofs[1].responses[7] + 1) % R.L
    with pytest.raises(T.VerificationFailure):
        st.batch_verify(proofs, [merlin.Transcript(b"CMZ") for _ in range(N)], encs, rng)
    # size mismatch and identity-encoding errors (batch_verifier.rs:72-74,120-122; toolbox/mod.rs:191)
    with pytest.raises(T.BatchSizeMismatch):
        T.BatchVerifier(b"x", 2, [merlin.Transcript(b"t")])
    bv = T.BatchVerifier(b"x", 1, [merlin.Transcript(b"t")])
    with pytest.raises(T.BatchSizeMismatch):
        bv.allocate_instance_point(b"A", [bytes(32), bytes(32)])
    with pytest.raises(T.VerificationFailure):
        bv.allocate_static_point(b"A", bytes(32))
