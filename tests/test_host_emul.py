"""The device arithmetic headers (fe.cuh / ge.cuh / sc.cuh) compiled for the HOST with -DZKP_HOST_EMUL
(asm carry chains replaced by 64-bit C) and checked against the big-int oracle: validates the kernel logic
without a GPU.  The emulation library is test infrastructure only."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import ristretto as R

HERE = os.path.dirname(os.path.abspath(__file__))
P = R.P


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "host_emul", "emul.cpp")
    out = os.path.join(HERE, "host_emul", "libemul.so")
    deps = [src] + [os.path.join(HERE, "..", "zkp_b200", "csrc", f) for f in ("fe.cuh", "fe64.cuh", "ge.cuh", "sc.cuh", "hash.cuh", "scl.cuh", "comb.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-DZKP_HOST_EMUL", "-DZKP_ABLATIONS", "-x", "c++", src, "-o", out])
    return ctypes.CDLL(out)


def _c2(f, a, b):
    r = ctypes.create_string_buffer(32)
    f(r, a.to_bytes(32, "little"), b.to_bytes(32, "little"))
    return int.from_bytes(r.raw, "little")


def _c1(f, a):
    r = ctypes.create_string_buffer(32)
    f(r, a.to_bytes(32, "little"))
    return int.from_bytes(r.raw, "little")


def test_field_ops(lib):
    rnd = random.Random(1)
    edge = [0, 1, 2, 19, 38, P - 1, P, P + 1, 2**255 - 1, 2**255, 2**256 - 1, 2**256 - 38, 2**256 - 39, 2 * P, 2 * P + 1,
            2**32 - 1, 2**64 - 1, (2**256 - 1) // 3, 2**224 - 1, 2**256 - 2**224]
    vals = edge + [rnd.getrandbits(256) for _ in range(150)] + [rnd.getrandbits(256) | (2**256 - 2**200) for _ in range(30)]
    for a in vals:
        assert _c1(lib.emul_fe_sq, a) % P == a * a % P
        assert _c1(lib.emul_fe_sq_vt, a) == _c1(lib.emul_fe_sq, a)
        assert _c1(lib.emul_fe_canon, a) == a % P
        for b in rnd.sample(vals, 8) + edge:
            assert _c2(lib.emul_fe_mul, a, b) % P == a * b % P
            assert _c2(lib.emul_fe_add, a, b) % P == (a + b) % P
            assert _c2(lib.emul_fe_sub, a, b) % P == (a - b) % P
            # the variable-time tails return the very same 256-bit representative as the constant-time ones
            assert _c2(lib.emul_fe_mul_vt, a, b) == _c2(lib.emul_fe_mul, a, b)
            assert _c2(lib.emul_fe_add_vt, a, b) == _c2(lib.emul_fe_add, a, b)
            assert _c2(lib.emul_fe_sub_vt, a, b) == _c2(lib.emul_fe_sub, a, b)
    # operands that drive the cold branches of the tails: a carry / borrow leaving limb 0, rippling through all-ones or
    # all-zero limbs, and wrapping past 2^256 / below 0 a second time
    M = 2**256
    cold = [(M - 1, 1), (M - 1, M - 1), (M - 38, 38), (M - 37, 75), (M - 1, 38), (2**255, 2**255), (M - 2**32, 2**32 + 5),
            (0, 1), (0, M - 1), (37, 38), (5, M - 20), (2**32, 2**32 + 1), (2**224, 2**224 + 1), (0, 2**255), (1, 39),
            (M - 19, M - 19), (M - 2**32 + 1, 2**32 - 1 + 38)]
    for a, b in cold + [(b, a) for a, b in cold]:
        for f_vt, f_ct, op in ((lib.emul_fe_add_vt, lib.emul_fe_add, lambda x, y: x + y),
                               (lib.emul_fe_sub_vt, lib.emul_fe_sub, lambda x, y: x - y),
                               (lib.emul_fe_mul_vt, lib.emul_fe_mul, lambda x, y: x * y)):
            r = _c2(f_vt, a, b)
            assert r == _c2(f_ct, a, b) and r % P == op(a, b) % P
    # products whose first-pass sum lo + 38*hi = c8*2^256 + x has limb 0 of x within 38*c8 of 2^32: the carry out of
    # limb 0 in the tail of fe_reduce512 (probability ~1e-7 for random operands, so constructed: b = S / a mod 2p)
    M2P = M - 38
    found = 0
    for trial in range(4000):
        c8 = rnd.randrange(1, 38)
        hi_limbs = rnd.choice([0, M - 2**32, rnd.getrandbits(224) << 32])   # all-zero / all-ones / random upper limbs
        x = hi_limbs + 2**32 - 1 - rnd.randrange(38 * c8)
        S = c8 * M + x
        a = rnd.getrandbits(256) | 1
        if a % P == 0:
            continue
        b = S * pow(a, -1, M2P) % M2P
        t = a * b
        lo9 = (t % M) + 38 * (t >> 256)
        if (lo9 & 0xFFFFFFFF) + 38 * (lo9 >> 256) < 2**32:
            continue
        found += 1
        r = _c2(lib.emul_fe_mul_vt, a, b)
        assert r == _c2(lib.emul_fe_mul, a, b) and r % P == a * b % P
    assert found >= 20, found
    for a in vals[:40]:
        if a % P:
            assert _c1(lib.emul_fe_invert, a) * a % P == 1
        assert _c1(lib.emul_fe_pow22523, a) % P == pow(a, (P - 5) // 8, P)
        assert _c1(lib.emul_fe_pow22523_vt, a) == _c1(lib.emul_fe_pow22523, a)
    # 51-bit limb interface, including unreduced limbs up to 2^54
    for _ in range(50):
        limbs = [rnd.getrandbits(54) for _ in range(5)]
        val = sum(l << (51 * i) for i, l in enumerate(limbs))
        arr = (ctypes.c_uint64 * 5)(*limbs)
        r = ctypes.create_string_buffer(32)
        lib.emul_fe_from_limbs51(r, arr)
        assert int.from_bytes(r.raw, "little") % P == val % P
        out = (ctypes.c_uint64 * 5)()
        lib.emul_fe_to_limbs51(out, r.raw)
        assert sum(int(out[i]) << (51 * i) for i in range(5)) == val % P and all(int(o) < 2**51 for o in out)


def test_fp64_field(lib):
    """fe64.cuh: the FP64-pipe twin of the field arithmetic is exact (same residues as the integer path and the oracle),
    its limbs stay integral and balanced, and its columns stay below 2^53."""
    rnd = random.Random(11)
    lib.emul_fe64_sq.restype = ctypes.c_double
    lib.emul_fe64_mul.restype = ctypes.c_double
    lib.emul_fe64_column_bound.restype = ctypes.c_double
    M = 2**256
    edge = [0, 1, 2, 19, P - 1, P, P + 1, 2**255 - 1, 2**255, M - 1, M - 38, 2 * P, (M - 1) // 3, 2**234, 2**234 - 1,
            sum(((1 << 21) - 1) << o for o in (0, 22, 43, 64, 85, 107, 128, 149, 170, 192, 213, 234)),      # limbs at +max
            sum((1 << 21) << o for o in (0, 22, 43, 64, 85, 107, 128, 149, 170, 192, 213, 234)) % M]        # limbs at the rounding tie
    vals = edge + [rnd.getrandbits(256) for _ in range(200)]
    r = ctypes.create_string_buffer(32)
    for a in vals:
        ab = a.to_bytes(32, "little")
        lib.emul_fe64_roundtrip(r, ab)
        v = int.from_bytes(r.raw, "little")
        assert v % P == a % P and v < M
        for n in (1, 2, 7):
            mq = lib.emul_fe64_sq(r, ab, n)
            assert int.from_bytes(r.raw, "little") % P == pow(a, 2**n, P)
            assert mq <= 2**21 + 2**13          # balanced limbs (+ the slack of the two second-round carries)
        cb = lib.emul_fe64_column_bound(ab)      # 0 would also flag a limb that is not an integer multiple of its weight
        assert cb < 2**52 and (cb > 0) == (a % P != 0)
        for b in rnd.sample(vals, 6) + edge[:8]:
            mq = lib.emul_fe64_mul(r, ab, b.to_bytes(32, "little"))
            assert int.from_bytes(r.raw, "little") % P == a * b % P
            assert mq <= 2**21 + 2**13
    for a in vals[:60]:
        lib.emul_fe_pow22523_fp64(r, a.to_bytes(32, "little"))
        assert int.from_bytes(r.raw, "little") % P == pow(a, (P - 5) // 8, P)
    # a long squaring chain keeps the invariants (what the 254-squaring exponentiation relies on)
    a = rnd.getrandbits(255)
    mq = lib.emul_fe64_sq(r, a.to_bytes(32, "little"), 2000)
    assert int.from_bytes(r.raw, "little") % P == pow(a, 2**2000, P) and mq <= 2**21 + 2**13


def _pb(p):
    return b"".join((c % P).to_bytes(32, "little") for c in p)


def _pf(b):
    return tuple(int.from_bytes(b[i * 32:(i + 1) * 32], "little") % P for i in range(4))


def test_group_and_codec(lib):
    rnd = random.Random(7)
    pts = [R.from_uniform_bytes(rnd.randbytes(64)) for _ in range(24)] + [R.IDENTITY, R.BASEPOINT]
    encs = [R.compress(p) for p in pts]
    affs = []
    for p, e in zip(pts, encs):
        out = ctypes.create_string_buffer(32)
        lib.emul_encode(out, _pb(p))
        assert out.raw == e
        z = rnd.randrange(1, P)
        lib.emul_encode(out, _pb(tuple(c * z % P for c in p)))
        assert out.raw == e
        a = ctypes.create_string_buffer(96)
        assert lib.emul_decode(a, e) == 1
        d = R.decompress(e)
        assert [int.from_bytes(a.raw[i * 32:(i + 1) * 32], "little") % P for i in range(3)] == [d[0], d[1], d[3]]
        a2 = ctypes.create_string_buffer(96)
        assert lib.emul_decode_vt(a2, e) == 1 and a2.raw == a.raw
        affs.append(a.raw)
    for _ in range(300):
        e = rnd.randbytes(32)
        if rnd.random() < 0.5:
            e = e[:31] + bytes([e[31] & 0x7F])
        a = ctypes.create_string_buffer(96)
        assert lib.emul_decode(a, e) == (R.decompress(e) is not None)
        assert lib.emul_decode_vt(a, e) == (R.decompress(e) is not None)
    for i in range(len(pts) - 2):
        p, q = pts[i], pts[i + 1]
        out = ctypes.create_string_buffer(128)
        for neg in (0, 1):
            lib.emul_madd(out, _pb(p), affs[i + 1], neg)
            exp = R.pt_sub(p, q) if neg else R.pt_add(p, q)
            r = _pf(out.raw)
            assert R.on_curve(r) and R.compress(r) == R.compress(exp)
            lib.emul_madd_signed(out, _pb(p), affs[i + 1], neg)
            r = _pf(out.raw)
            assert R.on_curve(r) and R.compress(r) == R.compress(exp)
            out2 = ctypes.create_string_buffer(128)
            lib.emul_madd_signed_vt(out2, _pb(p), affs[i + 1], neg)
            assert out2.raw == out.raw
        lib.emul_add(out, _pb(p), _pb(q))
        r, exp = _pf(out.raw), R.pt_add(p, q)
        assert R.on_curve(r) and (r[0] * exp[2] - exp[0] * r[2]) % P == 0 and (r[1] * exp[2] - exp[1] * r[2]) % P == 0
        lib.emul_double(out, _pb(p))
        r, exp = _pf(out.raw), R.pt_double(p)
        assert R.on_curve(r) and (r[0] * exp[2] - exp[0] * r[2]) % P == 0 and (r[1] * exp[2] - exp[1] * r[2]) % P == 0
        lib.emul_madd(out, _pb(q), affs[i + 1], 1)
        assert lib.emul_is_identity(out.raw) == 1
    # every representative of the identity coset counts as the identity (verifier.rs:168, batch_verifier.rs:230)
    for rep in [R.IDENTITY, (0, P - 1, 1, 0), (R.SQRT_M1, 0, 1, 0), (P - R.SQRT_M1, 0, 1, 0)]:
        assert lib.emul_is_identity(_pb(rep)) == 1
    assert lib.emul_is_identity(_pb(R.BASEPOINT)) == 0


def test_scalar_recode(lib):
    rnd = random.Random(3)
    L = R.L
    for c in (4, 7, 8, 13, 16, 19, 20, 23, 24):   # 19 = the bench's window; 24 = the widest the digit extractor takes
        W = -(-253 // c)
        cases = [0, 1, 2, L - 1, L - 2, L // 2, L // 2 + 1, L // 2 - 1, 2**128, 2**252 - 1, 2**252, L, L + 1, 2**256 - 1]
        cases += [rnd.randrange(L) for _ in range(100)] + [L - rnd.getrandbits(128) for _ in range(30)]
        for s in cases:
            dg = (ctypes.c_int32 * W)()
            k32 = ctypes.create_string_buffer(32)
            neg = ctypes.c_int()
            rc = lib.emul_recode(dg, k32, ctypes.byref(neg), s.to_bytes(32, "little"), c, W)
            assert (rc & 1) == (s < L)
            if s < L:
                assert rc >> 1 == 0
                k = int.from_bytes(k32.raw, "little")
                assert k == min(s, L - s) and neg.value == (k != s)
                assert sum(d * (1 << (c * i)) for i, d in enumerate(dg)) == k
                assert all(abs(d) <= 1 << (c - 1) for d in dg)


def test_device_hash_and_scalar_logic(lib):
    """hash.cuh / scl.cuh (device Keccak, STROBE, Merlin, SHAKE weights, scalars mod l) compiled for the host."""
    import hashlib
    from tests import util_data as U
    o = ctypes.create_string_buffer(32)
    lib.emul_transcript_test(o)
    assert o.raw.hex() == U.golden("merlin.json")["complex"]
    # TranscriptRng (prover front end: rekey with witnesses, finalize with entropy, 64-byte draws) against the oracle
    from oracle import merlin as OM
    rr = random.Random(21)
    for n_w, n_out in ((0, 1), (1, 1), (3, 4), (21, 21)):
        wit = [rr.randbytes(32) for _ in range(n_w)]
        ent = rr.randbytes(32)
        t = OM.Transcript(b"test protocol")
        t.append_message(b"step1", b"some data")
        bld = t.build_rng()
        for w in wit:
            bld.rekey_with_witness_bytes(b"", w)
        rng = bld.finalize(ent)
        exp = b"".join(rng.fill_bytes(64) for _ in range(n_out))
        out = ctypes.create_string_buffer(64 * n_out)
        lib.emul_transcript_rng(out, b"".join(wit), n_w, ent, n_out)
        assert out.raw == exp, (n_w, n_out)
    for n, m in [(16, b"abc"), (176, b"seed" * 8 + b"\x01" * 8), (300, b"x" * 135), (136, b"")]:
        out = ctypes.create_string_buffer(n)
        lib.emul_shake(out, n, m, len(m))
        assert out.raw == hashlib.shake_256(m).digest(n)
    L = R.L
    rnd = random.Random(9)
    vals = [0, 1, L - 1, 2**252 - 1, 2**252, L - 2] + [rnd.randrange(L) for _ in range(150)]
    for a in vals:
        for b in rnd.sample(vals, 6):
            for f, op in ((lib.emul_scl_mul, lambda x, y: x * y), (lib.emul_scl_add, lambda x, y: x + y),
                          (lib.emul_scl_sub, lambda x, y: x - y)):
                f(o, a.to_bytes(32, "little"), b.to_bytes(32, "little"))
                assert int.from_bytes(o.raw, "little") == op(a, b) % L
    # the 128-bit-weight product of the coefficient fold (scl_mul_128: 4 x 8 limbs, short reduction)
    for a in [0, 1, 2**128 - 1, 2**127, 2**64, 2**128 - 2**64] + [rnd.getrandbits(128) for _ in range(300)]:
        for b in [0, 1, L - 1, L - 2, 2**252 - 1, 2**252, L // 2] + [rnd.randrange(L) for _ in range(4)]:
            lib.emul_scl_mul_128(o, a.to_bytes(32, "little"), b.to_bytes(32, "little"))
            assert int.from_bytes(o.raw, "little") == a * b % L, (a, b)
    for w in [0, 2**512 - 1, (L << 256) - 1, L << 200, L * L - 1] + [rnd.getrandbits(512) for _ in range(300)]:
        lib.emul_scl_wide(o, w.to_bytes(64, "little"))
        assert int.from_bytes(o.raw, "little") == w % L


def test_field_ops_property_based(lib):
    """hypothesis over the whole 256-bit input space, biased towards limb boundaries: the integer field ops (constant- and
    variable-time tails) and their FP64 twins agree with big-int arithmetic mod p."""
    from hypothesis import given, settings, strategies as st_
    lib.emul_fe64_sq.restype = ctypes.c_double
    lib.emul_fe64_mul.restype = ctypes.c_double
    limb = st_.one_of(st_.sampled_from([0, 1, 2, 18, 19, 37, 38, 2**31 - 1, 2**31, 2**32 - 39, 2**32 - 38, 2**32 - 19, 2**32 - 2,
                                        2**32 - 1]), st_.integers(0, 2**32 - 1))
    elem = st_.lists(limb, min_size=8, max_size=8).map(lambda ws: sum(w << (32 * i) for i, w in enumerate(ws)))

    @settings(max_examples=300, deadline=None)
    @given(elem, elem)
    def check(a, b):
        r = ctypes.create_string_buffer(32)
        assert _c2(lib.emul_fe_mul, a, b) % P == a * b % P
        assert _c2(lib.emul_fe_mul_vt, a, b) == _c2(lib.emul_fe_mul, a, b)
        assert _c1(lib.emul_fe_sq_vt, a) % P == a * a % P
        assert _c2(lib.emul_fe_add_vt, a, b) % P == (a + b) % P and _c2(lib.emul_fe_add_vt, a, b) == _c2(lib.emul_fe_add, a, b)
        assert _c2(lib.emul_fe_sub_vt, a, b) % P == (a - b) % P and _c2(lib.emul_fe_sub_vt, a, b) == _c2(lib.emul_fe_sub, a, b)
        assert _c1(lib.emul_fe_canon, a) == a % P
        assert lib.emul_fe64_mul(r, a.to_bytes(32, "little"), b.to_bytes(32, "little")) <= 2**21 + 2**13
        assert int.from_bytes(r.raw, "little") % P == a * b % P
        lib.emul_fe64_sq(r, a.to_bytes(32, "little"), 3)
        assert int.from_bytes(r.raw, "little") % P == pow(a, 8, P)
    check()


def test_scalar_chains_property_based(lib):
    """hypothesis over the operand space of the device scalar arithmetic (scl.cuh on the multiply-add chains of fe.cuh), biased
    towards limb boundaries: the 128-bit-weight product and its short reduction, the full product, the wide reduction of
    the challenge, addition and subtraction agree with big integers mod l."""
    from hypothesis import given, settings, strategies as st_
    L = R.L
    limb = st_.one_of(st_.sampled_from([0, 1, 2, 2**28 - 1, 2**28, 2**31 - 1, 2**31, 2**32 - 2, 2**32 - 1, 0x5cf5d3ed, 0x14def9de]),
                      st_.integers(0, 2**32 - 1))
    w128 = st_.lists(limb, min_size=4, max_size=4).map(lambda ws: sum(w << (32 * i) for i, w in enumerate(ws)))
    red = st_.lists(limb, min_size=8, max_size=8).map(lambda ws: sum(w << (32 * i) for i, w in enumerate(ws)) % L)
    wide = st_.lists(limb, min_size=16, max_size=16).map(lambda ws: sum(w << (32 * i) for i, w in enumerate(ws)))
    o = ctypes.create_string_buffer(32)

    @settings(max_examples=400, deadline=None)
    @given(w128, red, red, wide)
    def check(a, b, c, w):
        lib.emul_scl_mul_128(o, a.to_bytes(32, "little"), b.to_bytes(32, "little"))
        assert int.from_bytes(o.raw, "little") == a * b % L
        lib.emul_scl_mul(o, b.to_bytes(32, "little"), c.to_bytes(32, "little"))
        assert int.from_bytes(o.raw, "little") == b * c % L
        lib.emul_scl_add(o, b.to_bytes(32, "little"), c.to_bytes(32, "little"))
        assert int.from_bytes(o.raw, "little") == (b + c) % L
        lib.emul_scl_sub(o, b.to_bytes(32, "little"), c.to_bytes(32, "little"))
        assert int.from_bytes(o.raw, "little") == (b - c) % L
        lib.emul_scl_wide(o, w.to_bytes(64, "little"))
        assert int.from_bytes(o.raw, "little") == w % L
    check()


def test_signed_comb_of_the_batch_prover(lib):
    """comb.cuh (the four-tooth signed comb behind k_small_msm_comb): the recoding reproduces the scalar (or scalar + l
    for even scalars) as +-1 digits, the eight table entries are T_3 +- T_2 +- T_1 +- T_0, and a comb MSM -- built,
    recoded, selected and accumulated with the functions the kernel uses -- encodes to what the oracle's MSM encodes
    to, for edge scalars (0, 1, 2, l-1, l-2, 2^252, powers of two) and random ones, one to twelve terms."""
    from oracle import msm as M, scalar as S
    L = S.L
    rnd = random.Random(21)
    edge = [0, 1, 2, 3, L - 1, L - 2, 2**252, 2**251, 2**64, 2**64 - 1, 2**128 + 1, 2**192, (L - 1) // 2]
    for s in edge + [rnd.randrange(L) for _ in range(200)]:
        m = ctypes.create_string_buffer(32)
        lib.emul_comb_recode(m, s.to_bytes(32, "little"))
        mi = int.from_bytes(m.raw, "little")
        k = sum((2 * ((mi >> i) & 1) - 1) << i for i in range(256))
        assert k == (s if s & 1 else s + L) and (mi >> 255) == 1
    # table entries
    pts = [R.from_uniform_bytes(rnd.randbytes(64)) for _ in range(6)] + [R.BASEPOINT, R.IDENTITY]
    def mulp(p, k):
        return R.pt_mul(k, p)
    for p in pts[:3] + [R.IDENTITY]:
        out = ctypes.create_string_buffer(1024)
        lib.emul_comb_build(out, _pb(p))
        teeth = [mulp(p, 1 << (64 * x)) for x in range(4)]
        for idx in range(8):
            e = out.raw[128 * idx:128 * (idx + 1)]
            ypx, ymx, z, t2d = (int.from_bytes(e[32 * i:32 * (i + 1)], "little") % P for i in range(4))
            inv2 = pow(2, P - 2, P)
            x, y = (ypx - ymx) * inv2 % P, (ypx + ymx) * inv2 % P
            t = x * y % P * pow(z, P - 2, P) % P
            assert t2d == 2 * R.D * t % P
            exp = teeth[3]
            for b in range(3):
                exp = R.pt_add(exp, teeth[b]) if (idx >> b) & 1 else R.pt_sub(exp, teeth[b])
            assert R.compress((x, y, z, t)) == R.compress(exp), idx
    # whole MSMs
    for n in (1, 2, 3, 11, 12):
        for trial in range(4):
            sc = [rnd.choice(edge) if rnd.random() < 0.3 else rnd.randrange(L) for _ in range(n)]
            ps = [rnd.choice(pts) for _ in range(n)]
            out = ctypes.create_string_buffer(128)
            lib.emul_comb_msm(out, b"".join(s.to_bytes(32, "little") for s in sc), b"".join(_pb(p) for p in ps), n)
            enc = ctypes.create_string_buffer(32)
            lib.emul_encode(enc, out.raw)
            exp = M.msm_bytes([s.to_bytes(32, "little") for s in sc], [R.compress(p) for p in ps])
            assert enc.raw == exp, (n, trial)


@pytest.fixture(scope="module")
def klib():
    """The batch prover's kernels compiled for the host (tests/host_emul/cuda_shim.h) and run thread by thread."""
    from tests import util_data as U
    if not U.can_spawn_threads():
        pytest.skip("this environment does not allow ~1100 threads per process (needed to emulate 1024-thread blocks)")
    src = os.path.join(HERE, "host_emul", "kernels_emul.cpp")
    out = os.path.join(HERE, "host_emul", "libkemul.so")
    csrc = os.path.join(HERE, "..", "zkp_b200", "csrc")
    deps = [src, os.path.join(HERE, "host_emul", "cuda_shim.h")] + [os.path.join(csrc, f) for f in os.listdir(csrc)
                                                                    if f.endswith((".cuh", ".hpp"))]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-pthread",
                               "-DZKP_HOST_EMUL", "-DZKP_ABLATIONS", "-I/usr/local/cuda/include", src, "-o", out])
    return ctypes.CDLL(out)


def test_batch_prover_msm_kernels_run_on_the_host(klib):
    """k_pv_gather, the Straus tables / MSMs (k_build_tables, k_small_msm_ct) and the comb path (k_build_combs,
    k_comb_recode, k_small_msm_comb) executed thread by thread on the CPU with the launch sequence and the plan of
    zkp_prove_batch: for 40 proofs (one full interleave group of 32 and a partial one) of a statement whose bases are
    shared between constraints, used once, instance and common, every commitment equals the oracle's -- with and without
    shared tables for the batch-static points, and a batch whose "common" points differ raises the not-uniform flag."""
    import numpy as np
    from oracle import scalar as S
    rnd = random.Random(33)
    m, ni, nc = 5, 3, 3
    lhs = np.array([0, 2, 0, 2], dtype=np.int32)
    terms = [[(0, 1), (1, 3)], [(2, 1), (3, 4), (4, 5), (0, 3)], [(1, 1)], [(2, 3), (3, 2)]]
    k = len(terms)
    off = np.cumsum([0] + [len(t) for t in terms]).astype(np.int32)
    ts = np.array([s for t in terms for s, _ in t], dtype=np.int32)
    tp = np.array([q for t in terms for _, q in t], dtype=np.int32)
    N = 40
    common = [R.from_uniform_bytes(rnd.randbytes(64)) for _ in range(nc)]
    pts = [[R.from_uniform_bytes(rnd.randbytes(64)) for _ in range(ni)] + common for _ in range(N)]
    edge = [0, 1, 2, S.L - 1, S.L - 2, 2**252]
    blind = [[rnd.choice(edge) if rnd.random() < 0.15 else rnd.randrange(S.L) for _ in range(m)] for _ in range(N)]

    def limbs_of(pt):
        z = rnd.randrange(1, P)                        # any projective representative
        x, y, zz = pt[0] * z % P, pt[1] * z % P, pt[2] * z % P
        t = x * y % P * pow(zz, P - 2, P) % P
        return [(c >> (51 * i)) & ((1 << 51) - 1) for c in (x, y, zz, t) for i in range(5)]

    klib.emul_prove_msms.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p] * 4 + [ctypes.c_size_t, ctypes.c_void_p,
                                                                                ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                                                                ctypes.c_void_p]

    def run(points, share, comb, same_representative):
        if same_representative:     # the uniformity check compares limbs, so every copy of a common point is the same bytes
            fixed = {id(q): limbs_of(q) for q in common}
            rows = [[fixed.get(id(q)) or limbs_of(q) for q in row] for row in points]
        else:
            rows = [[limbs_of(q) for q in row] for row in points]
        limbs = np.array(rows, dtype=np.uint64)
        bl = b"".join(s.to_bytes(32, "little") for row in blind for s in row)
        com = ctypes.create_string_buffer(N * k * 32)
        flag = klib.emul_prove_msms(m, ni, nc, k, lhs.ctypes.data, off.ctypes.data, ts.ctypes.data, tp.ctypes.data, N,
                                    limbs.ctypes.data, bl, share, comb, ctypes.cast(com, ctypes.c_void_p))
        return flag, com.raw

    expected = b""
    for j in range(N):
        for t in terms:
            acc = R.IDENTITY
            for s, q in t:
                acc = R.pt_add(acc, R.pt_mul(blind[j][s], pts[j][q]))
            expected += R.compress(acc)
    for share in (1, 0):
        for comb, piece in ((0, 2), (1, 2), (2, 2), (2, 1), (2, 3)):   # 2 = CTA-staged comb kernel, units of <= piece terms
            klib.emul_set_piece(piece)
            flag, com = run(pts, share, comb, same_representative=bool(share))
            assert flag == 0, (share, comb, piece)
            bad = [i for i in range(N * k) if com[32 * i:32 * i + 32] != expected[32 * i:32 * i + 32]]
            assert not bad, (share, comb, piece, bad[:8])
    klib.emul_set_piece(2)
    # a "common" point that differs in one proof: the kernels say so (api.cu then redoes the call without sharing)
    odd = [list(row) for row in pts]
    odd[17][ni + 1] = R.from_uniform_bytes(rnd.randbytes(64))
    for comb in (0, 1, 2):
        flag, _ = run(odd, 1, comb, same_representative=True)
        assert flag != 0


def test_device_prover_pipeline_on_the_host_matches_the_oracle_prover(klib):
    """All of zkp_prove_batch's kernels (k_compress_limbs, k_pv_blind, k_pv_gather, the Straus or the comb MSMs,
    k_pv_finish) run thread by thread on the CPU from the same inputs as the oracle's Prover (/root/reference/src/toolbox/
    prover.rs:76-112 restated): encodings, commitments and responses of CMZ'13 proofs (11 constraints, 31 terms) and of DLEQ
    proofs are byte-identical, with shared and per-proof tables, Straus and comb; a non-canonical secret is reported."""
    import numpy as np
    from oracle import merlin as OM, msm as M, scalar as S, toolbox as OT

    class OneShot:
        def __init__(self, b): self.b = b
        def bytes(self, n): return self.b

    def limbs_of(pt):
        return [((c % P) >> (51 * j)) & ((1 << 51) - 1) for c in pt for j in range(5)]

    klib.emul_prove_batch.argtypes = ([ctypes.c_int] * 4 + [ctypes.c_char_p] + [ctypes.c_void_p] * 5 + [ctypes.c_size_t]
                                      + [ctypes.c_void_p] * 3 + [ctypes.c_int] * 2 + [ctypes.c_void_p] * 3)

    def device_prove(ost, tlabel, secs, ptss, entropy, share, comb):
        names = ost.instance + ost.common
        m, ni, nc, k, N = len(ost.secrets), len(ost.instance), len(ost.common), len(ost.constraints), len(secs)
        lhs = np.array([names.index(l) for l, _ in ost.constraints], dtype=np.int32)
        off = np.cumsum([0] + [len(r) for _, r in ost.constraints]).astype(np.int32)
        ts = np.array([ost.secrets.index(s) for _, r in ost.constraints for s, _ in r], dtype=np.int32)
        tp = np.array([names.index(q) for _, r in ost.constraints for _, q in r], dtype=np.int32)
        # the transcript every proof starts from: Transcript::new(label), dom-sep, scalar labels (macros.rs:206-214)
        t = OM.Transcript(tlabel)
        OT.domain_sep(t, ost.label)
        for s_ in ost.secrets:
            OT.append_scalar_var(t, s_.encode())
        st = t.strobe
        state = bytes(st.state)
        prefix = np.array([int.from_bytes(state[4 * i:4 * i + 4], "little") for i in range(50)]
                          + [st.pos, st.pos_begin, st.cur_flags], dtype=np.uint32)
        sec = np.frombuffer(b"".join(S.to_bytes(s[n]) if isinstance(s[n], int) else s[n] for s in secs for n in ost.secrets),
                            dtype=np.uint8).copy()
        lim = np.array([[limbs_of(pp[n]) for n in names] for pp in ptss], dtype=np.uint64)
        ent = np.frombuffer(b"".join(entropy), dtype=np.uint8).copy()
        enc, com, resp = (np.zeros((N, len(names), 32), np.uint8), np.zeros((N, k, 32), np.uint8), np.zeros((N, m, 32), np.uint8))
        rc = klib.emul_prove_batch(m, ni, nc, k, b"".join(n.encode() + b"\0" for n in names), lhs.ctypes.data, off.ctypes.data,
                                   ts.ctypes.data, tp.ctypes.data, prefix.ctypes.data, N, sec.ctypes.data, lim.ctypes.data,
                                   ent.ctypes.data, share, comb, enc.ctypes.data, com.ctypes.data, resp.ctypes.data)
        return rc, enc, com, resp

    rng = OT.SeededRng(b"emul-prover")
    ost = OT.CMZ10
    common = {n: R.from_uniform_bytes(rng.bytes(64)) for n in ost.common}
    secs, ptss = [], []
    N = 3
    for _ in range(N):
        sec = {n: int.from_bytes(rng.bytes(64), "little") % S.L for n in ost.secrets}
        Pp, Q = R.from_uniform_bytes(rng.bytes(64)), R.from_uniform_bytes(rng.bytes(64))
        pts = dict(common)
        pts["P"], pts["Q"] = Pp, Q
        for i in range(1, 11):
            pts["C_%d" % i] = M.naive_msm([sec["m_%d" % i], sec["z_%d" % i]], [Pp, pts["A"]])
        pts["V"] = M.naive_msm([sec["m_%d" % i] for i in range(1, 11)] + [sec["minus_z_Q"]],
                               [pts["X_%d" % i] for i in range(1, 11)] + [Q])
        secs.append(sec)
        ptss.append(pts)
    entropy = [rng.bytes(32) for _ in range(N)]
    want = [ost.prove_batchable(OM.Transcript(b"CMZ"), secs[j], ptss[j], OneShot(entropy[j])) for j in range(N)]
    for share, comb in ((1, 0), (1, 1), (0, 1)):
        rc, enc, com, resp = device_prove(ost, b"CMZ", secs, ptss, entropy, share, comb)
        assert rc == 0, (share, comb)
        for j in range(N):
            proof, oenc = want[j]
            assert [bytes(c) for c in com[j]] == proof.commitments, (share, comb, j)
            assert [bytes(r) for r in resp[j]] == [S.to_bytes(r) for r in proof.responses], (share, comb, j)
            assert [bytes(e) for e in enc[j]] == [oenc[n] for n in ost.instance + ost.common]
    # DLEQ (tests/zkp.rs:28 form: H is an instance variable), comb path
    G = R.BASEPOINT
    H = R.hash_from_bytes_sha512(R.compress(G))
    xs = [89327492234 + j for j in range(2)]
    dsecs = [{"x": x} for x in xs]
    dpts = [{"A": R.pt_mul(x, G), "B": R.pt_mul(x, H), "H": H, "G": G} for x in xs]
    dent = [rng.bytes(32) for _ in xs]
    for comb in (0, 1):
        rc, enc, com, resp = device_prove(OT.DLEQ, b"DLEQTest", dsecs, dpts, dent, 1, comb)
        assert rc == 0
        for j in range(2):
            proof, _ = OT.DLEQ.prove_batchable(OM.Transcript(b"DLEQTest"), dsecs[j], dpts[j], OneShot(dent[j]))
            assert [bytes(c) for c in com[j]] == proof.commitments and [bytes(r) for r in resp[j]] == [S.to_bytes(r) for r in proof.responses]
    # a non-canonical secret comes back as ZKP_ERR_SCALAR
    bad = [dict(d) for d in dsecs]
    bad[1]["x"] = b"\xff" * 32
    rc, *_ = device_prove(OT.DLEQ, b"DLEQTest", bad, dpts, dent, 1, 0)
    assert rc == 3


def test_small_msm_kernels_on_the_host_against_the_golden_kats(klib):
    """The batched small-MSM kernels behind verify_compact and the single-proof prover -- k_decompress_valid,
    k_prep_scalars_vt, k_small_msm_vt (variable time, signed binary digits of k and 3k) and k_decompress_ext,
    k_build_tables, k_small_msm_ct (constant time, radix 16) -- run thread by thread on the CPU over the golden MSM KATs as
    one CSR batch: every result equals the KAT (edge scalars, identity sums, sizes 1..36); an undecodable point and a
    non-canonical scalar are reported per MSM."""
    import numpy as np
    from tests import util_data as U
    kats = [k_ for k_ in U.golden("msm_kat.json")["kats"] if k_["n"] <= 36]
    sc = b"".join(bytes.fromhex(x) for k_ in kats for x in k_["scalars"])
    pt = b"".join(bytes.fromhex(x) for k_ in kats for x in k_["points"])
    off = np.cumsum([0] + [k_["n"] for k_ in kats]).astype(np.uint64)
    M = len(kats)
    out = ctypes.create_string_buffer(M * 32)
    status = np.zeros(M, dtype=np.int32)
    klib.emul_msm_vartime_batched.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_void_p]
    klib.emul_msm_ct_batched.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                         ctypes.c_void_p]
    _vt, _ct = klib.emul_msm_vartime_batched, klib.emul_msm_ct_batched
    klib_vt = lambda s_, p_, o_, m_, out_, st_: _vt(s_, p_, o_, m_, 0, out_, st_)
    klib_ct = lambda s_, p_, o_, m_, out_: _ct(s_, p_, o_, m_, 0, out_)
    klib_vt(sc, pt, off.ctypes.data, M, ctypes.cast(out, ctypes.c_void_p), status.ctypes.data)
    assert (status == 0).all()
    assert [out.raw[32 * j:32 * j + 32].hex() for j in range(M)] == [k_["expected"] for k_ in kats]
    out2 = ctypes.create_string_buffer(M * 32)
    assert klib_ct(sc, pt, off.ctypes.data, M, ctypes.cast(out2, ctypes.c_void_p)) == 0
    assert out2.raw == out.raw
    # the four-lanes-per-MSM schedules (ge_double_coop4 / ge_madd_coop4 / ge_add_pniels_coop4: shuffles between the lanes
    # of a group, several groups with different term counts per warp) on the first MSMs of the batch
    Mc = min(M, 9)
    oc, stc = ctypes.create_string_buffer(Mc * 32), np.zeros(Mc, dtype=np.int32)
    _vt(sc, pt, off.ctypes.data, Mc, 1, ctypes.cast(oc, ctypes.c_void_p), stc.ctypes.data)
    assert (stc == 0).all() and oc.raw == out.raw[:32 * Mc]
    assert _ct(sc, pt, off.ctypes.data, Mc, 1, ctypes.cast(oc, ctypes.c_void_p)) == 0 and oc.raw == out.raw[:32 * Mc]
    # failures are local to their MSM (vartime) / reported with the entry point's codes (constant time)
    victim = next(j for j, k_ in enumerate(kats) if k_["n"] >= 2)
    t0 = int(off[victim])
    bad_pt = bytearray(pt); bad_pt[32 * t0:32 * t0 + 32] = b"\xff" * 32
    bad_sc = bytearray(sc); bad_sc[32 * (t0 + 1):32 * (t0 + 1) + 32] = b"\xff" * 32
    klib_vt(sc, bytes(bad_pt), off.ctypes.data, M, ctypes.cast(out2, ctypes.c_void_p), status.ctypes.data)
    assert status[victim] == 1 and (np.delete(status, victim) == 0).all() and out2.raw[32 * victim:32 * victim + 32] == bytes(32)
    klib_vt(bytes(bad_sc), pt, off.ctypes.data, M, ctypes.cast(out2, ctypes.c_void_p), status.ctypes.data)
    assert status[victim] == 3 and (np.delete(status, victim) == 0).all()
    assert klib_ct(sc, bytes(bad_pt), off.ctypes.data, M, ctypes.cast(out2, ctypes.c_void_p)) == 1
    assert klib_ct(bytes(bad_sc), pt, off.ctypes.data, M, ctypes.cast(out2, ctypes.c_void_p)) == 3


def test_headline_msm_pipeline_kernel_by_kernel_on_the_host(klib):
    """The variable-time MSM of the north-star path run kernel by kernel on the CPU (k_ingest2 histogram / scatter phases,
    k_scan, k_plan, k_items, k_len_hist / k_len_scatter, k_accumulate, k_merge, k_chunk_reduce, k_tree_sum, k_finish -- the
    cooperating ones with one OS thread per CUDA thread, tests/host_emul/cuda_shim.h) in the launch order of api.cu: the
    encoding equals the C port's (oracle) for several window widths, work-item lengths that cut buckets into many
    chunks, odd sizes, skewed scalars (one bucket takes everything), a cancelling instance (identity), and the first
    undecodable point / non-canonical scalar is reported with its index."""
    import struct
    import numpy as np
    from oracle import cref, scalar as S
    from tests import util_data as U
    klib.emul_msm_vartime.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_void_p]
    base = U.base_points(64)

    def run(sc, pts, c, S_forced=0, balance=1, host_window_scan=1):
        # host_window_scan: the W-block scan is the slow part of the emulation (1024 OS threads per block); it runs as a
        # kernel in ONE of the cases below and as a host prefix sum in the others
        out = ctypes.create_string_buffer(64)
        klib.emul_msm_vartime(bytes(sc), bytes(pts), len(sc) // 32, c, S_forced, balance, host_window_scan,
                              ctypes.cast(out, ctypes.c_void_p))
        status, ident, first_bad = struct.unpack_from("<iiq", out.raw, 32)
        return out.raw[:32], status, ident, first_bad

    n = 1500
    sc = U.random_scalars(n, seed=5).tobytes()
    pts = b"".join(base[(3 * i) % 64] for i in range(n))
    want = cref.msm_vartime(np.frombuffer(sc, np.uint8).reshape(-1, 32), np.frombuffer(pts, np.uint8).reshape(-1, 32))
    cases = [(8, 0, 1, 1), (8, 12, 0, 1)]
    if os.environ.get("ZKP_SLOW_TESTS") == "1":     # the W-block scan as a kernel: also covered by tests/test_api_emul.py
        cases.append((10, 0, 1, 0))
    for c, s_forced, balance, hws in cases:
        enc, status, ident, first_bad = run(sc, pts, c, s_forced, balance, hws)
        assert (enc, status, ident, first_bad) == (want, 0, 0, -1), (c, s_forced, balance)
    # odd size, one term, empty
    for m in (1, 333):
        w2 = cref.msm_vartime(np.frombuffer(sc[:32 * m], np.uint8).reshape(-1, 32), np.frombuffer(pts[:32 * m], np.uint8).reshape(-1, 32))
        assert run(sc[:32 * m], pts[:32 * m], 6, 64)[0] == w2, m
    assert run(b"", b"", 8) == (bytes(32), 0, 1, -1)
    # all scalars equal: one bucket per window takes every term (its work items are cut and merged)
    same = (12345678901234567890123456789 % S.L).to_bytes(32, "little") * 600
    w3 = cref.msm_vartime(np.frombuffer(same, np.uint8).reshape(-1, 32), np.frombuffer(pts[:32 * 600], np.uint8).reshape(-1, 32))
    assert run(same, pts[:32 * 600], 6, 8)[0] == w3
    # a cancelling instance: s * P + (l - s) * P for every pair -> the identity
    half = U.random_scalars(200, seed=9)
    neg = b"".join(((S.L - int.from_bytes(bytes(x), "little")) % S.L).to_bytes(32, "little") for x in half)
    enc, status, ident, _ = run(half.tobytes() + neg, pts[:32 * 200] * 2, 6, 64)
    assert (enc, status, ident) == (bytes(32), 0, 1)
    # failures carry the index of the first offender
    bad_pts = bytearray(pts); bad_pts[32 * 777:32 * 778] = b"\xff" * 32; bad_pts[32 * 1400:32 * 1401] = b"\xff" * 32
    assert run(sc, bad_pts, 6, 64)[1:] == (1, 0, 777)
    bad_sc = bytearray(sc); bad_sc[32 * 41:32 * 42] = b"\xff" * 32
    assert run(bad_sc, pts, 6, 64)[1:] == (3, 0, 41)


def test_batch_verification_front_end_kernels_on_the_host(klib):
    """k_bv_prepare2 (host-compiled transcript script), k_bv_prepare (byte-wise STROBE) and k_bv_static_sum run on the CPU
    with the plan of zkp_batch_verify_proofs (bv_plan.hpp): for real CMZ'13 proofs and for 130 DLEQ proofs (two launches,
    the second block of the first only partly filled) the MSM inputs -- per-proof challenges, weights, folded
    coefficients, block-reduced static coefficients, point rows -- are byte-equal to the oracle's BatchVerifier
    (/root/reference/src/toolbox/batch_verifier.rs:137-228 restated); a tampered commitment changes them, an identity
    commitment and a non-canonical response are reported with the proof's index."""
    import numpy as np
    from oracle import merlin as OM, msm as M, scalar as S, toolbox as OT

    class OneShot:
        def __init__(self, b): self.b = b
        def bytes(self, n): return self.b

    klib.emul_bv_front_end.argtypes = ([ctypes.c_int] * 4 + [ctypes.c_char_p] + [ctypes.c_void_p] * 5 + [ctypes.c_size_t]
                                       + [ctypes.c_char_p] * 5 + [ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                                                  ctypes.POINTER(ctypes.c_longlong)])

    def front_end(ost, tlabel, encs, proofs, seed, compiled, chunk=128):
        names = ost.instance + ost.common
        m, ni, nc, k, N = len(ost.secrets), len(ost.instance), len(ost.common), len(ost.constraints), len(proofs)
        lhs = np.array([names.index(l) for l, _ in ost.constraints], dtype=np.int32)
        off = np.cumsum([0] + [len(r) for _, r in ost.constraints]).astype(np.int32)
        ts = np.array([ost.secrets.index(s) for _, r in ost.constraints for s, _ in r], dtype=np.int32)
        tp = np.array([names.index(q) for _, r in ost.constraints for _, q in r], dtype=np.int32)
        t = OM.Transcript(tlabel)
        OT.domain_sep(t, ost.label)
        for s_ in ost.secrets:
            OT.append_scalar_var(t, s_.encode())
        st = t.strobe
        state = bytes(st.state)
        prefix = np.array([int.from_bytes(state[4 * i:4 * i + 4], "little") for i in range(50)]
                          + [st.pos, st.pos_begin, st.cur_flags], dtype=np.uint32)
        inst = b"".join(encs[n][j] for n in ost.instance for j in range(N))
        comm = b"".join(encs[n] for n in ost.common)
        com = b"".join(c for p in proofs for c in p.commitments)
        resp = b"".join(r if isinstance(r, bytes) else S.to_bytes(r) for p in proofs for r in p.responses)
        n = nc + (ni + k) * N
        co, po = ctypes.create_string_buffer(n * 32), ctypes.create_string_buffer(n * 32)
        bad = ctypes.c_longlong(-2)
        rc = klib.emul_bv_front_end(m, ni, nc, k, b"".join(x.encode() + b"\0" for x in names), lhs.ctypes.data, off.ctypes.data,
                                    ts.ctypes.data, tp.ctypes.data, prefix.ctypes.data, N, inst or b"\0", comm or b"\0", com or b"\0",
                                    resp or b"\0", seed, compiled, chunk, ctypes.cast(co, ctypes.c_void_p),
                                    ctypes.cast(po, ctypes.c_void_p), ctypes.byref(bad))
        return rc, bad.value, [co.raw[32 * i:32 * i + 32] for i in range(n)], [po.raw[32 * i:32 * i + 32] for i in range(n)]

    def oracle_inputs(ost, tlabel, encs, proofs, seed):
        N = len(proofs)
        bv = ost.build_batch_verifier(N, [OM.Transcript(tlabel) for _ in range(N)], encs)
        oscal, opts = bv.batch_coeffs(proofs, OT.PerProofRng(seed))
        return [S.to_bytes(s) for s in oscal], opts

    seed = bytes(range(32))
    rng = OT.SeededRng(b"emul-bv")
    # ---- CMZ'13, 3 real proofs ----
    ost = OT.CMZ10
    common = {n: R.from_uniform_bytes(rng.bytes(64)) for n in ost.common}
    proofs, encs = [], {n: [] for n in ost.instance}
    for _ in range(3):
        sec = {n: int.from_bytes(rng.bytes(64), "little") % S.L for n in ost.secrets}
        Pp, Q = R.from_uniform_bytes(rng.bytes(64)), R.from_uniform_bytes(rng.bytes(64))
        pts = dict(common)
        pts["P"], pts["Q"] = Pp, Q
        for i in range(1, 11):
            pts["C_%d" % i] = M.naive_msm([sec["m_%d" % i], sec["z_%d" % i]], [Pp, pts["A"]])
        pts["V"] = M.naive_msm([sec["m_%d" % i] for i in range(1, 11)] + [sec["minus_z_Q"]],
                               [pts["X_%d" % i] for i in range(1, 11)] + [Q])
        proof, enc = ost.prove_batchable(OM.Transcript(b"CMZ"), sec, pts, OneShot(rng.bytes(32)))
        proofs.append(proof)
        for n in ost.instance:
            encs[n].append(enc[n])
        for n in ost.common:
            encs[n] = enc[n]
    want_c, want_p = oracle_inputs(ost, b"CMZ", encs, proofs, seed)
    for compiled in (1, 0):
        rc, bad, co, po = front_end(ost, b"CMZ", encs, proofs, seed, compiled)
        assert (rc, bad) == (0, -1) and po == want_p and co == want_c, compiled
    tampered = [OT.BatchableProof(list(p.commitments), list(p.responses)) for p in proofs]
    tampered[1].commitments[4] = encs["P"][0]
    rc, bad, co2, _ = front_end(ost, b"CMZ", encs, tampered, seed, 1)
    assert rc == 0 and co2 != want_c
    tampered[1].commitments[4] = bytes(32)                      # identity encoding (toolbox/mod.rs:215)
    assert front_end(ost, b"CMZ", encs, tampered, seed, 1)[:2] == (1, 1)
    noncanon = [OT.BatchableProof(list(p.commitments), list(p.responses)) for p in proofs]
    noncanon[2].responses[7] = b"\xff" * 32
    assert front_end(ost, b"CMZ", encs, noncanon, seed, 1)[:2] == (3, 2)
    # ---- DLEQ, 130 proofs: two launches of 128 and 2 proofs ----
    ost = OT.DLEQ
    G = R.BASEPOINT
    H = R.hash_from_bytes_sha512(R.compress(G))
    N = 130
    proofs, encs = [], {n: [] for n in ost.instance}
    for j in range(N):
        x = 89327492234 + j
        pts = {"A": R.pt_mul(x, G), "B": R.pt_mul(x, H), "H": H, "G": G}
        proof, enc = ost.prove_batchable(OM.Transcript(b"DLEQBatchTest"), {"x": x}, pts, OneShot(rng.bytes(32)))
        proofs.append(proof)
        for n in ost.instance:
            encs[n].append(enc[n])
        encs["G"] = enc["G"]
    want_c, want_p = oracle_inputs(ost, b"DLEQBatchTest", encs, proofs, seed)
    for compiled in (1, 0):
        rc, bad, co, po = front_end(ost, b"DLEQBatchTest", encs, proofs, seed, compiled)
        assert (rc, bad) == (0, -1) and po == want_p and co == want_c, compiled
    # one slab of 130 proofs on a grid of ONE block (the resident grid of the slab pipeline): the block takes both groups
    rc, bad, co, po = front_end(ost, b"DLEQBatchTest", encs, proofs, seed, 1, chunk=256)
    assert (rc, bad) == (0, -1) and po == want_p and co == want_c
