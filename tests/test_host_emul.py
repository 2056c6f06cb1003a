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
    deps = [src] + [os.path.join(HERE, "..", "zkp_b200", "csrc", f) for f in ("fe.cuh", "ge.cuh", "sc.cuh", "hash.cuh", "scl.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-DZKP_HOST_EMUL", "-x", "c++", src, "-o", out])
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
        assert _c1(lib.emul_fe_canon, a) == a % P
        for b in rnd.sample(vals, 8) + edge:
            assert _c2(lib.emul_fe_mul, a, b) % P == a * b % P
            assert _c2(lib.emul_fe_add, a, b) % P == (a + b) % P
            assert _c2(lib.emul_fe_sub, a, b) % P == (a - b) % P
    for a in vals[:40]:
        if a % P:
            assert _c1(lib.emul_fe_invert, a) * a % P == 1
        assert _c1(lib.emul_fe_pow22523, a) % P == pow(a, (P - 5) // 8, P)
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
        affs.append(a.raw)
    for _ in range(300):
        e = rnd.randbytes(32)
        if rnd.random() < 0.5:
            e = e[:31] + bytes([e[31] & 0x7F])
        a = ctypes.create_string_buffer(96)
        assert lib.emul_decode(a, e) == (R.decompress(e) is not None)
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
    for c in (4, 7, 8, 13, 16, 20, 23):
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
    for w in [0, 2**512 - 1, (L << 256) - 1, L << 200, L * L - 1] + [rnd.getrandbits(512) for _ in range(300)]:
        lib.emul_scl_wide(o, w.to_bytes(64, "little"))
        assert int.from_bytes(o.raw, "little") == w % L
