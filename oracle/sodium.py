"""libsodium 1.0.20 ristretto255 (found inside the pyzmq wheel) as an INDEPENDENT cross-check of the oracle.
TEST INFRASTRUCTURE; only present in this container's image, so it is used when generating the golden
fixtures (tests/golden/make_golden.py) and in CPU tests that skip when it is absent."""
import ctypes
import glob
import os
import site

_L = 2**252 + 27742317777372353535851937790883648493


def find():
    roots = list(site.getsitepackages()) + [os.path.dirname(os.path.dirname(ctypes.__file__))]
    for r in roots:
        for p in glob.glob(os.path.join(r, "pyzmq.libs", "libsodium*.so*")):
            return p
    for p in glob.glob("/opt/prime-rl/.venv/lib/python3*/site-packages/pyzmq.libs/libsodium*.so*"):
        return p
    return None


def load():
    p = find()
    if p is None:
        return None
    lib = ctypes.CDLL(p)
    if lib.sodium_init() < 0 or not hasattr(lib, "crypto_scalarmult_ristretto255"):
        return None
    return lib


def is_valid_point(lib, enc):
    return lib.crypto_core_ristretto255_is_valid_point(bytes(enc)) == 1


def from_hash(lib, h64):
    out = ctypes.create_string_buffer(32)
    lib.crypto_core_ristretto255_from_hash(out, bytes(h64))
    return out.raw


def scalarmult(lib, k, enc):
    """k * P as an encoding; libsodium returns -1 for the identity / zero scalar: map that to 32 zero bytes."""
    out = ctypes.create_string_buffer(32)
    rc = lib.crypto_scalarmult_ristretto255(out, (k % _L).to_bytes(32, "little"), bytes(enc))
    return bytes(32) if rc != 0 else out.raw


def scalarmult_base(lib, k):
    out = ctypes.create_string_buffer(32)
    rc = lib.crypto_scalarmult_ristretto255_base(out, (k % _L).to_bytes(32, "little"))
    return bytes(32) if rc != 0 else out.raw


def add(lib, a, b):
    out = ctypes.create_string_buffer(32)
    assert lib.crypto_core_ristretto255_add(out, bytes(a), bytes(b)) == 0
    return out.raw


def msm(lib, scalars, encs):
    acc = bytes(32)
    for k, e in zip(scalars, encs):
        acc = add(lib, acc, scalarmult(lib, k, e))
    return acc
