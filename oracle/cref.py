"""ctypes face of oracle/_ref/libref_u64.so (the C port of the reference's CPU path).  TEST INFRASTRUCTURE."""
import ctypes

import numpy as np

from .c import build as _build

_lib = None


def load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_build.build())
        vp, sz = ctypes.c_void_p, ctypes.c_size_t
        lib.ref_msm_vartime.argtypes = [vp, vp, sz, vp, ctypes.POINTER(ctypes.c_int64)]
        lib.ref_msm_vartime_mt.argtypes = [vp, vp, sz, ctypes.c_int, vp, ctypes.POINTER(ctypes.c_int64)]
        lib.ref_msm_vartime_simd.argtypes = [vp, vp, sz, vp, ctypes.POINTER(ctypes.c_int64)]
        lib.ref_msm_vartime_mt_simd.argtypes = [vp, vp, sz, ctypes.c_int, vp, ctypes.POINTER(ctypes.c_int64)]
        lib.ref_simd_available.restype = ctypes.c_int
        lib.ref_simd_selftest_fe.argtypes = [vp, vp, vp, vp]
        lib.ref_msm_ct_batched.argtypes = [vp, vp, vp, sz, vp]
        lib.ref_msm_vartime_batched.argtypes = [vp, vp, vp, sz, vp, vp]
        lib.ref_decompress.argtypes = [vp, sz, vp, vp]
        lib.ref_compress.argtypes = [vp, sz, vp]
        _lib = lib
    return _lib


def _a(x, w=32):
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.uint8).reshape(-1, w)
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8).reshape(-1, w)
    return np.frombuffer(b"".join(bytes(b) for b in x), dtype=np.uint8).reshape(-1, w)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def simd_available():
    """True when this CPU can run the 4-way vector restatement of the reference's simd_backend (oracle/c/ref_ifma.h)."""
    return bool(load().ref_simd_available())


def msm_vartime(scalars, points, threads=1, simd=False):
    """optional_multiscalar_mul with dalek's dispatch; returns encoding bytes or None.
    simd=True runs the vector backend (raises when the CPU lacks AVX-512 IFMA: callers check simd_available())."""
    lib = load()
    s, p = _a(scalars), _a(points)
    out = np.zeros(32, dtype=np.uint8)
    bad = ctypes.c_int64(-1)
    if simd:
        if threads == 1:
            rc = lib.ref_msm_vartime_simd(_p(s), _p(p), s.shape[0], _p(out), ctypes.byref(bad))
        else:
            rc = lib.ref_msm_vartime_mt_simd(_p(s), _p(p), s.shape[0], int(threads), _p(out), ctypes.byref(bad))
        if rc < 0:
            raise RuntimeError("vector backend unavailable on this CPU (needs avx2, avx512f, avx512vl, avx512ifma)")
    elif threads == 1:
        rc = lib.ref_msm_vartime(_p(s), _p(p), s.shape[0], _p(out), ctypes.byref(bad))
    else:
        rc = lib.ref_msm_vartime_mt(_p(s), _p(p), s.shape[0], int(threads), _p(out), ctypes.byref(bad))
    return None if rc else out.tobytes()


def simd_field_selftest(a, b):
    """(a*b, a*a) of four field-element pairs through the vector field; a, b: uint64 [4][5] radix-2^51 limbs."""
    lib = load()
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(4, 5)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(4, 5)
    m, q = np.zeros((4, 5), np.uint64), np.zeros((4, 5), np.uint64)
    if lib.ref_simd_selftest_fe(_p(a), _p(b), _p(m), _p(q)) != 0:
        raise RuntimeError("vector backend unavailable on this CPU")
    return m, q


def msm_ct_batched(scalars, points, offsets):
    lib = load()
    s, p = _a(scalars), _a(points)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    out = np.zeros((off.shape[0] - 1, 32), dtype=np.uint8)
    rc = lib.ref_msm_ct_batched(_p(s), _p(p), _p(off), off.shape[0] - 1, _p(out))
    return None if rc else out


def msm_vartime_batched(scalars, points, offsets):
    lib = load()
    s, p = _a(scalars), _a(points)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    M = off.shape[0] - 1
    out = np.zeros((M, 32), dtype=np.uint8)
    valid = np.zeros(M, dtype=np.uint8)
    lib.ref_msm_vartime_batched(_p(s), _p(p), _p(off), M, _p(out), _p(valid))
    return out, valid


def decompress(encs):
    lib = load()
    e = _a(encs)
    limbs = np.zeros((e.shape[0], 4, 5), dtype=np.uint64)
    valid = np.zeros(e.shape[0], dtype=np.uint8)
    lib.ref_decompress(_p(e), e.shape[0], _p(limbs), _p(valid))
    return limbs, valid


def compress(limbs):
    lib = load()
    l = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4, 5)
    out = np.zeros((l.shape[0], 32), dtype=np.uint8)
    lib.ref_compress(_p(l), l.shape[0], _p(out))
    return out
