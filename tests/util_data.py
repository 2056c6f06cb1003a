"""Seeded test-data helpers shared by the CPU and GPU tests (uses the oracle: test infrastructure)."""
import hashlib
import json
import os

import numpy as np

from oracle import ristretto as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


_POINT_CACHE = {}


def base_points(k, seed=b"zkp-b200-points"):
    """k distinct valid ristretto255 encodings (hash-to-group of a counter), cached; returns list of bytes."""
    key = (k, seed)
    if key not in _POINT_CACHE:
        have = _POINT_CACHE.get(("max", seed), [])
        while len(have) < k:
            h = hashlib.sha512(seed + len(have).to_bytes(8, "little")).digest()
            have.append(R.compress(R.from_uniform_bytes(h)))
        _POINT_CACHE[("max", seed)] = have
        _POINT_CACHE[key] = have[:k]
    return _POINT_CACHE[key]


def random_scalars(n, seed, bits=None):
    """n canonical scalars as an (n,32) uint8 array; `bits` limits the magnitude (e.g. 128)."""
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(n, 64), dtype=np.uint8)
    out = np.zeros((n, 32), dtype=np.uint8)
    for i in range(n):
        v = int.from_bytes(raw[i].tobytes(), "little")
        v = v % R.L if bits is None else v % (1 << bits)
        out[i] = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8)
    return out


def scalars_to_ints(arr):
    return [int.from_bytes(arr[i].tobytes(), "little") for i in range(arr.shape[0])]


def tiled_expected(scalars_arr, base_encs):
    """Expected encoding of sum_i s_i * P_{i mod K} via the K-term reduction (exact for any n)."""
    from oracle import msm as M
    K = len(base_encs)
    sums = [0] * K
    n = scalars_arr.shape[0]
    # vectorised accumulation of scalar columns per residue class using Python ints on 64-bit words
    words = scalars_arr.reshape(n, 4, 8).copy().view(np.uint64).reshape(n, 4)
    for k in range(K):
        col = words[k::K]
        if col.shape[0] == 0:
            continue
        tot = 0
        for w in range(4):
            # split to avoid uint64 overflow
            lo = int((col[:, w] & np.uint64(0xffffffff)).sum(dtype=np.uint64)) if col.shape[0] < (1 << 31) else None
            hi = int((col[:, w] >> np.uint64(32)).sum(dtype=np.uint64))
            tot += (lo + (hi << 32)) << (64 * w)
        sums[k] = tot % R.L
    pts = [R.decompress(e) for e in base_encs]
    return R.compress(M.naive_msm(sums, pts))


def can_spawn_threads(n=1100):
    """The host emulation of cooperating kernels runs one OS thread per CUDA thread (1024 per block): check that this
    environment lets a process have that many before a test relies on it."""
    import threading
    ev, started, ths = threading.Event(), [], []
    def body():
        started.append(1)
        ev.wait()
    try:
        for _ in range(n):
            t = threading.Thread(target=body)
            t.start()
            ths.append(t)
        ok = True
    except RuntimeError:
        ok = False
    ev.set()
    for t in ths:
        t.join()
    return ok
