"""Multi-scalar multiplication exactly as curve25519-dalek 2.x performs it.  TEST INFRASTRUCTURE.

Restates `backend/serial/scalar_mul/{straus,pippenger}.rs` and the dispatch in `edwards.rs` [ext] behind the
reference call sites
  RistrettoPoint::multiscalar_mul           /root/reference/src/toolbox/prover.rs:94-97       (constant time)
  RistrettoPoint::vartime_multiscalar_mul   /root/reference/src/toolbox/verifier.rs:97-106
  RistrettoPoint::optional_multiscalar_mul  /root/reference/src/toolbox/verifier.rs:162-166,
                                            /root/reference/src/toolbox/batch_verifier.rs:219-228
The group element returned is independent of the algorithm; the algorithms are kept distinct so that the
oracle also documents (and the CPU baseline in oracle/c times) the work the reference actually does.
"""
from . import ristretto as R
from . import scalar as S


def naive_msm(scalars, points):
    acc = R.IDENTITY
    for k, p in zip(scalars, points):
        acc = R.pt_add(acc, R.pt_mul(k % R.L, p))
    return acc


def straus_ct(scalars, points):
    """Straus::multiscalar_mul: radix-16 signed digits, LookupTable of 1P..8P, 4 doublings per digit."""
    tables = []
    for p in points:
        t = [p]
        for _ in range(7):
            t.append(R.pt_add(t[-1], p))
        tables.append(t)
    digits = [S.to_radix_16(k) for k in scalars]
    q = R.IDENTITY
    for j in reversed(range(64)):
        for _ in range(4):
            q = R.pt_double(q)
        for d, t in zip(digits, tables):
            dj = d[j]
            if dj > 0:
                q = R.pt_add(q, t[dj - 1])
            elif dj < 0:
                q = R.pt_sub(q, t[-dj - 1])
            # dj == 0: the constant-time code adds the identity
    return q


def straus_vartime(scalars, points):
    """Straus::optional_multiscalar_mul: NAF(5), NafLookupTable5 of odd multiples 1P,3P,..,15P."""
    nafs = [S.non_adjacent_form(k, 5) for k in scalars]
    tables = []
    for p in points:
        p2 = R.pt_double(p)
        t = [p]
        for _ in range(7):
            t.append(R.pt_add(t[-1], p2))
        tables.append(t)
    r = R.IDENTITY
    for i in reversed(range(256)):
        r = R.pt_double(r)
        for naf, t in zip(nafs, tables):
            d = naf[i]
            if d > 0:
                r = R.pt_add(r, t[d // 2])
            elif d < 0:
                r = R.pt_sub(r, t[(-d) // 2])
    return r


def pippenger_window(n):
    return 6 if n < 500 else (7 if n < 800 else 8)


def pippenger(scalars, points, w=None):
    """Pippenger::optional_multiscalar_mul: signed radix-2^w digits, 2^(w-1) buckets, running-sum reduce."""
    n = len(scalars)
    if w is None:
        w = pippenger_window(n)
    max_digit = 1 << w
    digits_count = S.to_radix_2w_size_hint(w)
    buckets_count = max_digit // 2
    digits = [S.to_radix_2w(k, w) for k in scalars]

    def column(idx):
        buckets = [R.IDENTITY] * buckets_count
        for d, p in zip(digits, points):
            di = d[idx]
            if di > 0:
                buckets[di - 1] = R.pt_add(buckets[di - 1], p)
            elif di < 0:
                buckets[-di - 1] = R.pt_sub(buckets[-di - 1], p)
        run = buckets[buckets_count - 1]
        acc = buckets[buckets_count - 1]
        for i in reversed(range(buckets_count - 1)):
            run = R.pt_add(run, buckets[i])
            acc = R.pt_add(acc, run)
        return acc

    total = column(digits_count - 1)
    for idx in reversed(range(digits_count - 1)):
        for _ in range(w):
            total = R.pt_double(total)
        total = R.pt_add(total, column(idx))
    return total


def multiscalar_mul(scalars, points):
    """RistrettoPoint::multiscalar_mul (constant time) -> Straus CT for every size."""
    return straus_ct(list(scalars), list(points))


def vartime_multiscalar_mul(scalars, points):
    r = optional_multiscalar_mul(scalars, points)
    assert r is not None
    return r


def optional_multiscalar_mul(scalars, opt_points):
    """EdwardsPoint::optional_multiscalar_mul: size < 190 -> Straus vartime, else Pippenger; None if any
    point is None."""
    scalars = list(scalars)
    pts = list(opt_points)
    assert len(scalars) == len(pts)
    if any(p is None for p in pts):
        return None
    if len(scalars) < 190:
        return straus_vartime(scalars, pts)
    return pippenger(scalars, pts)


def msm_bytes(scalar_bytes, point_bytes):
    """The C-ABI shape: 32-byte scalars and 32-byte encodings in, 32-byte encoding (or None) out.
    Uses the fast naive sum (the group element is algorithm independent)."""
    pts = [R.decompress(bytes(b)) for b in point_bytes]
    if any(p is None for p in pts):
        return None
    ks = [int.from_bytes(bytes(b), "little") for b in scalar_bytes]
    return R.compress(naive_msm(ks, pts))
