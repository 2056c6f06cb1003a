"""ristretto255 in Python big-int arithmetic.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates RFC 9496 section 4 (= curve25519-dalek 2.x `ristretto.rs` / `field.rs` [ext]) for the calls the
reference makes at:
  decompress  /root/reference/src/toolbox/verifier.rs:90, :164, batch_verifier.rs:226
  compress    /root/reference/src/toolbox/mod.rs:180, :204
  is_identity /root/reference/src/toolbox/mod.rs:191, :215 (on encodings), verifier.rs:168,
              batch_verifier.rs:230 (on points; coset-aware)
Points are tuples (X, Y, Z, T) of ints mod p in extended twisted-Edwards coordinates (a = -1).
"""
import hashlib

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
D = (-121665 * pow(121666, P - 2, P)) % P
D2 = (2 * D) % P
SQRT_M1 = pow(2, (P - 1) // 4, P)
# RFC 9496 section 4.1 constants
INVSQRT_A_MINUS_D = 54469307008909316920995813868745141605393597292927456921205312896311721017578
SQRT_AD_MINUS_ONE = 25063068953384623474111414158702152701244531502492656460079210482610430750235
ONE_MINUS_D_SQ = 1159843021668779879193775521855586647937357759715417654439879720876111806838
D_MINUS_ONE_SQ = 40440834346308536858101042469323190826248399146238708352240133220865137265952

IDENTITY = (0, 1, 1, 0)
# Ed25519 basepoint (RFC 8032); ristretto255 generator is its coset.
_BY = (4 * pow(5, P - 2, P)) % P
_BX = 15112221349535400772501151409588531511454012693041857206046113283949847762202
BASEPOINT = (_BX, _BY, 1, (_BX * _BY) % P)


def is_negative(x):
    """RFC 9496 4.1: low bit of the canonical little-endian encoding."""
    return (x % P) & 1


def fe_abs(x):
    x %= P
    return (P - x) % P if x & 1 else x


def sqrt_ratio_i(u, v):
    """RFC 9496 4.2 SQRT_RATIO_M1 (= dalek FieldElement::sqrt_ratio_i [ext])."""
    u %= P
    v %= P
    v3 = (v * v % P) * v % P
    v7 = (v3 * v3 % P) * v % P
    r = (u * v3 % P) * pow(u * v7 % P, (P - 5) // 8, P) % P
    check = v * r % P * r % P
    correct = check == u
    flipped = check == (P - u) % P
    flipped_i = check == ((P - u) * SQRT_M1) % P
    if flipped or flipped_i:
        r = r * SQRT_M1 % P
    r = fe_abs(r)
    return (correct or flipped), r


def decompress(b):
    """RFC 9496 4.3.1 Decode.  Returns extended point or None (= CompressedRistretto::decompress [ext])."""
    assert len(b) == 32
    s = int.from_bytes(b, "little")
    if s >= P or (s & 1):
        return None
    ss = s * s % P
    u1 = (1 - ss) % P
    u2 = (1 + ss) % P
    u2_sqr = u2 * u2 % P
    v = (-(D * u1 % P * u1) - u2_sqr) % P
    ok, invsqrt = sqrt_ratio_i(1, v * u2_sqr % P)
    den_x = invsqrt * u2 % P
    den_y = invsqrt * den_x % P * v % P
    x = fe_abs(2 * s * den_x % P)
    y = u1 * den_y % P
    t = x * y % P
    if (not ok) or is_negative(t) or y == 0:
        return None
    return (x, y, 1, t)


def compress(pt):
    """RFC 9496 4.3.2 Encode (= RistrettoPoint::compress [ext])."""
    X, Y, Z, T = (c % P for c in pt)
    u1 = (Z + Y) * (Z - Y) % P
    u2 = X * Y % P
    _, invsqrt = sqrt_ratio_i(1, u1 * u2 % P * u2 % P)
    den1 = invsqrt * u1 % P
    den2 = invsqrt * u2 % P
    z_inv = den1 * den2 % P * T % P
    ix0 = X * SQRT_M1 % P
    iy0 = Y * SQRT_M1 % P
    enchanted = den1 * INVSQRT_A_MINUS_D % P
    rotate = is_negative(T * z_inv % P)
    if rotate:
        x, y, den_inv = iy0, ix0, enchanted
    else:
        x, y, den_inv = X, Y, den2
    if is_negative(x * z_inv % P):
        y = (P - y) % P
    s = fe_abs(den_inv * ((Z - y) % P) % P)
    return s.to_bytes(32, "little")


# ---- Edwards group law (dalek "models": Extended + ProjectiveNiels -> Completed -> Extended) -------------
def pt_add(p, q):
    X1, Y1, Z1, T1 = p
    X2, Y2, Z2, T2 = q
    PP = (Y1 + X1) * (Y2 + X2) % P
    MM = (Y1 - X1) * (Y2 - X2) % P
    TT2d = T1 * T2 % P * D2 % P
    ZZ2 = 2 * Z1 * Z2 % P
    cx, cy, cz, ct = (PP - MM) % P, (PP + MM) % P, (ZZ2 + TT2d) % P, (ZZ2 - TT2d) % P
    return (cx * ct % P, cy * cz % P, cz * ct % P, cx * cy % P)


def pt_neg(p):
    X, Y, Z, T = p
    return ((P - X) % P, Y, Z, (P - T) % P)


def pt_sub(p, q):
    return pt_add(p, pt_neg(q))


def pt_double(p):
    X, Y, Z, _ = p
    XX = X * X % P
    YY = Y * Y % P
    ZZ2 = 2 * Z * Z % P
    XpY2 = (X + Y) * (X + Y) % P
    cx, cy, cz = (XpY2 - YY - XX) % P, (YY + XX) % P, (YY - XX) % P
    ct = (ZZ2 - cz) % P
    return (cx * ct % P, cy * cz % P, cz * ct % P, cx * cy % P)


def pt_mul(k, p):
    """Plain double-and-add; k any non-negative int (reduced mod l by the caller if desired)."""
    acc = IDENTITY
    for bit in bin(k)[2:] if k else "":
        acc = pt_double(acc)
        if bit == "1":
            acc = pt_add(acc, p)
    return acc


def pt_eq(p, q):
    """Ristretto (coset-aware) equality: X1*Y2 == Y1*X2 or X1*X2 == Y1*Y2  (dalek ct_eq [ext])."""
    X1, Y1, _, _ = p
    X2, Y2, _, _ = q
    return (X1 * Y2 - Y1 * X2) % P == 0 or (X1 * X2 - Y1 * Y2) % P == 0


def is_identity(p):
    """Coset-aware identity (verifier.rs:168, batch_verifier.rs:230): equal to (0,1) as a ristretto point."""
    return pt_eq(p, IDENTITY)


def is_identity_encoding(b):
    """CompressedRistretto::is_identity (toolbox/mod.rs:191,215): 32 zero bytes."""
    return bytes(b) == bytes(32)


def on_curve(p):
    X, Y, Z, T = p
    lhs = (-X * X + Y * Y) % P * Z % P * Z % P
    rhs = (pow(Z, 4, P) + D * X % P * X % P * Y % P * Y) % P
    return lhs == rhs and (X * Y - Z * T) % P == 0


# ---- hash-to-group (RFC 9496 4.3.4; dalek from_uniform_bytes / hash_from_bytes [ext]) ---------------------
def _elligator_map(t):
    r = SQRT_M1 * t % P * t % P
    u = (r + 1) * ONE_MINUS_D_SQ % P
    v = (-1 - r * D) % P * ((r + D) % P) % P
    was_square, s = sqrt_ratio_i(u, v)
    s_prime = (P - fe_abs(s * t % P)) % P
    c = P - 1
    if not was_square:
        s = s_prime
        c = r
    N = (c * ((r - 1) % P) % P * D_MINUS_ONE_SQ - v) % P
    w0 = 2 * s * v % P
    w1 = N * SQRT_AD_MINUS_ONE % P
    w2 = (1 - s * s) % P
    w3 = (1 + s * s) % P
    return (w0 * w3 % P, w2 * w1 % P, w1 * w3 % P, w0 * w2 % P)


def from_uniform_bytes(b):
    assert len(b) == 64
    r0 = int.from_bytes(b[:32], "little") & ((1 << 255) - 1)
    r1 = int.from_bytes(b[32:], "little") & ((1 << 255) - 1)
    return pt_add(_elligator_map(r0 % P), _elligator_map(r1 % P))


def hash_from_bytes_sha512(msg):
    """RistrettoPoint::hash_from_bytes::<Sha512> as used in tests/zkp.rs:34, benches/dleq.rs:52."""
    return from_uniform_bytes(hashlib.sha512(msg).digest())


BASEPOINT_COMPRESSED = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")
