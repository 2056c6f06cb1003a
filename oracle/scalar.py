"""Scalars mod l and dalek's digit recodings.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates curve25519-dalek 2.x `scalar.rs` [ext] for the operations the reference's callers use
(SURVEY.md section 8a row a13): Neg (verifier.rs:95,142), From<u128> (verifier.rs:153, batch_verifier.rs:179),
mul/add/sub (verifier.rs:155-158, batch_verifier.rs:183-201), from_bytes_mod_order_wide (toolbox/mod.rs:226),
s*c+b (prover.rs:108); and the recodings the MSM algorithms consume (to_radix_16, non_adjacent_form(5),
to_radix_2w).  Scalars are Python ints in [0, l); wire form is 32 bytes little-endian canonical.
"""
from .ristretto import L


def from_bytes_mod_order_wide(b64):
    assert len(b64) == 64
    return int.from_bytes(b64, "little") % L


def from_bytes_mod_order(b32):
    assert len(b32) == 32
    return int.from_bytes(b32, "little") % L


def from_canonical_bytes(b32):
    v = int.from_bytes(b32, "little")
    return v if v < L else None


def to_bytes(s):
    return (s % L).to_bytes(32, "little")


def neg(s):
    return (L - s) % L


def to_radix_16(k):
    """Scalar::to_radix_16: 64 signed digits in [-8, 8), k = sum d_i 16^i (requires k < 2^255)."""
    assert 0 <= k < 2**255
    d = [(k >> (4 * i)) & 15 for i in range(64)]
    for i in range(63):
        carry = (d[i] + 8) >> 4
        d[i] -= carry << 4
        d[i + 1] += carry
    return d


def non_adjacent_form(k, w=5):
    """Scalar::non_adjacent_form(w): 256 digits, non-zero ones odd in (-2^(w-1), 2^(w-1))."""
    assert 2 <= w <= 8
    naf = [0] * 256
    width = 1 << w
    window_mask = width - 1
    pos, carry = 0, 0
    while pos < 256:
        window = carry + ((k >> pos) & window_mask)
        if window & 1 == 0:
            pos += 1
            continue
        if window < width // 2:
            carry = 0
            naf[pos] = window
        else:
            carry = 1
            naf[pos] = window - width
        pos += w
    return naf


def to_radix_2w_size_hint(w):
    assert 6 <= w <= 8
    return (256 + w - 1) // w + (1 if w == 8 else 0)


def to_radix_2w(k, w):
    """Scalar::to_radix_2w(w), w in {6,7,8}: signed digits in [-2^(w-1), 2^(w-1)), fixed 43-entry array."""
    assert 6 <= w <= 8
    radix = 1 << w
    digits_count = (256 + w - 1) // w
    digits = [0] * 43
    carry = 0
    for i in range(digits_count):
        coef = carry + ((k >> (w * i)) & (radix - 1))
        carry = (coef + radix // 2) >> w
        digits[i] = coef - (carry << w)
    if w == 8:
        digits[digits_count] += carry
    else:
        digits[digits_count - 1] += carry << w
    return digits
