// Edwards25519 group law + ristretto255 encode/decode for sm_100a, in registers.
//
// Replaces curve25519-dalek 2.x `edwards.rs` / `backend/serial/curve_models` (Extended, ProjectiveNiels,
// AffineNiels, Completed) and `ristretto.rs` (CompressedRistretto::decompress, RistrettoPoint::compress,
// coset-aware identity) [ext] -- SURVEY.md section 2.2 rows E6/E7, section 8a rows a10-a12 -- as used by
//   /root/reference/src/toolbox/verifier.rs:90,164  /root/reference/src/toolbox/batch_verifier.rs:226  (decompress)
//   /root/reference/src/toolbox/mod.rs:180,204                                                (compress)
//   /root/reference/src/toolbox/verifier.rs:168  /root/reference/src/toolbox/batch_verifier.rs:230 (is_identity)
// Formulas are the complete a=-1 twisted-Edwards ones of Hisil-Wong-Carter-Dawson 2008 (the same ones dalek
// evaluates); encode/decode follow RFC 9496 section 4.3.  Outputs are canonical encodings, hence bit-exact with
// the reference whatever the internal representation.
#pragma once
#include "fe.cuh"

namespace zkp {

#define ZKP_FE_CONST(name, w0, w1, w2, w3, w4, w5, w6, w7) \
  ZKP_DEV fe name() {                                       \
    fe r;                                                   \
    r.v[0] = w0; r.v[1] = w1; r.v[2] = w2; r.v[3] = w3;     \
    r.v[4] = w4; r.v[5] = w5; r.v[6] = w6; r.v[7] = w7;     \
    return r;                                               \
  }
ZKP_FE_CONST(fe_D, 0x135978a3u, 0x75eb4dcau, 0x4141d8abu, 0x00700a4du, 0x7779e898u, 0x8cc74079u, 0x2b6ffe73u, 0x52036ceeu)
ZKP_FE_CONST(fe_D2, 0x26b2f159u, 0xebd69b94u, 0x8283b156u, 0x00e0149au, 0xeef3d130u, 0x198e80f2u, 0x56dffce7u, 0x2406d9dcu)
ZKP_FE_CONST(fe_SQRT_M1, 0x4a0ea0b0u, 0xc4ee1b27u, 0xad2fe478u, 0x2f431806u, 0x3dfbd7a7u, 0x2b4d0099u, 0x4fc1df0bu, 0x2b832480u)
ZKP_FE_CONST(fe_INVSQRT_A_MINUS_D, 0x805d40eau, 0x99c8fdaau, 0x5a4172beu, 0x9d2f1617u, 0xfe01d840u, 0x16c27b91u, 0xcfaffca2u, 0x786c8905u)

struct ge_ext { fe X, Y, Z, T; };              // extended (X:Y:Z:T), T = XY/Z
struct ge_aniels { fe yplusx, yminusx, xy2d; };  // affine Niels, Z = 1 (96 bytes in HBM)
struct ge_pniels { fe YplusX, YminusX, Z, T2d; };  // projective Niels

ZKP_DEV void ge_identity(ge_ext& r) {
  fe_zero(r.X); fe_one(r.Y); fe_one(r.Z); fe_zero(r.T);
}
ZKP_DEV void ge_aniels_identity(ge_aniels& r) {
  fe_one(r.yplusx); fe_one(r.yminusx); fe_zero(r.xy2d);
}

// r = p + q (q affine Niels): 7M
ZKP_DEV void ge_madd(ge_ext& r, const ge_ext& p, const ge_aniels& q) {
  fe a, b, c, d, e, f, g, h;
  fe_sub(a, p.Y, p.X);
  fe_add(b, p.Y, p.X);
  fe_mul(a, a, q.yminusx);
  fe_mul(b, b, q.yplusx);
  fe_mul(c, p.T, q.xy2d);
  fe_add(d, p.Z, p.Z);
  fe_sub(e, b, a);
  fe_sub(f, d, c);
  fe_add(g, d, c);
  fe_add(h, b, a);
  fe_mul(r.X, e, f);
  fe_mul(r.Y, g, h);
  fe_mul(r.Z, f, g);
  fe_mul(r.T, e, h);
}

// r = p + (neg ? -q : q) with the sign folded into operand selection (no field negation, 32 selects):
// -q swaps (y+x, y-x) and flips the sign of 2dxy, i.e. swaps F = D - C and G = D + C.
// VT selects the variable-time tails of fe.cuh (public data only: the verifier's bucket accumulation).
template <bool VT = false>
ZKP_DEV void ge_madd_signed(ge_ext& r, const ge_ext& p, const ge_aniels& q, uint32_t neg) {
  fe a, b, c, d, e, f0, g0, f, g, h, qa, qb;
  fe_select(qa, q.yminusx, q.yplusx, neg);
  fe_select(qb, q.yplusx, q.yminusx, neg);
  fe_sub_t<VT>(a, p.Y, p.X);
  fe_add_t<VT>(b, p.Y, p.X);
  fe_mulx<VT>(a, a, qa);
  fe_mulx<VT>(b, b, qb);
  fe_mulx<VT>(c, p.T, q.xy2d);
  fe_add_t<VT>(d, p.Z, p.Z);
  fe_sub_t<VT>(e, b, a);
  fe_sub_t<VT>(f0, d, c);
  fe_add_t<VT>(g0, d, c);
  fe_select(f, f0, g0, neg);
  fe_select(g, g0, f0, neg);
  fe_add_t<VT>(h, b, a);
  fe_mulx<VT>(r.X, e, f);
  fe_mulx<VT>(r.Y, g, h);
  fe_mulx<VT>(r.Z, f, g);
  fe_mulx<VT>(r.T, e, h);
}

// conditional negation of an affine Niels point (branch-free): swap the first two, negate the third
ZKP_DEV void ge_aniels_cneg(ge_aniels& q, uint32_t neg) {
  fe_cswap(q.yplusx, q.yminusx, neg);
  fe_cneg(q.xy2d, q.xy2d, neg);
}
ZKP_DEV void ge_pniels_cneg(ge_pniels& q, uint32_t neg) {
  fe_cswap(q.YplusX, q.YminusX, neg);
  fe_cneg(q.T2d, q.T2d, neg);
}

ZKP_DEV void ge_to_pniels(ge_pniels& r, const ge_ext& p) {
  fe_add(r.YplusX, p.Y, p.X);
  fe_sub(r.YminusX, p.Y, p.X);
  r.Z = p.Z;
  fe_mul(r.T2d, p.T, fe_D2());
}

// r = p + q (q projective Niels): 8M
ZKP_DEV void ge_add_pniels(ge_ext& r, const ge_ext& p, const ge_pniels& q) {
  fe a, b, c, d, e, f, g, h;
  fe_sub(a, p.Y, p.X);
  fe_add(b, p.Y, p.X);
  fe_mul(a, a, q.YminusX);
  fe_mul(b, b, q.YplusX);
  fe_mul(c, p.T, q.T2d);
  fe_mul(d, p.Z, q.Z);
  fe_add(d, d, d);
  fe_sub(e, b, a);
  fe_sub(f, d, c);
  fe_add(g, d, c);
  fe_add(h, b, a);
  fe_mul(r.X, e, f);
  fe_mul(r.Y, g, h);
  fe_mul(r.Z, f, g);
  fe_mul(r.T, e, h);
}

// r = p + q, both extended: 9M
ZKP_DEV void ge_add(ge_ext& r, const ge_ext& p, const ge_ext& q) {
  ge_pniels n;
  ge_to_pniels(n, q);
  ge_add_pniels(r, p, n);
}

// r = 2p: 4S + 4M
ZKP_DEV void ge_double(ge_ext& r, const ge_ext& p) {
  fe xx, yy, zz2, xy2, e, g, f, h;
  fe_sq(xx, p.X);
  fe_sq(yy, p.Y);
  fe_sq(zz2, p.Z);
  fe_add(zz2, zz2, zz2);
  fe_add(xy2, p.X, p.Y);
  fe_sq(xy2, xy2);
  fe_add(h, yy, xx);        // Y' = YY + XX
  fe_sub(g, yy, xx);        // Z' = YY - XX
  fe_sub(e, xy2, h);        // X' = (X+Y)^2 - YY - XX
  fe_sub(f, zz2, g);        // T' = 2ZZ - Z'
  fe_mul(r.X, e, f);
  fe_mul(r.Y, h, g);
  fe_mul(r.Z, g, f);
  fe_mul(r.T, e, h);
}

ZKP_DEV void ge_neg(ge_ext& r, const ge_ext& p) {
  fe_neg(r.X, p.X);
  r.Y = p.Y;
  r.Z = p.Z;
  fe_neg(r.T, p.T);
}

// ristretto identity (coset-aware): X == 0 or Y == 0   (dalek ct_eq against the identity [ext])
ZKP_DEV uint32_t ge_is_identity_coset(const ge_ext& p) { return fe_is_zero(p.X) | fe_is_zero(p.Y); }

// (was_square, r) = sqrt_ratio_i(1, v)   RFC 9496 4.2 with u = 1
template <bool VT = false>
ZKP_DEV uint32_t fe_invsqrt(fe& r, const fe& v) {
  fe v3, v7, t, check, one, m1, mi;
  fe_sqx<VT>(v3, v);
  fe_mulx<VT>(v3, v3, v);
  fe_sqx<VT>(v7, v3);
  fe_mulx<VT>(v7, v7, v);
  fe_pow22523<VT>(t, v7);
  fe_mulx<VT>(t, t, v3);          // r = v^3 * (v^7)^((p-5)/8)
  fe_sqx<VT>(check, t);
  fe_mulx<VT>(check, check, v);   // v * r^2
  fe_one(one);
  fe_neg(m1, one);
  fe_neg(mi, fe_SQRT_M1());
  uint32_t correct = fe_eq(check, one);
  uint32_t flipped = fe_eq(check, m1);
  uint32_t flipped_i = fe_eq(check, mi);
  fe ti;
  fe_mul(ti, t, fe_SQRT_M1());
  fe_select(t, t, ti, flipped | flipped_i);
  fe_abs(r, t);
  return correct | flipped;
}

// (was_square, r) with r = |1/sqrt(v)| for a nonzero square v -- the decoder's use of sqrt_ratio_i(1, v).
// RFC 9496 4.2 evaluates r0 = v^3 (v^7)^((p-5)/8) = v^((p-5)/8) * zeta with zeta = (v^((p-1)/4))^3; for a square v,
// zeta = +-1, so v r0^2 = v (v^((p-5)/8))^2 and the two candidates agree up to the sign that the final |.| removes: the
// exponentiation can start from v itself (2 squarings and 3 multiplications fewer per point).  For v = 0 or a non-square
// both forms report was_square = 0 and the decoder rejects; r is then unspecified, so this form is NOT sqrt_ratio_i and
// is used by ristretto_decode only.
template <bool VT = false>
ZKP_DEV uint32_t fe_invsqrt_decode(fe& r, const fe& v) {
  fe t, check, one, m1, ti;
  fe_pow22523<VT>(t, v);
  fe_sqx<VT>(check, t);
  fe_mulx<VT>(check, check, v);   // v * t^2 = v^((p-1)/4): +1 or -1 for a nonzero square
  fe_one(one);
  fe_neg(m1, one);
  uint32_t correct = fe_eq(check, one);
  uint32_t flipped = fe_eq(check, m1);
  fe_mulx<VT>(ti, t, fe_SQRT_M1());
  fe_select(t, t, ti, flipped);
  fe_abs(r, t);
  return correct | flipped;
}

// RFC 9496 4.3.1.  w = the 32 encoding bytes as 8 little-endian words.  Returns 1 if valid.
// Output: affine extended coordinates (x, y, 1, t) delivered as x, y, t.
template <bool VT = false>
ZKP_DEV uint32_t ristretto_decode(fe& x, fe& y, fe& t, const uint32_t* w) {
  fe s, ss, u1, u2, u2sq, v, inv, dx, dy, one, tmp;
  fe_from_words(s, w);
  // canonical (s < p, bit 255 clear) and non-negative (even)
  uint32_t cw[8];
  fe_to_words(cw, s);
  uint32_t diff = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) diff |= cw[i] ^ w[i];
  uint32_t ok = (diff == 0) & ((w[0] & 1u) == 0);
  fe_one(one);
  fe_sqx<VT>(ss, s);
  fe_sub_t<VT>(u1, one, ss);
  fe_add_t<VT>(u2, one, ss);
  fe_sqx<VT>(u2sq, u2);
  fe_sqx<VT>(tmp, u1);
  fe_mulx<VT>(tmp, tmp, fe_D());
  fe_neg(tmp, tmp);
  fe_sub_t<VT>(v, tmp, u2sq);       // v = -(D*u1^2) - u2^2
  fe_mulx<VT>(tmp, v, u2sq);
  uint32_t sq = fe_invsqrt_decode<VT>(inv, tmp);
  fe_mulx<VT>(dx, inv, u2);
  fe_mulx<VT>(dy, inv, dx);
  fe_mulx<VT>(dy, dy, v);
  fe_mulx<VT>(tmp, s, dx);
  fe_add_t<VT>(tmp, tmp, tmp);
  fe_abs(x, tmp);
  fe_mulx<VT>(y, u1, dy);
  fe_mulx<VT>(t, x, y);
  ok &= sq & (fe_is_negative(t) ^ 1u) & (fe_is_zero(y) ^ 1u);
  return ok;
}

ZKP_DEV void ge_aniels_from_affine(ge_aniels& r, const fe& x, const fe& y, const fe& t) {
  fe_add(r.yplusx, y, x);
  fe_sub(r.yminusx, y, x);
  fe_mul(r.xy2d, t, fe_D2());
}

// RFC 9496 4.3.2: w = canonical 32-byte encoding as 8 words
ZKP_DEV void ristretto_encode(uint32_t* w, const ge_ext& p) {
  fe u1, u2, tmp, inv, den1, den2, zinv, ix0, iy0, ench, x, y, deninv, s;
  fe_add(u1, p.Z, p.Y);
  fe_sub(tmp, p.Z, p.Y);
  fe_mul(u1, u1, tmp);
  fe_mul(u2, p.X, p.Y);
  fe_sq(tmp, u2);
  fe_mul(tmp, tmp, u1);
  fe_invsqrt(inv, tmp);
  fe_mul(den1, inv, u1);
  fe_mul(den2, inv, u2);
  fe_mul(zinv, den1, den2);
  fe_mul(zinv, zinv, p.T);
  fe_mul(ix0, p.X, fe_SQRT_M1());
  fe_mul(iy0, p.Y, fe_SQRT_M1());
  fe_mul(ench, den1, fe_INVSQRT_A_MINUS_D());
  fe_mul(tmp, p.T, zinv);
  uint32_t rotate = fe_is_negative(tmp);
  fe_select(x, p.X, iy0, rotate);
  fe_select(y, p.Y, ix0, rotate);
  fe_select(deninv, den2, ench, rotate);
  fe_mul(tmp, x, zinv);
  fe_cneg(y, y, fe_is_negative(tmp));
  fe_sub(tmp, p.Z, y);
  fe_mul(s, deninv, tmp);
  fe_abs(s, s);
  fe_to_words(w, s);
}

}  // namespace zkp
