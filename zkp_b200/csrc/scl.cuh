// Scalars mod l = 2^252 + c on the device (8 x 32-bit limbs, always fully reduced).
//
// SURVEY.md section 8f-2 ("next" row): the coefficient fold of /root/reference/src/toolbox/batch_verifier.rs:173-206 and
// the challenge reduction of toolbox/mod.rs:223-227 (Scalar::from_bytes_mod_order_wide, curve25519-dalek scalar.rs
// [ext]) next to the MSM.  42 multiplications per CMZ proof: small against the point arithmetic, but the front-end kernel
// they run in (k_bv_prepare2) is bound by the ALU pipe (Keccak: profiles/r02_front_end_ncu.md), and 64-bit C arithmetic
// spends four ALU instructions per limb product on carries.  The products and the carry chains therefore use the fused
// multiply-add chains of fe.cuh (IMAD.WIDE.X on the FMA pipe, carries in the condition code); with -DZKP_HOST_EMUL the
// same chains are 64-bit C, so the host-emulation tests run this very source.
#pragma once
#include <stdint.h>
#include "fe.cuh"

namespace zkp {

struct scl { uint32_t v[8]; };

// c = l - 2^252 (125 bits), little-endian limbs
ZKP_DEV uint32_t scl_c(int i) {
  const uint32_t C[4] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu};
  return C[i];
}
ZKP_DEV uint32_t scl_l(int i) {
  const uint32_t L[8] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0, 0, 0, 0x10000000u};
  return L[i];
}

// r[0..n) = a - b, returns borrow
template <int N>
ZKP_DEV uint32_t limbs_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  if (N == 8) return sub8(r, a, b);   // one carry chain (fe.cuh)
  uint64_t br = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint64_t d = (uint64_t)a[i] - b[i] - br;
    r[i] = (uint32_t)d;
    br = (d >> 32) & 1;
  }
  return (uint32_t)br;
}
template <int N>
ZKP_DEV uint32_t limbs_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  if (N == 8) return add8(r, a, b);
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint64_t s = (uint64_t)a[i] + b[i] + c;
    r[i] = (uint32_t)s;
    c = s >> 32;
  }
  return (uint32_t)c;
}
// r[0..NA+NB) = a * b
template <int NA, int NB>
ZKP_DEV void limbs_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#pragma unroll
  for (int i = 0; i < NA + NB; i++) r[i] = 0;
#pragma unroll
  for (int i = 0; i < NA; i++) {
    uint64_t carry = 0;
#pragma unroll
    for (int j = 0; j < NB; j++) {
      uint64_t p = (uint64_t)a[i] * b[j] + r[i + j] + carry;
      r[i + j] = (uint32_t)p;
      carry = p >> 32;
    }
    r[i + NB] = (uint32_t)carry;
  }
}

// ---- products on the multiply-add chains of fe.cuh -----------------------------------------------------------------------
// Limb products a_i * b_j land on position i + j: those on even positions accumulate in E (pairs (0,1)(2,3)...), those on odd
// positions in O (O[p] = position p + 1), so that every row is a chain of lo/hi pairs; x = E + (O << 32) at the end.
// x[0..12) = a[0..4) * b[0..8)
ZKP_DEV void limbs_mul_4x8(uint32_t* x, const uint32_t* a, const uint32_t* b) {
  uint32_t E[16], O[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { E[i] = 0; O[i] = 0; }
  mulw(E[0], E[1], b[0], a[0]); mulw(E[2], E[3], b[2], a[0]); mulw(E[4], E[5], b[4], a[0]); mulw(E[6], E[7], b[6], a[0]);
  mulw(O[0], O[1], b[1], a[0]); mulw(O[2], O[3], b[3], a[0]); mulw(O[4], O[5], b[5], a[0]); mulw(O[6], O[7], b[7], a[0]);
  O[8] = mad4(O, b[0], b[2], b[4], b[6], a[1]);            // positions 1 3 5 7
  E[10] = mad4(E + 2, b[1], b[3], b[5], b[7], a[1]);       // positions 2 4 6 8
  E[10] += mad4(E + 2, b[0], b[2], b[4], b[6], a[2]);      // positions 2 4 6 8
  O[10] = mad4(O + 2, b[1], b[3], b[5], b[7], a[2]);       // positions 3 5 7 9
  O[10] += mad4(O + 2, b[0], b[2], b[4], b[6], a[3]);      // positions 3 5 7 9
  mad4(E + 4, b[1], b[3], b[5], b[7], a[3]);               // positions 4 6 8 10 (the product fits in 12 limbs)
  uint32_t t[15];
  add15(t, E + 1, O);
  x[0] = E[0];
#pragma unroll
  for (int i = 0; i < 11; i++) x[1 + i] = t[i];
}
// x[0..9) = a[0..4) * b[0..5)
ZKP_DEV void limbs_mul_4x5(uint32_t* x, const uint32_t* a, const uint32_t* b) {
  uint32_t E[10], O[10];
#pragma unroll
  for (int i = 0; i < 10; i++) { E[i] = 0; O[i] = 0; }
  mulw(E[0], E[1], b[0], a[0]); mulw(E[2], E[3], b[2], a[0]); mulw(E[4], E[5], b[4], a[0]);
  mulw(O[0], O[1], b[1], a[0]); mulw(O[2], O[3], b[3], a[0]);
  O[6] = mad3(O, b[0], b[2], b[4], a[1]);                  // positions 1 3 5
  E[6] = mad2(E + 2, b[1], b[3], a[1]);                    // positions 2 4
  E[8] = mad3(E + 2, b[0], b[2], b[4], a[2]);              // positions 2 4 6
  O[6] += mad2(O + 2, b[1], b[3], a[2]);                   // positions 3 5
  mad3(O + 2, b[0], b[2], b[4], a[3]);                     // positions 3 5 7 (the product fits in 9 limbs)
  E[8] += mad2(E + 4, b[1], b[3], a[3]);                   // positions 4 6
  uint32_t t[8];
  add8(t, E + 1, O);
  x[0] = E[0];
#pragma unroll
  for (int i = 0; i < 8; i++) x[1 + i] = t[i];
}
// x[0..5) = z * c[0..4)
ZKP_DEV void limbs_mul_1x4(uint32_t* x, uint32_t z, const uint32_t* c) {
  uint32_t E[9], O[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { E[i] = 0; O[i] = 0; }
  E[8] = 0;
  mulw(E[0], E[1], c[0], z); mulw(E[2], E[3], c[2], z);
  mulw(O[0], O[1], c[1], z); mulw(O[2], O[3], c[3], z);
  uint32_t t[8];
  add8(t, E + 1, O);
  x[0] = E[0];
#pragma unroll
  for (int i = 0; i < 4; i++) x[1 + i] = t[i];
}

// conditional r -= l while r >= l (r < 4l on entry), r has 8 limbs (+ optional 9th handled by caller)
template <int TIMES = 4>
ZKP_DEV void scl_final_sub(uint32_t* r) {
  uint32_t l[8], t[8];
#pragma unroll
  for (int i = 0; i < 8; i++) l[i] = scl_l(i);
#pragma unroll 1
  for (int k = 0; k < TIMES; k++) {
    uint32_t br = limbs_sub<8>(t, r, l);
    uint32_t keep = 0u - br;  // borrow -> keep r
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = (r[i] & keep) | (t[i] & ~keep);
  }
}

// x (16 limbs, any 512-bit value) mod l
ZKP_DEV void scl_reduce512(scl& out, const uint32_t* x) {
  // x = xh * 2^252 + xl
  uint32_t xl[8], xh[9];
#pragma unroll
  for (int i = 0; i < 8; i++) xl[i] = x[i];
  xl[7] &= 0x0fffffffu;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    uint32_t lo = x[7 + i] >> 28;
    uint32_t hi = (7 + i + 1 < 16) ? (x[8 + i] << 4) : 0u;
    xh[i] = lo | hi;
  }
  uint32_t c[4];
#pragma unroll
  for (int i = 0; i < 4; i++) c[i] = scl_c(i);
  // y = c * xh (13 limbs, < 2^385) = yh * 2^252 + yl
  uint32_t y[13], y12[12], top[5], pad[8], sum[8];
  limbs_mul_4x8(y12, c, xh);              // c * xh[0..8)
  limbs_mul_1x4(top, xh[8], c);           // c * xh[8], at limb 8
#pragma unroll
  for (int i = 0; i < 8; i++) {
    y[i] = y12[i];
    pad[i] = i < 4 ? y12[8 + i] : 0u;
    sum[i] = i < 5 ? top[i] : 0u;
  }
  add8(sum, sum, pad);
#pragma unroll
  for (int i = 0; i < 5; i++) y[8 + i] = sum[i];
  uint32_t yl[8], yh[5];
#pragma unroll
  for (int i = 0; i < 8; i++) yl[i] = y[i];
  yl[7] &= 0x0fffffffu;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint32_t lo = y[7 + i] >> 28;
    uint32_t hi = (8 + i < 13) ? (y[8 + i] << 4) : 0u;
    yh[i] = lo | hi;
  }
  // z = c * yh (9 limbs, < 2^258) = zh * 2^252 + zl, zh < 2^6
  uint32_t z[9];
  limbs_mul_4x5(z, c, yh);
  uint32_t zl[8];
#pragma unroll
  for (int i = 0; i < 8; i++) zl[i] = z[i];
  zl[7] &= 0x0fffffffu;
  uint32_t zh = (z[7] >> 28) | (z[8] << 4);
  // w = c * zh (5 limbs, < 2^131)
  uint32_t w[8], w5[5];
  limbs_mul_1x4(w5, zh, c);
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = i < 5 ? w5[i] : 0u;
  // x = xl - yl + zl - w  (mod l); every term < 2^252 < l: r = (xl + zl) + 2l - yl - w  in [0, 4l)
  uint32_t r[8], l2[8];
  limbs_add<8>(r, xl, zl);  // < 2^253
  uint64_t cy = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {  // 2l
    uint64_t s = ((uint64_t)scl_l(i) << 1) + cy;
    l2[i] = (uint32_t)s;
    cy = s >> 32;
  }
  limbs_add<8>(r, r, l2);   // < 2^253 + 2^254 < 2^255
  limbs_sub<8>(r, r, yl);
  limbs_sub<8>(r, r, w);    // still >= 0: 2l > yl + w
  scl_final_sub<3>(r);      // r < 4l
#pragma unroll
  for (int i = 0; i < 8; i++) out.v[i] = r[i];
}

// x (12 limbs, < 2^381: the product of a 128-bit weight and a reduced scalar) mod l -- the common case of the
// coefficient fold (batch_verifier.rs:183-201: every product has the 128-bit rho as one factor), with 24 instead of 60
// limb products in the reduction
ZKP_DEV void scl_reduce384(scl& out, const uint32_t* x) {
  // x = xh * 2^252 + xl, xh < 2^129 (5 limbs)
  uint32_t xl[8], xh[5];
#pragma unroll
  for (int i = 0; i < 8; i++) xl[i] = x[i];
  xl[7] &= 0x0fffffffu;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint32_t lo = x[7 + i] >> 28;
    uint32_t hi = (8 + i < 12) ? (x[8 + i] << 4) : 0u;
    xh[i] = lo | hi;
  }
  uint32_t c[4];
#pragma unroll
  for (int i = 0; i < 4; i++) c[i] = scl_c(i);
  // y = c * xh (9 limbs, < 2^254) = yh * 2^252 + yl, yh < 4
  uint32_t y[9];
  limbs_mul_4x5(y, c, xh);
  uint32_t yl[8];
#pragma unroll
  for (int i = 0; i < 8; i++) yl[i] = y[i];
  yl[7] &= 0x0fffffffu;
  uint32_t yh = (y[7] >> 28) | (y[8] << 4);
  // z = c * yh (5 limbs, < 2^127)
  uint32_t z[8], z5[5];
  limbs_mul_1x4(z5, yh, c);
#pragma unroll
  for (int i = 0; i < 8; i++) z[i] = i < 5 ? z5[i] : 0u;
  // x = xl - yl + z (mod l): r = xl + z + l - yl  in [0, 3l)
  uint32_t r[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; i++) l[i] = scl_l(i);
  limbs_add<8>(r, xl, z);   // < 2^252 + 2^127
  limbs_add<8>(r, r, l);    // < 2^254
  limbs_sub<8>(r, r, yl);   // >= 0: l > yl
  scl_final_sub<2>(r);      // r < 3l
#pragma unroll
  for (int i = 0; i < 8; i++) out.v[i] = r[i];
}
// r = a * b with a < 2^128 (its four high limbs are zero) and b reduced
ZKP_DEV void scl_mul_128(scl& r, const scl& a, const scl& b) {
  uint32_t x[12];
  limbs_mul_4x8(x, a.v, b.v);
  scl_reduce384(r, x);
}

ZKP_DEV void scl_from_wide(scl& out, const uint8_t* b64) {
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; i++)
    x[i] = (uint32_t)b64[4 * i] | ((uint32_t)b64[4 * i + 1] << 8) | ((uint32_t)b64[4 * i + 2] << 16) |
           ((uint32_t)b64[4 * i + 3] << 24);
  scl_reduce512(out, x);
}
ZKP_DEV void scl_mul(scl& r, const scl& a, const scl& b) {
  uint32_t x[16];
  limbs_mul<8, 8>(x, a.v, b.v);
  scl_reduce512(r, x);
}
ZKP_DEV void scl_add(scl& r, const scl& a, const scl& b) {
  uint32_t t[8], u[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; i++) l[i] = scl_l(i);
  limbs_add<8>(t, a.v, b.v);  // < 2^254: no carry
  uint32_t br = limbs_sub<8>(u, t, l);
  uint32_t keep = 0u - br;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = (t[i] & keep) | (u[i] & ~keep);
}
ZKP_DEV void scl_sub(scl& r, const scl& a, const scl& b) {
  uint32_t t[8], u[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; i++) l[i] = scl_l(i);
  uint32_t br = limbs_sub<8>(t, a.v, b.v);
  limbs_add<8>(u, t, l);
  uint32_t fix = 0u - br;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = (u[i] & fix) | (t[i] & ~fix);
}
ZKP_DEV void scl_zero(scl& r) {
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
}
ZKP_DEV void scl_neg(scl& r, const scl& a) {
  scl z;
  scl_zero(z);
  scl_sub(r, z, a);
}
// canonical check of 8 words
ZKP_DEV uint32_t scl_is_canonical(const uint32_t* w) {
  uint32_t t[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; i++) l[i] = scl_l(i);
  return limbs_sub<8>(t, w, l);  // borrow <=> w < l
}

}  // namespace zkp
