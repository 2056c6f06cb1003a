// Scalar decoding for the MSM kernels: canonical check, sign folding, signed window digits.
//
// Replaces the recodings of curve25519-dalek 2.x `scalar.rs` (to_radix_2w / to_radix_16 /
// non_adjacent_form) [ext] -- SURVEY.md section 2.2 row E8 -- that `Pippenger::optional_multiscalar_mul` and
// `Straus::*` apply to the scalars handed over at /root/reference/src/toolbox/batch_verifier.rs:219-228,
// verifier.rs:97,162 and prover.rs:94.  The digit set differs from dalek's (any window width c <= 24, and the
// scalar is first replaced by min(s, l - s) with the point negated, which halves the windows of the 128-bit
// weights -rho that batch_verifier.rs:183 produces); the group element computed is the same, and so is its
// canonical encoding.
#pragma once
#include <stdint.h>
#include "fe.cuh"

namespace zkp {

// l = 2^252 + 27742317777372353535851937790883648493, little-endian words
#define ZKP_L0 0x5cf5d3edu
#define ZKP_L1 0x5812631au
#define ZKP_L2 0xa2f79cd6u
#define ZKP_L3 0x14def9deu
#define ZKP_L7 0x10000000u

ZKP_DEV void sc_load_l(uint32_t* l) {
  l[0] = ZKP_L0; l[1] = ZKP_L1; l[2] = ZKP_L2; l[3] = ZKP_L3;
  l[4] = 0; l[5] = 0; l[6] = 0; l[7] = ZKP_L7;
}

// Input: s = 8 words.  Returns canonical (s < l).  On return k = min(s, l-s) (< 2^252) and neg = (k != s).
ZKP_DEV uint32_t sc_fold_sign(uint32_t* k, uint32_t& neg, const uint32_t* s) {
  uint32_t l[8], d[8], h[8];
  sc_load_l(l);
  uint32_t br = sub8(d, s, l);       // borrow <=> s < l
  uint32_t canonical = br;
  sub8(d, l, s);                     // d = l - s
  // neg if s > l - s  <=>  (l - s) - s borrows
  uint32_t lt = sub8(h, d, s);
  neg = lt & canonical;
  uint32_t m = 0u - neg;
#pragma unroll
  for (int i = 0; i < 8; i++) k[i] = s[i] ^ (m & (s[i] ^ d[i]));
  return canonical;
}

// Signed window digits, lowest window first: every call returns the digit of the lowest c bits of k (plus the incoming
// carry) in (-2^(c-1), 2^(c-1)] -- |digit| in `mag` (0..2^(c-1)), its sign in `dneg` -- and shifts k right by c bits
// (eight funnel shifts: no limb of k is ever indexed by a run-time value).  1 <= c <= 24.
ZKP_DEV uint32_t sc_shr_limb(uint32_t lo, uint32_t hi, int s) {
#if ZKP_DEVICE_ASM
  return __funnelshift_r(lo, hi, s);
#else
  return (lo >> s) | (hi << (32 - s));
#endif
}
ZKP_DEV void sc_next_digit(uint32_t& mag, uint32_t& dneg, uint32_t& carry, uint32_t* k, int c) {
  const uint32_t raw = (k[0] & ((1u << c) - 1u)) + carry;
#pragma unroll
  for (int i = 0; i < 7; i++) k[i] = sc_shr_limb(k[i], k[i + 1], c);
  k[7] >>= c;
  const uint32_t half = 1u << (c - 1);
  if (raw > half) {
    mag = (1u << c) - raw;
    dneg = 1;
    carry = 1;
  } else {
    mag = raw;
    dneg = 0;
    carry = 0;
  }
}
// nothing left: every further digit is zero
ZKP_DEV bool sc_digits_done(const uint32_t* k, uint32_t carry) {
  return (k[0] | k[1] | k[2] | k[3] | k[4] | k[5] | k[6] | k[7] | carry) == 0;
}

}  // namespace zkp
