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

// c-bit field of k starting at bit position pos (pos + c may run past 256: zero-extended).  c <= 24.
ZKP_DEV uint32_t sc_bits(const uint32_t* k, int pos, int c) {
  int wi = pos >> 5, sh = pos & 31;
  uint32_t lo = wi < 8 ? k[wi] : 0u;
  uint32_t hi = (wi + 1) < 8 ? k[wi + 1] : 0u;
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> sh) & ((1u << c) - 1u);
}

// Signed digit of window w (width c) with incoming carry; digits in (-2^(c-1), 2^(c-1)].
// Returns |digit| in `mag` (0..2^(c-1)), its sign in `dneg`, and updates carry.
ZKP_DEV void sc_digit(uint32_t& mag, uint32_t& dneg, uint32_t& carry, const uint32_t* k, int w, int c) {
  uint32_t raw = sc_bits(k, w * c, c) + carry;
  uint32_t half = 1u << (c - 1);
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

}  // namespace zkp
