#include "scalar.hpp"

namespace zkp_host {

typedef unsigned __int128 u128;
static const uint64_t L[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0, 0x1000000000000000ULL};

// r = a - b over n words, returns borrow
static uint64_t sub_n(uint64_t* r, const uint64_t* a, const uint64_t* b, int n) {
  uint64_t br = 0;
  for (int i = 0; i < n; i++) {
    u128 d = (u128)a[i] - b[i] - br;
    r[i] = (uint64_t)d;
    br = (uint64_t)(d >> 64) & 1;
  }
  return br;
}
static uint64_t add_n(uint64_t* r, const uint64_t* a, const uint64_t* b, int n) {
  uint64_t c = 0;
  for (int i = 0; i < n; i++) {
    u128 s = (u128)a[i] + b[i] + c;
    r[i] = (uint64_t)s;
    c = (uint64_t)(s >> 64);
  }
  return c;
}
static bool ge_l(const uint64_t* a) {  // a (4 words) >= l ?   (used on public data only: canonicity of wire scalars)
  uint64_t t[4];
  return sub_n(t, a, L, 4) == 0;
}
// Branch-free selects: the arithmetic below runs on witnesses and blindings (Prover::prove_impl, prover.rs:85-109), and
// the reference's curve25519-dalek Scalar arithmetic is constant time, so nothing here may branch on a value.
// r = a + (b & mask) over 4 words; mask is 0 or ~0
static void add_masked(uint64_t* r, const uint64_t* a, const uint64_t* b, uint64_t mask) {
  uint64_t m[4] = {b[0] & mask, b[1] & mask, b[2] & mask, b[3] & mask};
  add_n(r, a, m, 4);
}
// r = a - l if a >= l else a   (one conditional subtraction, selected by mask)
static void cond_sub_l(uint64_t* r, const uint64_t* a) {
  uint64_t t[4];
  const uint64_t keep = 0 - sub_n(t, a, L, 4);   // borrow -> a < l -> keep a
  for (int i = 0; i < 4; i++) r[i] = (a[i] & keep) | (t[i] & ~keep);
}

// Reduce an n-word (n <= 9) little-endian value mod l by binary long division on 64-bit words:
// process from the top word down, keeping a running remainder r < l (fits 253 bits), r = r * 2^64 + word mod l.
// r * 2^64 is reduced with the identity 2^252 = -c (mod l), c = l - 2^252 (125 bits).
static void reduce_words(uint64_t out[4], const uint64_t* in, int n) {
  static const uint64_t C[2] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL};  // c = l - 2^252
  uint64_t r[4] = {0, 0, 0, 0};
  for (int i = n - 1; i >= 0; i--) {
    // t = r * 2^64 + in[i]   (5 words, < 2^317)
    uint64_t t[5] = {in[i], r[0], r[1], r[2], r[3]};
    // split t = hi * 2^252 + lo ; t = lo - hi * c (mod l); hi < 2^65, hi*c < 2^190
    uint64_t lo[4] = {t[0], t[1], t[2], t[3] & 0x0fffffffffffffffULL};
    uint64_t hi0 = (t[3] >> 60) | (t[4] << 4), hi1 = t[4] >> 60;  // hi = hi1:hi0 (hi1 < 2)
    // prod = hi * c  (up to 3 words + a bit)
    uint64_t prod[4] = {0, 0, 0, 0};
    u128 p = (u128)hi0 * C[0];
    prod[0] = (uint64_t)p;
    u128 carry = p >> 64;
    p = (u128)hi0 * C[1] + carry;
    prod[1] = (uint64_t)p;
    prod[2] = (uint64_t)(p >> 64);
    {  // + (hi1 ? c : 0) * 2^64, selected by mask
      const uint64_t m1 = 0 - hi1;
      u128 s = (u128)prod[1] + (C[0] & m1);
      prod[1] = (uint64_t)s;
      s = (u128)prod[2] + (C[1] & m1) + (uint64_t)(s >> 64);
      prod[2] = (uint64_t)s;
      prod[3] = (uint64_t)(s >> 64);
    }
    // r = lo - prod (mod l): lo < 2^252 < l, prod < 2^191 < l, so lo - prod + (borrow ? l : 0) is in [0, l)
    uint64_t d[4];
    const uint64_t borrow = sub_n(d, lo, prod, 4);
    add_masked(d, d, L, 0 - borrow);
    memcpy(r, d, 32);
  }
  cond_sub_l(out, r);   // r < l already (see above); one fixed conditional subtraction keeps the bound explicit
}

Scalar Scalar::from_bytes_mod_order(const uint8_t b[32]) {
  uint64_t in[4];
  memcpy(in, b, 32);
  Scalar s;
  reduce_words(s.w, in, 4);
  return s;
}
Scalar Scalar::from_bytes_mod_order_wide(const uint8_t b[64]) {
  uint64_t in[8];
  memcpy(in, b, 64);
  Scalar s;
  reduce_words(s.w, in, 8);
  return s;
}
bool Scalar::from_canonical_bytes(Scalar* out, const uint8_t b[32]) {
  uint64_t in[4];
  memcpy(in, b, 32);
  if (ge_l(in)) return false;
  memcpy(out->w, in, 32);
  return true;
}
Scalar sc_add(const Scalar& a, const Scalar& b) {
  Scalar r;
  uint64_t t[4];
  add_n(t, a.w, b.w, 4);  // < 2^254: no carry out
  cond_sub_l(r.w, t);
  return r;
}
Scalar sc_sub(const Scalar& a, const Scalar& b) {
  Scalar r;
  const uint64_t borrow = sub_n(r.w, a.w, b.w, 4);
  add_masked(r.w, r.w, L, 0 - borrow);
  return r;
}
Scalar sc_neg(const Scalar& a) { return sc_sub(Scalar::zero(), a); }
Scalar sc_mul(const Scalar& a, const Scalar& b) {
  uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < 4; j++) {
      u128 p = (u128)a.w[i] * b.w[j] + t[i + j] + carry;
      t[i + j] = (uint64_t)p;
      carry = (uint64_t)(p >> 64);
    }
    t[i + 4] = carry;
  }
  Scalar r;
  reduce_words(r.w, t, 8);
  return r;
}

}  // namespace zkp_host
