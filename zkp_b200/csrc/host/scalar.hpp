// Scalars mod l = 2^252 + 27742317777372353535851937790883648493 on the host.
//
// Replaces curve25519-dalek 2.x `scalar.rs` (Scalar, Scalar52 arithmetic) [ext] for the operations the
// reference's callers perform: Neg (/root/reference/src/toolbox/verifier.rs:95,142),
// From<u128> (verifier.rs:153, batch_verifier.rs:179), mul / += / -= (verifier.rs:155-158, batch_verifier.rs:183-201),
// from_bytes_mod_order_wide (toolbox/mod.rs:226), s*c+b (prover.rs:108).  Wire form: 32 bytes little-endian canonical.
#pragma once
#include <stdint.h>
#include <string.h>

namespace zkp_host {

struct Scalar {
  uint64_t w[4];  // little-endian words, always fully reduced (< l)

  static Scalar zero() { Scalar s; memset(s.w, 0, 32); return s; }
  static Scalar from_u128(uint64_t lo, uint64_t hi) { Scalar s; s.w[0] = lo; s.w[1] = hi; s.w[2] = s.w[3] = 0; return s; }
  static Scalar from_bytes_mod_order(const uint8_t b[32]);
  static Scalar from_bytes_mod_order_wide(const uint8_t b[64]);
  static bool from_canonical_bytes(Scalar* out, const uint8_t b[32]);
  void to_bytes(uint8_t out[32]) const { memcpy(out, w, 32); }
  bool operator==(const Scalar& o) const { return memcmp(w, o.w, 32) == 0; }
};

Scalar sc_add(const Scalar& a, const Scalar& b);
Scalar sc_sub(const Scalar& a, const Scalar& b);
Scalar sc_neg(const Scalar& a);
Scalar sc_mul(const Scalar& a, const Scalar& b);
// a*b + c
inline Scalar sc_muladd(const Scalar& a, const Scalar& b, const Scalar& c) { return sc_add(sc_mul(a, b), c); }

}  // namespace zkp_host
