// Signed four-tooth combs for the constant-time MSMs of batch proving (second half).
//
// The reference's prover computes every commitment with Straus::multiscalar_mul (curve25519-dalek 2.x
// backend/serial/scalar_mul/straus.rs [ext], reached from /root/reference/src/toolbox/prover.rs:94-97): 256 doublings
// per MSM plus 64 table additions per term.  In a statement like CMZ'13 (benches/zkp.rs:27-46) the same base serves many
// constraints of a proof (P in ten of them) or every proof of the batch (X_1..X_10, A), so a table that is built ONCE
// per base can carry more: with the teeth T_x = 2^(64 x) P, x = 0..3, and the scalar written in signed digits
//   k' = sum_{i < 256} s_i 2^i,  s_i in {-1, +1}        (k' = k for odd k, k + l for even k: l P is 4-torsion, which
//                                                         the ristretto encoding does not see)
// column j = 0..63 of the digit matrix contributes  s_j T_0 + s_{j+64} T_1 + s_{j+128} T_2 + s_{j+192} T_3,  one of 16
// points that come in +- pairs: eight table entries.  An MSM then costs 64 doublings (instead of 256) and one
// constant-time table addition per term and column (as before); a comb costs 195 doublings and 10 additions, once.
// Everything is straight-line in the scalar: the digits only ever feed arithmetic masks (comb_select scans all eight
// entries), the Edwards formulas are complete, and no exceptional case needs a branch.
// Outputs are ristretto encodings, so they equal the reference's byte for byte whatever the digit set.
//
// This header holds the parts that do not touch device memory, so tests/host_emul checks them against the oracle; the
// kernels are in small_msm.cuh (k_build_combs, k_comb_recode, k_small_msm_comb).
#pragma once
#include "ge.cuh"
#include "sc.cuh"

namespace zkp {

// Sign bits of the signed-digit form of a canonical scalar s < l: bit i of m is 1 iff s_i = +1.
//   k' = s + (s even ? l : 0)  is odd and < 2^254;   sum_i (2 m_i - 1) 2^i = 2 m - (2^256 - 1) = k'   <=>   m = (k' - 1) / 2 + 2^255
ZKP_DEV void comb_recode(uint32_t* m, const uint32_t* s) {
  uint32_t l[8], t[8];
  sc_load_l(l);
  const uint32_t even_mask = (s[0] & 1u) - 1u;   // all ones iff s is even
#pragma unroll
  for (int i = 0; i < 8; i++) l[i] &= even_mask;
  add8(t, s, l);                                 // odd, < 2^254
#pragma unroll
  for (int i = 0; i < 7; i++) m[i] = (t[i] >> 1) | (t[i + 1] << 31);   // (k' - 1) >> 1 == k' >> 1 for odd k'
  m[7] = (t[7] >> 1) | 0x80000000u;
}

// Column j (0..63): the table index given by the three low teeth and whether the entry is negated (top digit -1:
// -(T_3 + sum (-s_x) T_x) is the negated entry of the complemented index).
// (w0..w3 = the words of m that hold bit j of each tooth: m[j >> 5 + 2 x]; sh = j & 31)
ZKP_DEV void comb_column_words(uint32_t& idx, uint32_t& neg, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int sh) {
  const uint32_t b0 = (w0 >> sh) & 1u, b1 = (w1 >> sh) & 1u, b2 = (w2 >> sh) & 1u, b3 = (w3 >> sh) & 1u;
  neg = b3 ^ 1u;
  idx = (b0 | (b1 << 1) | (b2 << 2)) ^ ((0u - neg) & 7u);
}
ZKP_DEV void comb_column(uint32_t& idx, uint32_t& neg, const uint32_t* m, int j) {
  const int w = j >> 5;
  comb_column_words(idx, neg, m[w], m[w + 2], m[w + 4], m[w + 6], j & 31);
}

// The eight entries E[idx] = T_3 + sum_{x < 3} (2 bit_x(idx) - 1) T_x of the comb of P.
ZKP_DEV void comb_build(ge_pniels* E, const ge_ext& P) {
  ge_ext t[4];
  t[0] = P;
#pragma unroll 1
  for (int x = 1; x < 4; x++) {
    t[x] = t[x - 1];
#pragma unroll 1
    for (int i = 0; i < 64; i++) ge_double(t[x], t[x]);
  }
  ge_ext e[8];
  ge_pniels n, d2[3];
  e[0] = t[3];
#pragma unroll 1
  for (int x = 0; x < 3; x++) {
    ge_to_pniels(n, t[x]);
    ge_pniels_cneg(n, 1u);
    ge_add_pniels(e[0], e[0], n);          // T_3 - T_0 - T_1 - T_2
    ge_ext u;
    ge_double(u, t[x]);
    ge_to_pniels(d2[x], u);                // flipping digit x from -1 to +1 adds 2 T_x
  }
#pragma unroll 1
  for (int idx = 1; idx < 8; idx++) {
    const int x = (idx & 1) ? 0 : ((idx & 2) ? 1 : 2);   // lowest set bit
    ge_add_pniels(e[idx], e[idx & (idx - 1)], d2[x]);
  }
#pragma unroll 1
  for (int idx = 0; idx < 8; idx++) ge_to_pniels(E[idx], e[idx]);
}

// Constant-time selection: every entry is read, the one with index idx is kept by masking, then negated if neg.
// word4(e, q, x, y, z, w) loads words 4q .. 4q+3 of entry e; an entry is Q4 groups of four words.
template <int Q4, class LOAD>
ZKP_DEV void comb_select_words(uint32_t* dst, LOAD&& word4, uint32_t idx) {
#pragma unroll
  for (int i = 0; i < 4 * Q4; i++) dst[i] = 0u;
#pragma unroll 1
  for (uint32_t e = 0; e < 8; e++) {
    const uint32_t mask = 0u - (uint32_t)(idx == e);
#pragma unroll
    for (int q = 0; q < Q4; q++) {
      uint32_t x, y, z, w;
      word4(e, q, x, y, z, w);
      dst[4 * q + 0] |= mask & x;
      dst[4 * q + 1] |= mask & y;
      dst[4 * q + 2] |= mask & z;
      dst[4 * q + 3] |= mask & w;
    }
  }
}
// projective Niels entries (32 words: Y+X, Y-X, Z, 2dT): the per-proof combs
template <class LOAD>
ZKP_DEV void comb_select(ge_pniels& sel, LOAD&& word4, uint32_t idx, uint32_t neg) {
  comb_select_words<8>((uint32_t*)&sel, word4, idx);
  ge_pniels_cneg(sel, neg);
}
// affine Niels entries (24 words: y+x, y-x, 2dxy): the shared combs of the batch-static bases, normalised once per
// batch so that their additions are mixed ones (7 M instead of 8 M) and their scans a quarter shorter
template <class LOAD>
ZKP_DEV void comb_select_affine(ge_aniels& sel, LOAD&& word4, uint32_t idx, uint32_t neg) {
  comb_select_words<6>((uint32_t*)&sel, word4, idx);
  ge_aniels_cneg(sel, neg);
}
// projective -> affine Niels: every component times 1 / Z (T2d = 2d X Y / Z, so T2d / Z = 2d x y)
ZKP_DEV void comb_entry_to_affine(ge_aniels& r, const ge_pniels& q) {
  fe zinv;
  fe_invert(zinv, q.Z);
  fe_mul(r.yplusx, q.YplusX, zinv);
  fe_mul(r.yminusx, q.YminusX, zinv);
  fe_mul(r.xy2d, q.T2d, zinv);
}

}  // namespace zkp
