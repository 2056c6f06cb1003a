// Host build of the device arithmetic headers (-DZKP_HOST_EMUL): the asm carry chains are replaced by 64-bit C
// so the kernel *logic* can be checked against the big-int oracle without a GPU.  TEST INFRASTRUCTURE ONLY:
// nothing in zkp_b200/ loads this library.
#include <cstring>
#include "../../zkp_b200/csrc/fe.cuh"
#include "../../zkp_b200/csrc/fe64.cuh"
#include "../../zkp_b200/csrc/ge.cuh"
#include "../../zkp_b200/csrc/sc.cuh"
#include "../../zkp_b200/csrc/hash.cuh"
#include "../../zkp_b200/csrc/scl.cuh"
#include "../../zkp_b200/csrc/comb.cuh"
#include <vector>
using namespace zkp;

static fe ld(const uint8_t* p) { fe r; memcpy(r.v, p, 32); return r; }
static void st(uint8_t* p, const fe& a) { memcpy(p, a.v, 32); }

extern "C" {
// raw (non-canonical) 256-bit in/out so the weak-reduction invariants themselves are testable
void emul_fe_mul(uint8_t* r, const uint8_t* a, const uint8_t* b) { fe x; fe_mul(x, ld(a), ld(b)); st(r, x); }
void emul_fe_sq(uint8_t* r, const uint8_t* a) { fe x; fe_sq(x, ld(a)); st(r, x); }
void emul_fe_add(uint8_t* r, const uint8_t* a, const uint8_t* b) { fe x; fe_add(x, ld(a), ld(b)); st(r, x); }
void emul_fe_sub(uint8_t* r, const uint8_t* a, const uint8_t* b) { fe x; fe_sub(x, ld(a), ld(b)); st(r, x); }
// the variable-time tails (cold-branch carry fold) used by decompression and bucket accumulation
void emul_fe_mul_vt(uint8_t* r, const uint8_t* a, const uint8_t* b) { fe x; fe_mul_vt(x, ld(a), ld(b)); st(r, x); }
void emul_fe_sq_vt(uint8_t* r, const uint8_t* a) { fe x; fe_sq_vt(x, ld(a)); st(r, x); }
void emul_fe_add_vt(uint8_t* r, const uint8_t* a, const uint8_t* b) { fe x; fe_add_vt(x, ld(a), ld(b)); st(r, x); }
void emul_fe_sub_vt(uint8_t* r, const uint8_t* a, const uint8_t* b) { fe x; fe_sub_vt(x, ld(a), ld(b)); st(r, x); }
void emul_fe_pow22523_vt(uint8_t* r, const uint8_t* a) { fe x; fe_pow22523<true>(x, ld(a)); st(r, x); }
void emul_fe_canon(uint8_t* r, const uint8_t* a) { fe x; fe_canon(x, ld(a)); st(r, x); }
void emul_fe_invert(uint8_t* r, const uint8_t* a) { fe x; fe_invert(x, ld(a)); st(r, x); }
void emul_fe_pow22523(uint8_t* r, const uint8_t* a) { fe x; fe_pow22523(x, ld(a)); st(r, x); }
void emul_fe_from_limbs51(uint8_t* r, const uint64_t* l) { fe x; fe_from_limbs51(x, l); st(r, x); }
void emul_fe_to_limbs51(uint64_t* l, const uint8_t* a) { fe_to_limbs51(l, ld(a)); }

// ---- the FP64 twin of the field arithmetic (fe64.cuh): in/out as 32-byte integers; `max_q` (optional) receives the
// largest |limb / weight| seen in the result, for the exactness bounds
static double fe64_maxq(const fe64& x) {
  double m = 0;
  for (int i = 0; i < 12; i++) { double q = fabs(ldexp(x.v[i], -ZKP_FE64_OFF(i))); if (q > m) m = q; }
  return m;
}
void emul_fe64_roundtrip(uint8_t* r, const uint8_t* a) { fe64 x; fe64_from_fe(x, ld(a)); fe y; fe64_to_fe(y, x); st(r, y); }
double emul_fe64_sq(uint8_t* r, const uint8_t* a, int n) {
  fe64 x; fe64_from_fe(x, ld(a)); double m = 0;
  for (int i = 0; i < n; i++) { fe64_sq(x, x); double q = fe64_maxq(x); if (q > m) m = q; }
  fe y; fe64_to_fe(y, x); st(r, y); return m;
}
double emul_fe64_mul(uint8_t* r, const uint8_t* a, const uint8_t* b) {
  fe64 x, y, z; fe64_from_fe(x, ld(a)); fe64_from_fe(y, ld(b)); fe64_mul(z, x, y);
  fe o; fe64_to_fe(o, z); st(r, o); return fe64_maxq(z);
}
void emul_fe_pow22523_fp64(uint8_t* r, const uint8_t* a) { fe x; fe_pow22523_fp64(x, ld(a)); st(r, x); }
// limbs are integers times their weight, and columns stay below 2^53: returns the largest |column| / 2^(o_k) of a^2
// computed in long double beside the double computation (0 if any limb is not an integer multiple of its weight)
double emul_fe64_column_bound(const uint8_t* a) {
  fe64 x; fe64_from_fe(x, ld(a));
  double worst = 0;
  for (int rep = 0; rep < 3; rep++) {
    for (int i = 0; i < 12; i++) { double q = ldexp(x.v[i], -ZKP_FE64_OFF(i)); if (q != floor(q)) return 0; }
    for (int k = 0; k < 12; k++) {
      long double s = 0;
      for (int i = 0; i < 12; i++) for (int j = 0; j < 12; j++) {
        long double p = fabsl((long double)x.v[i] * x.v[j]);
        if (i + j == k) s += p;
        if (i + j == k + 12) s += p * 19.0L * ldexpl(1.0L, -255);
      }
      double b = (double)ldexpl(s, -ZKP_FE64_OFF(k));
      if (b > worst) worst = b;
    }
    fe64_sq(x, x);
  }
  return worst;
}

// ---- group / ristretto ----
static ge_ext ldp(const uint8_t* p) { ge_ext r; memcpy(&r, p, 128); return r; }
// decode: out = x|y|t (96 B raw limbs), returns validity
int emul_decode(uint8_t* out, const uint8_t* enc) {
  uint32_t w[8]; memcpy(w, enc, 32);
  fe x, y, t; uint32_t ok = ristretto_decode(x, y, t, w);
  memcpy(out, x.v, 32); memcpy(out + 32, y.v, 32); memcpy(out + 64, t.v, 32);
  return (int)ok;
}
int emul_decode_vt(uint8_t* out, const uint8_t* enc) {
  uint32_t w[8]; memcpy(w, enc, 32);
  fe x, y, t; uint32_t ok = ristretto_decode<true>(x, y, t, w);
  memcpy(out, x.v, 32); memcpy(out + 32, y.v, 32); memcpy(out + 64, t.v, 32);
  return (int)ok;
}
void emul_encode(uint8_t* enc, const uint8_t* pt128) {
  uint32_t w[8]; ristretto_encode(w, ldp(pt128)); memcpy(enc, w, 32);
}
// p (128 B ext) + [neg] q (x|y|t affine 96 B) via the Niels path
void emul_madd(uint8_t* out, const uint8_t* pt128, const uint8_t* aff96, int neg) {
  fe x, y, t; memcpy(x.v, aff96, 32); memcpy(y.v, aff96 + 32, 32); memcpy(t.v, aff96 + 64, 32);
  ge_aniels q; ge_aniels_from_affine(q, x, y, t); ge_aniels_cneg(q, (uint32_t)neg);
  ge_ext r; ge_madd(r, ldp(pt128), q); memcpy(out, &r, 128);
}
void emul_madd_signed(uint8_t* out, const uint8_t* pt128, const uint8_t* aff96, int neg) {
  fe x, y, t; memcpy(x.v, aff96, 32); memcpy(y.v, aff96 + 32, 32); memcpy(t.v, aff96 + 64, 32);
  ge_aniels q; ge_aniels_from_affine(q, x, y, t);
  ge_ext r; ge_madd_signed(r, ldp(pt128), q, (uint32_t)neg); memcpy(out, &r, 128);
}
void emul_madd_signed_vt(uint8_t* out, const uint8_t* pt128, const uint8_t* aff96, int neg) {
  fe x, y, t; memcpy(x.v, aff96, 32); memcpy(y.v, aff96 + 32, 32); memcpy(t.v, aff96 + 64, 32);
  ge_aniels q; ge_aniels_from_affine(q, x, y, t);
  ge_ext r; ge_madd_signed<true>(r, ldp(pt128), q, (uint32_t)neg); memcpy(out, &r, 128);
}
void emul_add(uint8_t* out, const uint8_t* p, const uint8_t* q) { ge_ext r; ge_add(r, ldp(p), ldp(q)); memcpy(out, &r, 128); }
void emul_double(uint8_t* out, const uint8_t* p) { ge_ext r; ge_double(r, ldp(p)); memcpy(out, &r, 128); }
int emul_is_identity(const uint8_t* p) { return (int)ge_is_identity_coset(ldp(p)); }
// scalar: returns canonical; k32 = folded magnitude; *neg; digits[W] signed
int emul_recode(int32_t* digits, uint8_t* k32, int* neg, const uint8_t* s32, int c, int W) {
  uint32_t s[8], k[8]; memcpy(s, s32, 32);
  uint32_t ng; uint32_t can = sc_fold_sign(k, ng, s);
  memcpy(k32, k, 32); *neg = (int)ng;
  uint32_t carry = 0;
  for (int w = 0; w < W; w++) { uint32_t mag, dn; sc_next_digit(mag, dn, carry, k, c); digits[w] = dn ? -(int32_t)mag : (int32_t)mag; }
  return (int)can | ((int)carry << 1);
}

// ---- device hashing / scalar arithmetic mod l (hash.cuh, scl.cuh) ----
void emul_keccak(uint64_t* st) { keccak_f1600_dev(st); }
// Merlin transcript on the device state machine: new(label); a sequence of appends; challenge
void emul_transcript_test(uint8_t* out32) {
  strobe_t s; for (int i = 0; i < 25; i++) s.st[i] = 0;
  uint8_t* b = st_bytes(s); const uint8_t init[6] = {1, 168, 1, 0, 1, 96};
  memcpy(b, init, 6); memcpy(b + 6, "STROBEv1.0.2", 12); keccak_f1600_dev(s.st);
  s.pos = 0; s.pos_begin = 0; s.cur_flags = 0;
  strobe_meta_ad(s, (const uint8_t*)"Merlin v1.0", 11, false);
  transcript_append(s, (const uint8_t*)"dom-sep", 7, (const uint8_t*)"test protocol", 13);
  transcript_append(s, (const uint8_t*)"step1", 5, (const uint8_t*)"some data", 9);
  uint8_t ch[32], big[1024]; memset(big, 0x63, 1024);
  for (int i = 0; i < 32; i++) {
    transcript_challenge(s, (const uint8_t*)"challenge", 9, ch, 32);
    transcript_append(s, (const uint8_t*)"bigdata", 7, big, 1024);
    transcript_append(s, (const uint8_t*)"challengedata", 13, ch, 32);
  }
  memcpy(out32, ch, 32);
}
// TranscriptRng on the device state machine (prover front end): Transcript("test protocol"), one append, then
// build_rng -> rekey with n_w witnesses of 32 bytes -> finalize(entropy) -> n_out draws of 64 bytes
void emul_transcript_rng(uint8_t* out, const uint8_t* witnesses, int n_w, const uint8_t* entropy32, int n_out) {
  strobe_t s; for (int i = 0; i < 25; i++) s.st[i] = 0;
  uint8_t* b = st_bytes(s); const uint8_t init[6] = {1, 168, 1, 0, 1, 96};
  memcpy(b, init, 6); memcpy(b + 6, "STROBEv1.0.2", 12); keccak_f1600_dev(s.st);
  s.pos = 0; s.pos_begin = 0; s.cur_flags = 0;
  strobe_meta_ad(s, (const uint8_t*)"Merlin v1.0", 11, false);
  transcript_append(s, (const uint8_t*)"dom-sep", 7, (const uint8_t*)"test protocol", 13);
  transcript_append(s, (const uint8_t*)"step1", 5, (const uint8_t*)"some data", 9);
  for (int i = 0; i < n_w; i++) rng_rekey_with_witness(s, (const uint8_t*)"", 0, witnesses + 32 * i, 32);
  rng_finalize(s, entropy32);
  for (int i = 0; i < n_out; i++) rng_fill_bytes(s, out + 64 * i, 64);
}
void emul_shake(uint8_t* out, uint32_t n, const uint8_t* msg, uint32_t len) { shake256_short(out, n, msg, len); }
void emul_scl_mul(uint8_t* r, const uint8_t* a, const uint8_t* b) { scl x, y, z; memcpy(x.v, a, 32); memcpy(y.v, b, 32); scl_mul(z, x, y); memcpy(r, z.v, 32); }
void emul_scl_mul_128(uint8_t* r, const uint8_t* a, const uint8_t* b) { scl x, y, z; memcpy(x.v, a, 32); memcpy(y.v, b, 32); scl_mul_128(z, x, y); memcpy(r, z.v, 32); }
void emul_scl_add(uint8_t* r, const uint8_t* a, const uint8_t* b) { scl x, y, z; memcpy(x.v, a, 32); memcpy(y.v, b, 32); scl_add(z, x, y); memcpy(r, z.v, 32); }
void emul_scl_sub(uint8_t* r, const uint8_t* a, const uint8_t* b) { scl x, y, z; memcpy(x.v, a, 32); memcpy(y.v, b, 32); scl_sub(z, x, y); memcpy(r, z.v, 32); }
void emul_scl_wide(uint8_t* r, const uint8_t* a64) { scl z; scl_from_wide(z, a64); memcpy(r, z.v, 32); }

// ---- signed four-tooth combs (comb.cuh): the constant-time MSM of batch proving ----
void emul_comb_recode(uint8_t* m32, const uint8_t* s32) {
  uint32_t s[8], m[8]; memcpy(s, s32, 32); comb_recode(m, s); memcpy(m32, m, 32);
}
// the eight entries of the comb of P as extended points would be redundant: return them as projective Niels (8 x 128 B)
void emul_comb_build(uint8_t* out1024, const uint8_t* pt128) {
  ge_pniels E[8]; comb_build(E, ldp(pt128)); memcpy(out1024, E, 1024);
}
// sum_t scalars[t] * points[t] the way k_small_msm_comb does it: recode, one comb per base, 64 columns of
// (double; per term: column digits -> constant-time select -> add).  scalars [n][32] canonical, points [n][128] extended.
void emul_comb_msm(uint8_t* out128, const uint8_t* scalars, const uint8_t* points, int n) {
  std::vector<ge_pniels> tabs((size_t)8 * n);
  std::vector<uint32_t> m((size_t)8 * n);
  for (int t = 0; t < n; t++) {
    comb_build(&tabs[(size_t)8 * t], ldp(points + 128 * (size_t)t));
    uint32_t s[8]; memcpy(s, scalars + 32 * (size_t)t, 32);
    comb_recode(&m[(size_t)8 * t], s);
  }
  ge_ext acc; ge_identity(acc);
  for (int col = 63; col >= 0; col--) {
    ge_double(acc, acc);
    for (int t = 0; t < n; t++) {
      uint32_t idx, neg;
      comb_column(idx, neg, &m[(size_t)8 * t], col);
      ge_pniels sel;
      const uint32_t* tab = (const uint32_t*)&tabs[(size_t)8 * t];
      comb_select(sel, [&](uint32_t e, int q, uint32_t& x, uint32_t& y, uint32_t& z, uint32_t& w) {
        const uint32_t* p = tab + 32 * e + 4 * q; x = p[0]; y = p[1]; z = p[2]; w = p[3];
      }, idx, neg);
      ge_add_pniels(acc, acc, sel);
    }
  }
  memcpy(out128, &acc, 128);
}
}
