/* oracle/c/ref_u64.c -- C restatement of the reference's CPU hot path.  TEST INFRASTRUCTURE / CPU BASELINE.
 *
 * This is the checker and the timed CPU baseline ("port"), never the product: only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__ may build or call it (see oracle/__init__.py).
 *
 * What it restates: curve25519-dalek 2.x with the default `u64_backend` (/root/reference/Cargo.toml:27,37) [ext]
 *   backend/serial/u64/field.rs      FieldElement51: radix-2^51, u128 products, pow2k, invert, sqrt_ratio_i
 *   backend/serial/curve_models      Extended / ProjectiveNiels / Completed add + double formulas
 *   ristretto.rs                     CompressedRistretto::decompress, RistrettoPoint::compress
 *   scalar.rs                        to_radix_16, non_adjacent_form(5), to_radix_2w
 *   backend/serial/scalar_mul/       straus.rs (CT radix-16; vartime NAF-5), pippenger.rs (w = 6/7/8)
 *   edwards.rs                       optional_multiscalar_mul: size < 190 -> Straus, else Pippenger
 *   backend/vector                   the 4-way vector field and point types of `simd_backend`: ref_ifma.h (included below)
 * as reached from /root/reference/src/toolbox/prover.rs:94, verifier.rs:90,97,162,164,
 * batch_verifier.rs:219-230, toolbox/mod.rs:180,204.  The crate sources are not under /root/reference
 * (Cargo dependency), so this follows the published algorithms; it is pinned by the same golden vectors as
 * the Python oracle (tests/test_oracle_c.py).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef struct { u64 v[5]; } fe;
#define M51 0x7ffffffffffffULL

static const fe FE_ZERO = {{0, 0, 0, 0, 0}};
static const fe FE_ONE = {{1, 0, 0, 0, 0}};
/* d, 2d, sqrt(-1), 1/sqrt(a-d) in radix 2^51 (filled by init from the byte constants) */
static fe FE_D, FE_D2, FE_SQRTM1, FE_INVSQRT_A_MINUS_D;

static void fe_add(fe* r, const fe* a, const fe* b) { for (int i = 0; i < 5; i++) r->v[i] = a->v[i] + b->v[i]; }
/* a - b with 16p added so limbs stay positive (dalek FieldElement51::sub) then weak reduce */
static void fe_reduce(fe* r, const u64* l) {
  u64 c0 = l[0] >> 51, c1 = l[1] >> 51, c2 = l[2] >> 51, c3 = l[3] >> 51, c4 = l[4] >> 51;
  r->v[0] = (l[0] & M51) + c4 * 19;
  r->v[1] = (l[1] & M51) + c0;
  r->v[2] = (l[2] & M51) + c1;
  r->v[3] = (l[3] & M51) + c2;
  r->v[4] = (l[4] & M51) + c3;
}
static void fe_sub(fe* r, const fe* a, const fe* b) {
  u64 l[5];
  l[0] = (a->v[0] + 36028797018963664ULL) - b->v[0];
  l[1] = (a->v[1] + 36028797018963952ULL) - b->v[1];
  l[2] = (a->v[2] + 36028797018963952ULL) - b->v[2];
  l[3] = (a->v[3] + 36028797018963952ULL) - b->v[3];
  l[4] = (a->v[4] + 36028797018963952ULL) - b->v[4];
  fe_reduce(r, l);
}
static void fe_neg(fe* r, const fe* a) { fe_sub(r, &FE_ZERO, a); }

static void fe_mul(fe* r, const fe* a, const fe* b) {
  const u64 *x = a->v, *y = b->v;
  u64 b1 = y[1] * 19, b2 = y[2] * 19, b3 = y[3] * 19, b4 = y[4] * 19;
  u128 c0 = (u128)x[0] * y[0] + (u128)x[4] * b1 + (u128)x[3] * b2 + (u128)x[2] * b3 + (u128)x[1] * b4;
  u128 c1 = (u128)x[1] * y[0] + (u128)x[0] * y[1] + (u128)x[4] * b2 + (u128)x[3] * b3 + (u128)x[2] * b4;
  u128 c2 = (u128)x[2] * y[0] + (u128)x[1] * y[1] + (u128)x[0] * y[2] + (u128)x[4] * b3 + (u128)x[3] * b4;
  u128 c3 = (u128)x[3] * y[0] + (u128)x[2] * y[1] + (u128)x[1] * y[2] + (u128)x[0] * y[3] + (u128)x[4] * b4;
  u128 c4 = (u128)x[4] * y[0] + (u128)x[3] * y[1] + (u128)x[2] * y[2] + (u128)x[1] * y[3] + (u128)x[0] * y[4];
  u64 o[5];
  c1 += (u64)(c0 >> 51); o[0] = (u64)c0 & M51;
  c2 += (u64)(c1 >> 51); o[1] = (u64)c1 & M51;
  c3 += (u64)(c2 >> 51); o[2] = (u64)c2 & M51;
  c4 += (u64)(c3 >> 51); o[3] = (u64)c3 & M51;
  u64 carry = (u64)(c4 >> 51); o[4] = (u64)c4 & M51;
  o[0] += carry * 19;
  o[1] += o[0] >> 51; o[0] &= M51;
  memcpy(r->v, o, sizeof o);
}
static void fe_pow2k(fe* r, const fe* a, int k) {
  u64 x[5];
  memcpy(x, a->v, sizeof x);
  while (k-- > 0) {
    u64 a3_19 = 19 * x[3], a4_19 = 19 * x[4];
    u128 c0 = (u128)x[0] * x[0] + 2 * ((u128)x[1] * a4_19 + (u128)x[2] * a3_19);
    u128 c1 = (u128)x[3] * a3_19 + 2 * ((u128)x[0] * x[1] + (u128)x[2] * a4_19);
    u128 c2 = (u128)x[1] * x[1] + 2 * ((u128)x[0] * x[2] + (u128)x[4] * a3_19);
    u128 c3 = (u128)x[4] * a4_19 + 2 * ((u128)x[0] * x[3] + (u128)x[1] * x[2]);
    u128 c4 = (u128)x[2] * x[2] + 2 * ((u128)x[0] * x[4] + (u128)x[1] * x[3]);
    c1 += (u64)(c0 >> 51); x[0] = (u64)c0 & M51;
    c2 += (u64)(c1 >> 51); x[1] = (u64)c1 & M51;
    c3 += (u64)(c2 >> 51); x[2] = (u64)c2 & M51;
    c4 += (u64)(c3 >> 51); x[3] = (u64)c3 & M51;
    u64 carry = (u64)(c4 >> 51); x[4] = (u64)c4 & M51;
    x[0] += carry * 19;
    x[1] += x[0] >> 51; x[0] &= M51;
  }
  memcpy(r->v, x, sizeof x);
}
static void fe_sq(fe* r, const fe* a) { fe_pow2k(r, a, 1); }

static void fe_frombytes(fe* r, const uint8_t* b) {
  u64 w[4];
  memcpy(w, b, 32);
  r->v[0] = w[0] & M51;
  r->v[1] = ((w[0] >> 51) | (w[1] << 13)) & M51;
  r->v[2] = ((w[1] >> 38) | (w[2] << 26)) & M51;
  r->v[3] = ((w[2] >> 25) | (w[3] << 39)) & M51;
  r->v[4] = (w[3] >> 12) & M51; /* bit 255 ignored */
}
static void fe_tobytes(uint8_t* out, const fe* a) {
  u64 l[5];
  fe t;
  fe_reduce(&t, a->v);
  memcpy(l, t.v, sizeof l);
  u64 q = (l[0] + 19) >> 51;
  q = (l[1] + q) >> 51; q = (l[2] + q) >> 51; q = (l[3] + q) >> 51; q = (l[4] + q) >> 51;
  l[0] += 19 * q;
  l[1] += l[0] >> 51; l[0] &= M51;
  l[2] += l[1] >> 51; l[1] &= M51;
  l[3] += l[2] >> 51; l[2] &= M51;
  l[4] += l[3] >> 51; l[3] &= M51;
  l[4] &= M51;
  u64 w[4];
  w[0] = l[0] | (l[1] << 51);
  w[1] = (l[1] >> 13) | (l[2] << 38);
  w[2] = (l[2] >> 26) | (l[3] << 25);
  w[3] = (l[3] >> 39) | (l[4] << 12);
  memcpy(out, w, 32);
}
static int fe_is_negative(const fe* a) { uint8_t b[32]; fe_tobytes(b, a); return b[0] & 1; }
static int fe_is_zero(const fe* a) {
  uint8_t b[32]; fe_tobytes(b, a);
  uint8_t o = 0; for (int i = 0; i < 32; i++) o |= b[i];
  return o == 0;
}
static int fe_eq(const fe* a, const fe* b) {
  uint8_t x[32], y[32]; fe_tobytes(x, a); fe_tobytes(y, b);
  return memcmp(x, y, 32) == 0;
}
static void fe_cneg(fe* r, const fe* a, int neg) { if (neg) fe_neg(r, a); else *r = *a; }

static void fe_pow22501(fe* t19, fe* t3, const fe* a) {
  fe t0, t1, t2, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18;
  fe_sq(&t0, a); fe_pow2k(&t1, &t0, 2); fe_mul(&t2, a, &t1); fe_mul(t3, &t0, &t2);
  fe_sq(&t4, t3); fe_mul(&t5, &t2, &t4); fe_pow2k(&t6, &t5, 5); fe_mul(&t7, &t6, &t5);
  fe_pow2k(&t8, &t7, 10); fe_mul(&t9, &t8, &t7); fe_pow2k(&t10, &t9, 20); fe_mul(&t11, &t10, &t9);
  fe_pow2k(&t12, &t11, 10); fe_mul(&t13, &t12, &t7); fe_pow2k(&t14, &t13, 50); fe_mul(&t15, &t14, &t13);
  fe_pow2k(&t16, &t15, 100); fe_mul(&t17, &t16, &t15); fe_pow2k(&t18, &t17, 50); fe_mul(t19, &t18, &t13);
}
static void fe_pow_p58(fe* r, const fe* a) {
  fe t19, t3, t20;
  fe_pow22501(&t19, &t3, a);
  fe_pow2k(&t20, &t19, 2);
  fe_mul(r, a, &t20);
}
/* (was_square, r) = sqrt_ratio_i(u, v) -- RFC 9496 4.2 */
static int fe_sqrt_ratio_i(fe* r, const fe* u, const fe* v) {
  fe v3, v7, t, check, mu, mui;
  fe_sq(&v3, v); fe_mul(&v3, &v3, v);
  fe_sq(&v7, &v3); fe_mul(&v7, &v7, v);
  fe_mul(&t, u, &v7); fe_pow_p58(&t, &t);
  fe_mul(&t, &t, &v3); fe_mul(&t, &t, u);
  fe_sq(&check, &t); fe_mul(&check, &check, v);
  fe_neg(&mu, u); fe_mul(&mui, &mu, &FE_SQRTM1);
  int correct = fe_eq(&check, u), flipped = fe_eq(&check, &mu), flipped_i = fe_eq(&check, &mui);
  if (flipped || flipped_i) fe_mul(&t, &t, &FE_SQRTM1);
  fe_cneg(r, &t, fe_is_negative(&t));
  return correct || flipped;
}

/* ---- group ---- */
typedef struct { fe X, Y, Z, T; } ge;            /* extended */
typedef struct { fe YpX, YmX, Z, T2d; } pniels;  /* projective Niels */
typedef struct { fe X, Y, Z, T; } completed;

static void ge_identity(ge* r) { r->X = FE_ZERO; r->Y = FE_ONE; r->Z = FE_ONE; r->T = FE_ZERO; }
static void ge_to_pniels(pniels* r, const ge* p) {
  fe_add(&r->YpX, &p->Y, &p->X); fe_sub(&r->YmX, &p->Y, &p->X); r->Z = p->Z; fe_mul(&r->T2d, &p->T, &FE_D2);
}
static void completed_to_ext(ge* r, const completed* c) {
  fe_mul(&r->X, &c->X, &c->T); fe_mul(&r->Y, &c->Y, &c->Z); fe_mul(&r->Z, &c->Z, &c->T); fe_mul(&r->T, &c->X, &c->Y);
}
static void ge_add_pn(ge* r, const ge* p, const pniels* q, int neg) {
  fe ypx, ymx, pp, mm, tt2d, zz, zz2; completed c;
  fe_add(&ypx, &p->Y, &p->X); fe_sub(&ymx, &p->Y, &p->X);
  if (!neg) { fe_mul(&pp, &ypx, &q->YpX); fe_mul(&mm, &ymx, &q->YmX); }
  else { fe_mul(&pp, &ypx, &q->YmX); fe_mul(&mm, &ymx, &q->YpX); }
  fe_mul(&tt2d, &p->T, &q->T2d); fe_mul(&zz, &p->Z, &q->Z); fe_add(&zz2, &zz, &zz);
  fe_sub(&c.X, &pp, &mm); fe_add(&c.Y, &pp, &mm);
  if (!neg) { fe_add(&c.Z, &zz2, &tt2d); fe_sub(&c.T, &zz2, &tt2d); }
  else { fe_sub(&c.Z, &zz2, &tt2d); fe_add(&c.T, &zz2, &tt2d); }
  completed_to_ext(r, &c);
}
static void ge_add(ge* r, const ge* p, const ge* q) { pniels n; ge_to_pniels(&n, q); ge_add_pn(r, p, &n, 0); }
static void ge_double(ge* r, const ge* p) {
  fe xx, yy, zz2, xpy, xpy2; completed c;
  fe_sq(&xx, &p->X); fe_sq(&yy, &p->Y); fe_sq(&zz2, &p->Z); fe_add(&zz2, &zz2, &zz2);
  fe_add(&xpy, &p->X, &p->Y); fe_sq(&xpy2, &xpy);
  fe_add(&c.Y, &yy, &xx); fe_sub(&c.Z, &yy, &xx); fe_sub(&c.X, &xpy2, &c.Y); fe_sub(&c.T, &zz2, &c.Z);
  completed_to_ext(r, &c);
}
static void ge_mul_pow2(ge* r, const ge* p, int k) { *r = *p; while (k-- > 0) ge_double(r, r); }

static int ristretto_decode(ge* r, const uint8_t* b) {
  fe s, ss, u1, u2, u2sq, v, I, dx, dy, t; uint8_t chk[32];
  fe_frombytes(&s, b); fe_tobytes(chk, &s);
  if (memcmp(chk, b, 32) != 0 || (b[0] & 1)) return 0;
  fe_sq(&ss, &s); fe_sub(&u1, &FE_ONE, &ss); fe_add(&u2, &FE_ONE, &ss); fe_sq(&u2sq, &u2);
  fe_sq(&t, &u1); fe_mul(&t, &t, &FE_D); fe_neg(&t, &t); fe_sub(&v, &t, &u2sq);
  fe_mul(&t, &v, &u2sq);
  int ok = fe_sqrt_ratio_i(&I, &FE_ONE, &t);
  fe_mul(&dx, &I, &u2); fe_mul(&dy, &I, &dx); fe_mul(&dy, &dy, &v);
  fe_mul(&t, &s, &dx); fe_add(&t, &t, &t); fe_cneg(&r->X, &t, fe_is_negative(&t));
  fe_mul(&r->Y, &u1, &dy); r->Z = FE_ONE; fe_mul(&r->T, &r->X, &r->Y);
  if (!ok || fe_is_negative(&r->T) || fe_is_zero(&r->Y)) return 0;
  return 1;
}
static void ristretto_encode(uint8_t* out, const ge* p) {
  fe u1, u2, t, inv, i1, i2, zinv, ix, iy, ench, x, y, den, s;
  fe_add(&u1, &p->Z, &p->Y); fe_sub(&t, &p->Z, &p->Y); fe_mul(&u1, &u1, &t);
  fe_mul(&u2, &p->X, &p->Y);
  fe_sq(&t, &u2); fe_mul(&t, &t, &u1);
  fe_sqrt_ratio_i(&inv, &FE_ONE, &t);
  fe_mul(&i1, &inv, &u1); fe_mul(&i2, &inv, &u2);
  fe_mul(&zinv, &i1, &i2); fe_mul(&zinv, &zinv, &p->T);
  fe_mul(&ix, &p->X, &FE_SQRTM1); fe_mul(&iy, &p->Y, &FE_SQRTM1); fe_mul(&ench, &i1, &FE_INVSQRT_A_MINUS_D);
  fe_mul(&t, &p->T, &zinv);
  if (fe_is_negative(&t)) { x = iy; y = ix; den = ench; } else { x = p->X; y = p->Y; den = i2; }
  fe_mul(&t, &x, &zinv);
  if (fe_is_negative(&t)) fe_neg(&y, &y);
  fe_sub(&t, &p->Z, &y); fe_mul(&s, &den, &t);
  fe_cneg(&s, &s, fe_is_negative(&s));
  fe_tobytes(out, &s);
}

/* ---- scalar recodings (scalar.rs) ---- */
static void to_radix_16(int8_t* d, const uint8_t* s) {
  for (int i = 0; i < 32; i++) { d[2 * i] = s[i] & 15; d[2 * i + 1] = (s[i] >> 4) & 15; }
  for (int i = 0; i < 63; i++) { int8_t c = (int8_t)((d[i] + 8) >> 4); d[i] -= (int8_t)(c << 4); d[i + 1] += c; }
}
static void naf5(int8_t* naf, const uint8_t* s) {
  u64 x[5] = {0, 0, 0, 0, 0};
  memcpy(x, s, 32);
  memset(naf, 0, 256);
  const u64 width = 32, mask = 31;
  int pos = 0; u64 carry = 0;
  while (pos < 256) {
    int idx = pos / 64, b = pos % 64;
    u64 buf = b < 59 ? (x[idx] >> b) : ((x[idx] >> b) | (x[idx + 1] << (64 - b)));
    u64 window = carry + (buf & mask);
    if ((window & 1) == 0) { pos += 1; continue; }
    if (window < width / 2) { carry = 0; naf[pos] = (int8_t)window; }
    else { carry = 1; naf[pos] = (int8_t)((int64_t)window - (int64_t)width); }
    pos += 5;
  }
}
static int radix_2w_digits(int w) { return (256 + w - 1) / w + (w == 8 ? 1 : 0); }
static void to_radix_2w(int8_t* digits, const uint8_t* s, int w) {
  u64 x[5] = {0, 0, 0, 0, 0};
  memcpy(x, s, 32);
  const u64 radix = 1ull << w, mask = radix - 1;
  int cnt = (256 + w - 1) / w; u64 carry = 0;
  memset(digits, 0, 43);
  for (int i = 0; i < cnt; i++) {
    int bit = i * w, idx = bit / 64, b = bit % 64;
    u64 buf = (b < 64 - w || idx == 3) ? (x[idx] >> b) : ((x[idx] >> b) | (x[idx + 1] << (64 - b)));
    u64 coef = carry + (buf & mask);
    carry = (coef + radix / 2) >> w;
    digits[i] = (int8_t)((int64_t)coef - (int64_t)(carry << w));
  }
  if (w == 8) digits[cnt] += (int8_t)carry; else digits[cnt - 1] += (int8_t)(carry << w);
}

/* ---- MSM algorithms ---- */
/* Straus::multiscalar_mul (constant-time schedule: every digit does a table scan + add) */
static void straus_ct(ge* out, const uint8_t* scalars, const ge* pts, size_t n) {
  pniels* tab = (pniels*)malloc(n * 8 * sizeof(pniels));
  int8_t* dig = (int8_t*)malloc(n * 64);
  for (size_t i = 0; i < n; i++) {
    ge m = pts[i];
    ge_to_pniels(&tab[8 * i], &m);
    for (int k = 1; k < 8; k++) { ge_add_pn(&m, &m, &tab[8 * i], 0); ge_to_pniels(&tab[8 * i + k], &m); }
    to_radix_16(dig + 64 * i, scalars + 32 * i);
  }
  ge q; ge_identity(&q);
  pniels idn; idn.YpX = FE_ONE; idn.YmX = FE_ONE; idn.Z = FE_ONE; idn.T2d = FE_ZERO;
  for (int j = 63; j >= 0; j--) {
    ge_mul_pow2(&q, &q, 4);
    for (size_t i = 0; i < n; i++) {
      int d = dig[64 * i + j], mag = d < 0 ? -d : d;
      pniels sel = idn;
      for (int e = 1; e <= 8; e++) if (mag == e) sel = tab[8 * i + e - 1]; /* dalek scans all 8 with ct select */
      ge_add_pn(&q, &q, &sel, d < 0);
    }
  }
  *out = q;
  free(tab); free(dig);
}
/* Straus::optional_multiscalar_mul (vartime, NAF-5, odd multiples) */
static void straus_vt(ge* out, const uint8_t* scalars, const ge* pts, size_t n) {
  pniels* tab = (pniels*)malloc(n * 8 * sizeof(pniels));
  int8_t* nafs = (int8_t*)malloc(n * 256);
  for (size_t i = 0; i < n; i++) {
    ge p2, m = pts[i]; pniels p2n;
    ge_double(&p2, &pts[i]); ge_to_pniels(&p2n, &p2);
    ge_to_pniels(&tab[8 * i], &m);
    for (int k = 1; k < 8; k++) { ge_add_pn(&m, &m, &p2n, 0); ge_to_pniels(&tab[8 * i + k], &m); }
    naf5(nafs + 256 * i, scalars + 32 * i);
  }
  ge r; ge_identity(&r);
  for (int i = 255; i >= 0; i--) {
    ge_double(&r, &r);
    for (size_t k = 0; k < n; k++) {
      int d = nafs[256 * k + i];
      if (d > 0) ge_add_pn(&r, &r, &tab[8 * k + d / 2], 0);
      else if (d < 0) ge_add_pn(&r, &r, &tab[8 * k + (-d) / 2], 1);
    }
  }
  *out = r;
  free(tab); free(nafs);
}
/* Pippenger::optional_multiscalar_mul */
static void pippenger(ge* out, const uint8_t* scalars, const ge* pts, size_t n) {
  int w = n < 500 ? 6 : (n < 800 ? 7 : 8);
  int max_digit = 1 << w, digits_count = radix_2w_digits(w), buckets_count = max_digit / 2;
  int8_t* dig = (int8_t*)malloc(n * 43);
  pniels* pn = (pniels*)malloc(n * sizeof(pniels));
  for (size_t i = 0; i < n; i++) { to_radix_2w(dig + 43 * i, scalars + 32 * i, w); ge_to_pniels(&pn[i], &pts[i]); }
  ge* buckets = (ge*)malloc(buckets_count * sizeof(ge));
  ge total; ge_identity(&total);
  for (int idx = digits_count - 1; idx >= 0; idx--) {
    for (int b = 0; b < buckets_count; b++) ge_identity(&buckets[b]);
    for (size_t i = 0; i < n; i++) {
      int d = dig[43 * i + idx];
      if (d > 0) ge_add_pn(&buckets[d - 1], &buckets[d - 1], &pn[i], 0);
      else if (d < 0) ge_add_pn(&buckets[-d - 1], &buckets[-d - 1], &pn[i], 1);
    }
    ge run = buckets[buckets_count - 1], acc = buckets[buckets_count - 1];
    for (int b = buckets_count - 2; b >= 0; b--) { ge_add(&run, &run, &buckets[b]); ge_add(&acc, &acc, &run); }
    if (idx != digits_count - 1) ge_mul_pow2(&total, &total, w);
    ge_add(&total, &total, &acc);
  }
  *out = total;
  free(dig); free(pn); free(buckets);
}

#include "ref_ifma.h"   /* the 4-way vector restatement (simd_backend); defines REF_HAVE_SIMD and ref_simd_available() */

static pthread_once_t init_once = PTHREAD_ONCE_INIT;
static void init_consts(void) {
  static const uint8_t d[32] = {0xa3, 0x78, 0x59, 0x13, 0xca, 0x4d, 0xeb, 0x75, 0xab, 0xd8, 0x41, 0x41, 0x4d, 0x0a, 0x70, 0x00,
                                0x98, 0xe8, 0x79, 0x77, 0x79, 0x40, 0xc7, 0x8c, 0x73, 0xfe, 0x6f, 0x2b, 0xee, 0x6c, 0x03, 0x52};
  static const uint8_t sm1[32] = {0xb0, 0xa0, 0x0e, 0x4a, 0x27, 0x1b, 0xee, 0xc4, 0x78, 0xe4, 0x2f, 0xad, 0x06, 0x18, 0x43, 0x2f,
                                  0xa7, 0xd7, 0xfb, 0x3d, 0x99, 0x00, 0x4d, 0x2b, 0x0b, 0xdf, 0xc1, 0x4f, 0x80, 0x24, 0x83, 0x2b};
  static const uint8_t isad[32] = {0xea, 0x40, 0x5d, 0x80, 0xaa, 0xfd, 0xc8, 0x99, 0xbe, 0x72, 0x41, 0x5a, 0x17, 0x16, 0x2f, 0x9d,
                                   0x40, 0xd8, 0x01, 0xfe, 0x91, 0x7b, 0xc2, 0x16, 0xa2, 0xfc, 0xaf, 0xcf, 0x05, 0x89, 0x6c, 0x78};
  fe_frombytes(&FE_D, d);
  fe_add(&FE_D2, &FE_D, &FE_D);
  { u64 l[5]; memcpy(l, FE_D2.v, sizeof l); fe_reduce(&FE_D2, l); }
  fe_frombytes(&FE_SQRTM1, sm1);
  fe_frombytes(&FE_INVSQRT_A_MINUS_D, isad);
}

/* ---- exported C API (ctypes) ---- */
/* optional_multiscalar_mul over encodings with dalek's dispatch; returns 0 ok / 1 invalid point (first_bad) */
/* simd = 0: the serial u64 backend; simd = 1: the vector backend (decompression is serial in both, as in the crate) */
static int msm_vartime_points(ge* out, const uint8_t* scalars, const uint8_t* points, size_t n, int64_t* first_bad, int simd) {
  ge* pts = (ge*)malloc((n ? n : 1) * sizeof(ge));
  for (size_t i = 0; i < n; i++)
    if (!ristretto_decode(&pts[i], points + 32 * i)) { if (first_bad) *first_bad = (int64_t)i; free(pts); return 1; }
#if REF_HAVE_SIMD
  if (simd) msm_dispatch_simd(out, scalars, pts, n); else
#endif
  if (n < 190) straus_vt(out, scalars, pts, n); else pippenger(out, scalars, pts, n);
  free(pts);
  return 0;
}
static int msm_vartime_any(const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out32, int64_t* first_bad, int simd) {
  pthread_once(&init_once, init_consts);
  ge r;
  int rc = msm_vartime_points(&r, scalars, points, n, first_bad, simd);
  if (rc) return rc;
  ristretto_encode(out32, &r);
  return 0;
}
int ref_msm_vartime(const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out32, int64_t* first_bad) {
  return msm_vartime_any(scalars, points, n, out32, first_bad, 0);
}
/* the same through the vector backend; -1 when this cpu cannot run it */
int ref_msm_vartime_simd(const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out32, int64_t* first_bad) {
  if (!ref_simd_available()) return -1;
  return msm_vartime_any(scalars, points, n, out32, first_bad, 1);
}

typedef struct { const uint8_t *s, *p; size_t n; ge out; int rc; int64_t bad; int simd; } shard_t;
static void* shard_run(void* a) {
  shard_t* s = (shard_t*)a;
  s->bad = -1;
  s->rc = msm_vartime_points(&s->out, s->s, s->p, s->n, &s->bad, s->simd);
  return NULL;
}
/* the same MSM sharded over `threads` host threads (independent sub-sums added at the end) */
static int msm_vartime_mt_any(const uint8_t* scalars, const uint8_t* points, size_t n, int threads, uint8_t* out32,
                              int64_t* first_bad, int simd);
int ref_msm_vartime_mt(const uint8_t* scalars, const uint8_t* points, size_t n, int threads, uint8_t* out32,
                       int64_t* first_bad) {
  return msm_vartime_mt_any(scalars, points, n, threads, out32, first_bad, 0);
}
int ref_msm_vartime_mt_simd(const uint8_t* scalars, const uint8_t* points, size_t n, int threads, uint8_t* out32,
                            int64_t* first_bad) {
  if (!ref_simd_available()) return -1;
  return msm_vartime_mt_any(scalars, points, n, threads, out32, first_bad, 1);
}
static int msm_vartime_mt_any(const uint8_t* scalars, const uint8_t* points, size_t n, int threads, uint8_t* out32,
                              int64_t* first_bad, int simd) {
  pthread_once(&init_once, init_consts);
  if (threads < 1) threads = 1;
  if ((size_t)threads > n) threads = n ? (int)n : 1;
  shard_t* sh = (shard_t*)calloc(threads, sizeof(shard_t));
  pthread_t* th = (pthread_t*)malloc(threads * sizeof(pthread_t));
  size_t per = n / threads, rem = n % threads, off = 0;
  for (int t = 0; t < threads; t++) {
    size_t cnt = per + ((size_t)t < rem ? 1 : 0);
    sh[t].s = scalars + 32 * off; sh[t].p = points + 32 * off; sh[t].n = cnt; sh[t].simd = simd;
    off += cnt;
    pthread_create(&th[t], NULL, shard_run, &sh[t]);
  }
  ge tot; ge_identity(&tot);
  int rc = 0; off = 0;
  for (int t = 0; t < threads; t++) {
    pthread_join(th[t], NULL);
    if (sh[t].rc && !rc) { rc = sh[t].rc; if (first_bad) *first_bad = (int64_t)off + sh[t].bad; }
    if (!sh[t].rc) ge_add(&tot, &tot, &sh[t].out);
    off += sh[t].n;
  }
  free(sh); free(th);
  if (rc) return rc;
  ristretto_encode(out32, &tot);
  return 0;
}

/* M constant-time MSMs + compress (prover.rs:93-103); points as encodings */
int ref_msm_ct_batched(const uint8_t* scalars, const uint8_t* points, const uint64_t* offsets, size_t M, uint8_t* out) {
  pthread_once(&init_once, init_consts);
  for (size_t j = 0; j < M; j++) {
    size_t lo = offsets[j], n = offsets[j + 1] - lo;
    ge* pts = (ge*)malloc((n ? n : 1) * sizeof(ge));
    for (size_t i = 0; i < n; i++) if (!ristretto_decode(&pts[i], points + 32 * (lo + i))) { free(pts); return 1; }
    ge r; straus_ct(&r, scalars + 32 * lo, pts, n);
    ristretto_encode(out + 32 * j, &r);
    free(pts);
  }
  return 0;
}
int ref_msm_vartime_batched(const uint8_t* scalars, const uint8_t* points, const uint64_t* offsets, size_t M,
                            uint8_t* out, uint8_t* valid) {
  for (size_t j = 0; j < M; j++) {
    size_t lo = offsets[j], n = offsets[j + 1] - lo;
    int rc = ref_msm_vartime(scalars + 32 * lo, points + 32 * lo, n, out + 32 * j, NULL);
    valid[j] = rc == 0;
    if (rc) memset(out + 32 * j, 0, 32);
  }
  return 0;
}
int ref_decompress(const uint8_t* enc, size_t n, uint64_t* limbs, uint8_t* valid) {
  pthread_once(&init_once, init_consts);
  for (size_t i = 0; i < n; i++) {
    ge p;
    valid[i] = (uint8_t)ristretto_decode(&p, enc + 32 * i);
    if (!valid[i]) ge_identity(&p);
    const fe* c[4] = {&p.X, &p.Y, &p.Z, &p.T};
    for (int k = 0; k < 4; k++) {
      uint8_t b[32]; fe t;
      fe_tobytes(b, c[k]); fe_frombytes(&t, b);
      memcpy(limbs + 20 * i + 5 * k, t.v, 40);
    }
  }
  return 0;
}
int ref_compress(const uint64_t* limbs, size_t n, uint8_t* enc) {
  pthread_once(&init_once, init_consts);
  for (size_t i = 0; i < n; i++) {
    ge p;
    memcpy(p.X.v, limbs + 20 * i, 40); memcpy(p.Y.v, limbs + 20 * i + 5, 40);
    memcpy(p.Z.v, limbs + 20 * i + 10, 40); memcpy(p.T.v, limbs + 20 * i + 15, 40);
    ristretto_encode(enc + 32 * i, &p);
  }
  return 0;
}
