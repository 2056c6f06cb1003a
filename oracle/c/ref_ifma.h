/* oracle/c/ref_ifma.h -- 4-way vector restatement of the reference's `simd_backend`.  TEST INFRASTRUCTURE / CPU BASELINE.
 * Included at the end of ref_u64.c (it reuses that file's serial field, decode/encode and scalar recodings).
 *
 * What it restates: curve25519-dalek 2.x `backend/vector` [ext] -- the backend /root/reference/Cargo.toml:40
 * (`simd_backend = ["nightly", "curve25519-dalek/simd_backend"]`) and /root/reference/README.md:69-74 select.  Its
 * design: the FOUR coordinates of one point live in the four 64-bit lanes of a vector, a
 * field multiplication is done on all four at once, and the Hisil-Wong-Carter-Dawson point formulas are arranged so
 * that one addition is two vector multiplications and one doubling is one vector squaring plus one vector
 * multiplication, with lane shuffles in between; `scalar_mul/{pippenger,straus}.rs` run unchanged on top of those
 * point types, and decompression stays serial (it has no 4-way parallelism inside one point).  The crate has two
 * vector fields, AVX2 (10 limbs of 25.5 bits, vpmuludq) and AVX-512 IFMA (5 limbs of 51 bits, vpmadd52{l,h}uq); this
 * file restates the IFMA one -- the faster of the two, so the stronger baseline -- on 256-bit vectors.  The crate
 * sources are not under /root/reference: the lane layout and the order of the shuffles are this file's own, the
 * arithmetic they implement is the published one, and the result is pinned bit-for-bit against the serial port
 * (tests/test_oracle.py) through the canonical encoding of the MSM result.
 *
 * Built with a target pragma, so the library loads on any x86-64; ref_simd_available() says whether THIS cpu can run
 * it (avx2 + avx512f + avx512vl + avx512ifma), and callers fall back to the serial port when it cannot.
 */
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
/* compiled for the baseline ISA (it runs on every cpu): everything after the pragma below must not */
int ref_simd_available(void) {
  __builtin_cpu_init();
  return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl") &&
         __builtin_cpu_supports("avx512ifma");
}
#pragma GCC push_options
#pragma GCC target("avx2,avx512f,avx512vl,avx512ifma")

/* lane j of v[k] = limb k (radix 2^51) of coordinate j */
typedef struct { __m256i v[5]; } fe4;

#define V_SET1(x) _mm256_set1_epi64x((long long)(x))
#define V_ADD(a, b) _mm256_add_epi64(a, b)
#define V_SUB(a, b) _mm256_sub_epi64(a, b)
/* PERM(v, a, b, c, d): destination lanes 0..3 take source lanes a, b, c, d */
#define V_PERM(v, a, b, c, d) _mm256_permute4x64_epi64(v, (a) | ((b) << 2) | ((c) << 4) | ((d) << 6))
/* BLEND(x, y, lanes): the 64-bit lanes named in `lanes` come from y, the others from x */
#define LANE0 0x03
#define LANE1 0x0c
#define LANE2 0x30
#define LANE3 0xc0
#define V_BLEND(x, y, lanes) _mm256_blend_epi32(x, y, lanes)

/* One parallel carry round: limbs < 2^63 in, limbs < 2^51 + 19 * 2^12 out (what a multiplication accepts: < 2^52). */
static inline void fe4_carry(fe4* r) {
  const __m256i m = V_SET1(M51);
  __m256i c0 = _mm256_srli_epi64(r->v[0], 51), c1 = _mm256_srli_epi64(r->v[1], 51), c2 = _mm256_srli_epi64(r->v[2], 51),
          c3 = _mm256_srli_epi64(r->v[3], 51), c4 = _mm256_srli_epi64(r->v[4], 51);
  c4 = _mm256_mul_epu32(c4, V_SET1(19));   /* c4 < 2^13 */
  r->v[0] = V_ADD(_mm256_and_si256(r->v[0], m), c4);
  r->v[1] = V_ADD(_mm256_and_si256(r->v[1], m), c0);
  r->v[2] = V_ADD(_mm256_and_si256(r->v[2], m), c1);
  r->v[3] = V_ADD(_mm256_and_si256(r->v[3], m), c2);
  r->v[4] = V_ADD(_mm256_and_si256(r->v[4], m), c3);
}

/* x * 19 for 64-bit lanes below 2^59 */
static inline __m256i v_mul19(__m256i x) {
  return V_ADD(V_ADD(_mm256_slli_epi64(x, 4), _mm256_slli_epi64(x, 1)), x);
}

/* Reduction shared by mul and sq: t[0..9] are the ten product limbs (each < 2^57), radix 2^51. */
static inline void fe4_fold(fe4* r, const __m256i* t) {
  for (int k = 0; k < 5; k++) r->v[k] = V_ADD(t[k], v_mul19(t[k + 5]));
  fe4_carry(r);
}

/* r = x * y lane-wise in GF(2^255-19); limbs of x and y below 2^52.
 * x_i * y_j = hi * 2^52 + lo with lo, hi from vpmadd52{l,h}uq; in radix 2^51 the hi half weighs 2 * 2^(51 (i+j+1)). */
static inline void fe4_mul(fe4* r, const fe4* x, const fe4* y) {
  const __m256i z = _mm256_setzero_si256();
  __m256i lo[10] = {z, z, z, z, z, z, z, z, z, z}, hi[10] = {z, z, z, z, z, z, z, z, z, z};
#pragma GCC unroll 5
  for (int i = 0; i < 5; i++) {
#pragma GCC unroll 5
    for (int j = 0; j < 5; j++) {
      lo[i + j] = _mm256_madd52lo_epu64(lo[i + j], x->v[i], y->v[j]);
      hi[i + j + 1] = _mm256_madd52hi_epu64(hi[i + j + 1], x->v[i], y->v[j]);
    }
  }
  __m256i t[10];
  for (int k = 0; k < 10; k++) t[k] = V_ADD(lo[k], V_ADD(hi[k], hi[k]));   /* < 5 * 2^52 + 10 * 2^52 < 2^56 */
  fe4_fold(r, t);
}

/* r = x * x: the 10 cross products are formed once and doubled */
static inline void fe4_sq(fe4* r, const fe4* x) {
  const __m256i z = _mm256_setzero_si256();
  __m256i dl[10] = {z, z, z, z, z, z, z, z, z, z}, dh[10] = {z, z, z, z, z, z, z, z, z, z};
  __m256i cl[10] = {z, z, z, z, z, z, z, z, z, z}, ch[10] = {z, z, z, z, z, z, z, z, z, z};
#pragma GCC unroll 5
  for (int i = 0; i < 5; i++) {
    dl[2 * i] = _mm256_madd52lo_epu64(dl[2 * i], x->v[i], x->v[i]);
    dh[2 * i + 1] = _mm256_madd52hi_epu64(dh[2 * i + 1], x->v[i], x->v[i]);
#pragma GCC unroll 5
    for (int j = i + 1; j < 5; j++) {
      cl[i + j] = _mm256_madd52lo_epu64(cl[i + j], x->v[i], x->v[j]);
      ch[i + j + 1] = _mm256_madd52hi_epu64(ch[i + j + 1], x->v[i], x->v[j]);
    }
  }
  __m256i t[10];
  for (int k = 0; k < 10; k++) {   /* diag + 2 cross + 2 (diag_hi + 2 cross_hi): < 2^52 (1 + 4 + 2 + 8) < 2^56 */
    __m256i lo = V_ADD(dl[k], V_ADD(cl[k], cl[k])), h = V_ADD(dh[k], V_ADD(ch[k], ch[k]));
    t[k] = V_ADD(lo, V_ADD(h, h));
  }
  fe4_fold(r, t);
}

/* limbs of 2p: what a subtrahend with limbs below 2^52 - 38 may be taken from without a borrow */
static inline void fe4_bias(__m256i* b) {
  b[0] = V_SET1(0xfffffffffffdaULL);
  b[1] = b[2] = b[3] = b[4] = V_SET1(0xffffffffffffeULL);
}

static inline void fe4_pack(fe4* r, const fe* a, const fe* b, const fe* c, const fe* d) {
  for (int k = 0; k < 5; k++) r->v[k] = _mm256_set_epi64x((long long)d->v[k], (long long)c->v[k], (long long)b->v[k], (long long)a->v[k]);
}
static inline void fe4_unpack(fe* a, fe* b, fe* c, fe* d, const fe4* x) {
  u64 t[4] __attribute__((aligned(32)));
  for (int k = 0; k < 5; k++) {
    _mm256_store_si256((__m256i*)t, x->v[k]);
    a->v[k] = t[0]; b->v[k] = t[1]; c->v[k] = t[2]; d->v[k] = t[3];
  }
}

/* ---- points: lanes (X, Y, Z, T) of an extended point; lanes (Y-X, Y+X, 2Z, 2dT) of a cached addend ---- */
typedef fe4 ext4;
typedef fe4 cached4;

static inline void ext4_identity(ext4* r) {
  const __m256i one = _mm256_set_epi64x(0, 1, 1, 0), z = _mm256_setzero_si256();
  r->v[0] = one; r->v[1] = r->v[2] = r->v[3] = r->v[4] = z;
}
static void ext4_from_ge(ext4* r, const ge* p) { fe4_pack(r, &p->X, &p->Y, &p->Z, &p->T); fe4_carry(r); }
static void ext4_to_ge(ge* p, const ext4* x) { fe4_unpack(&p->X, &p->Y, &p->Z, &p->T, x); }
static void cached4_from_ge(cached4* r, const ge* p) {
  pniels n; fe z2;
  ge_to_pniels(&n, p);
  fe_add(&z2, &n.Z, &n.Z);
  fe4_pack(r, &n.YmX, &n.YpX, &z2, &n.T2d);
  fe4_carry(r);
}
/* -Q: swap (Y-X, Y+X), negate 2dT */
static inline void cached4_neg(cached4* r, const cached4* q) {
  __m256i bias[5];
  fe4_bias(bias);
  for (int k = 0; k < 5; k++) {
    __m256i s = V_PERM(q->v[k], 1, 0, 2, 3);
    r->v[k] = V_BLEND(s, V_SUB(bias[k], s), LANE3);   /* 2p - 2dT < 2^52: a valid multiplicand as it is */
  }
}

/* r = p + q.  Two vector multiplications:
 *   (Y1-X1, Y1+X1, Z1, T1) * (Y2-X2, Y2+X2, 2 Z2, 2d T2) = (MM, PP, ZZ2, TT2d)
 *   E = PP-MM, H = PP+MM, G = ZZ2+TT2d, F = ZZ2-TT2d;  (E, H, G, E) * (F, G, F, H) = (X3, Y3, Z3, T3)              */
static inline void ext4_add_cached(ext4* r, const ext4* p, const cached4* q) {
  __m256i bias[5];
  fe4_bias(bias);
  fe4 tmp, s, l, rr;
  for (int k = 0; k < 5; k++) {
    __m256i t = V_PERM(p->v[k], 1, 0, 2, 3);                    /* (Y, X, Z, T)        */
    __m256i sum = V_ADD(p->v[k], t);                            /* (X+Y, X+Y, ..)      */
    __m256i dif = V_SUB(V_ADD(t, bias[k]), p->v[k]);            /* (Y-X, ..)           */
    tmp.v[k] = V_BLEND(V_BLEND(p->v[k], sum, LANE1), dif, LANE0);
  }
  fe4_carry(&tmp);
  fe4_mul(&s, &tmp, q);                                         /* (MM, PP, ZZ2, TT2d) */
  fe4 sum, dif;
  for (int k = 0; k < 5; k++) {
    __m256i u = V_PERM(s.v[k], 1, 0, 3, 2);                     /* (PP, MM, TT2d, ZZ2) */
    sum.v[k] = V_ADD(s.v[k], u);                                /* (H, H, G, G)        */
    dif.v[k] = V_SUB(V_ADD(u, bias[k]), s.v[k]);                /* (E, -E, -F, F)      */
  }
  fe4_carry(&sum);
  fe4_carry(&dif);
  for (int k = 0; k < 5; k++) {
    l.v[k] = V_BLEND(V_PERM(dif.v[k], 0, 0, 0, 0), sum.v[k], LANE1 | LANE2);                           /* (E, H, G, E) */
    rr.v[k] = V_BLEND(V_PERM(dif.v[k], 3, 3, 3, 3), V_PERM(sum.v[k], 0, 2, 0, 0), LANE1 | LANE3);      /* (F, G, F, H) */
  }
  fe4_mul(r, &l, &rr);
}

/* r = 2 p.  One vector squaring and one vector multiplication:
 *   (X, Y, Z, X+Y)^2 = (XX, YY, ZZ, XPY2);  cY = XX+YY, cZ = YY-XX, cX = XPY2-cY, cT = 2ZZ-cZ;
 *   (cX, cY, cZ, cX) * (cT, cZ, cT, cY) = (X3, Y3, Z3, T3)                                                          */
static inline void ext4_double(ext4* r, const ext4* p) {
  __m256i bias[5];
  fe4_bias(bias);
  fe4 a, sq, sum, dif, w, l, rr;
  for (int k = 0; k < 5; k++) {
    __m256i t = V_PERM(p->v[k], 1, 0, 2, 3);
    __m256i s = V_ADD(p->v[k], t);                              /* lanes 0, 1 = X+Y    */
    a.v[k] = V_BLEND(p->v[k], V_PERM(s, 0, 0, 0, 0), LANE3);    /* (X, Y, Z, X+Y)      */
  }
  fe4_carry(&a);
  fe4_sq(&sq, &a);
  for (int k = 0; k < 5; k++) {
    __m256i u = V_PERM(sq.v[k], 1, 0, 2, 3);                    /* (YY, XX, ZZ, XPY2)  */
    sum.v[k] = V_ADD(sq.v[k], u);                               /* (cY, cY, 2ZZ, ..)   */
    dif.v[k] = V_SUB(V_ADD(sq.v[k], bias[k]), u);               /* (-cZ, cZ, 0, 0)     */
  }
  fe4_carry(&sum);
  fe4_carry(&dif);
  for (int k = 0; k < 5; k++) {
    __m256i v1 = V_BLEND(sq.v[k], sum.v[k], LANE2);                                                    /* (.., .., 2ZZ, XPY2) */
    __m256i v2 = V_BLEND(V_PERM(dif.v[k], 1, 1, 1, 1), V_PERM(sum.v[k], 0, 0, 0, 0), LANE3);           /* (.., .., cZ, cY)    */
    w.v[k] = V_SUB(V_ADD(v1, bias[k]), v2);                                                            /* (.., .., cT, cX)    */
  }
  fe4_carry(&w);
  for (int k = 0; k < 5; k++) {
    __m256i cy = V_PERM(sum.v[k], 0, 0, 0, 0), cz = V_PERM(dif.v[k], 1, 1, 1, 1);
    l.v[k] = V_BLEND(V_BLEND(V_PERM(w.v[k], 3, 3, 3, 3), cy, LANE1), cz, LANE2);                        /* (cX, cY, cZ, cX) */
    rr.v[k] = V_BLEND(V_BLEND(V_PERM(w.v[k], 2, 2, 2, 2), cz, LANE1), cy, LANE3);                       /* (cT, cZ, cT, cY) */
  }
  fe4_mul(r, &l, &rr);
}

/* ---- scalar_mul/pippenger.rs and straus.rs (vartime) over the vector point types ---- */
static void pippenger_simd(ge* out, const uint8_t* scalars, const ge* pts, size_t n) {
  int w = n < 500 ? 6 : (n < 800 ? 7 : 8);
  int max_digit = 1 << w, digits_count = radix_2w_digits(w), buckets_count = max_digit / 2;
  int8_t* dig = (int8_t*)malloc(n * 43);
  cached4* pc = (cached4*)aligned_alloc(32, (n ? n : 1) * sizeof(cached4));
  ext4* buckets = (ext4*)aligned_alloc(32, (size_t)buckets_count * sizeof(ext4));
  for (size_t i = 0; i < n; i++) { to_radix_2w(dig + 43 * i, scalars + 32 * i, w); cached4_from_ge(&pc[i], &pts[i]); }
  ge total; ge_identity(&total);
  for (int idx = digits_count - 1; idx >= 0; idx--) {
    for (int b = 0; b < buckets_count; b++) ext4_identity(&buckets[b]);
    for (size_t i = 0; i < n; i++) {
      int d = dig[43 * i + idx];
      if (d > 0) ext4_add_cached(&buckets[d - 1], &buckets[d - 1], &pc[i]);
      else if (d < 0) { cached4 m; cached4_neg(&m, &pc[i]); ext4_add_cached(&buckets[-d - 1], &buckets[-d - 1], &m); }
    }
    /* running sums of the buckets, highest first (pippenger.rs) */
    ext4 run = buckets[buckets_count - 1], acc = buckets[buckets_count - 1];
    for (int b = buckets_count - 2; b >= 0; b--) {
      ge t; cached4 c;
      ext4_to_ge(&t, &buckets[b]); cached4_from_ge(&c, &t); ext4_add_cached(&run, &run, &c);
      ext4_to_ge(&t, &run); cached4_from_ge(&c, &t); ext4_add_cached(&acc, &acc, &c);
    }
    ge accg; ext4_to_ge(&accg, &acc);
    if (idx != digits_count - 1) ge_mul_pow2(&total, &total, w);
    ge_add(&total, &total, &accg);
  }
  *out = total;
  free(dig); free(pc); free(buckets);
}

static void straus_vt_simd(ge* out, const uint8_t* scalars, const ge* pts, size_t n) {
  cached4* tab = (cached4*)aligned_alloc(32, (n ? n : 1) * 8 * sizeof(cached4));
  int8_t* nafs = (int8_t*)malloc((n ? n : 1) * 256);
  for (size_t i = 0; i < n; i++) {   /* odd multiples 1P, 3P, .., 15P */
    ge p2; ext4 m, d; cached4 p2c;
    ext4_from_ge(&m, &pts[i]);
    ext4_double(&d, &m); ext4_to_ge(&p2, &d); cached4_from_ge(&p2c, &p2);
    cached4_from_ge(&tab[8 * i], &pts[i]);
    for (int k = 1; k < 8; k++) { ge t; ext4_add_cached(&m, &m, &p2c); ext4_to_ge(&t, &m); cached4_from_ge(&tab[8 * i + k], &t); }
    naf5(nafs + 256 * i, scalars + 32 * i);
  }
  ext4 r; ext4_identity(&r);
  for (int i = 255; i >= 0; i--) {
    ext4_double(&r, &r);
    for (size_t k = 0; k < n; k++) {
      int d = nafs[256 * k + i];
      if (d > 0) ext4_add_cached(&r, &r, &tab[8 * k + d / 2]);
      else if (d < 0) { cached4 m; cached4_neg(&m, &tab[8 * k + (-d) / 2]); ext4_add_cached(&r, &r, &m); }
    }
  }
  ext4_to_ge(out, &r);
  free(tab); free(nafs);
}

static void msm_dispatch_simd(ge* out, const uint8_t* scalars, const ge* pts, size_t n) {
  if (n < 190) straus_vt_simd(out, scalars, pts, n); else pippenger_simd(out, scalars, pts, n);
}

/* test hook: lane-wise product / square / point ops on serial operands (tests pin the vector field against the serial one) */
int ref_simd_selftest_fe(const uint64_t* a20, const uint64_t* b20, uint64_t* mul20, uint64_t* sq20) {
  if (!ref_simd_available()) return -1;
  fe a[4], b[4], m[4], s[4]; fe4 x, y, r;
  for (int j = 0; j < 4; j++) { memcpy(a[j].v, a20 + 5 * j, 40); memcpy(b[j].v, b20 + 5 * j, 40); }
  fe4_pack(&x, &a[0], &a[1], &a[2], &a[3]); fe4_carry(&x);
  fe4_pack(&y, &b[0], &b[1], &b[2], &b[3]); fe4_carry(&y);
  fe4_mul(&r, &x, &y); fe4_unpack(&m[0], &m[1], &m[2], &m[3], &r);
  fe4_sq(&r, &x); fe4_unpack(&s[0], &s[1], &s[2], &s[3], &r);
  for (int j = 0; j < 4; j++) { memcpy(mul20 + 5 * j, m[j].v, 40); memcpy(sq20 + 5 * j, s[j].v, 40); }
  return 0;
}
#pragma GCC pop_options
#define REF_HAVE_SIMD 1
#else
int ref_simd_available(void) { return 0; }
#define REF_HAVE_SIMD 0
#endif
