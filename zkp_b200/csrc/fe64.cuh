// GF(2^255-19) on the FP64 pipe: the second arithmetic engine of the decompression kernel (sm_100a).
//
// Why.  tools/ubench2.cu (profiles/r01_issue_model.md) shows that the integer field arithmetic of fe.cuh is bound by
// the FMA-heavy pipe: IMAD.WIDE.U32 costs ~4.2 issue cycles of an SM sub-partition, and fe_sq / fe_mul sit at 85-93 %
// of that bound.  The FP64 pipe (DFMA, 2 cycles per warp instruction) is idle and runs beside the integer pipes almost
// for free.  Ristretto decompression (curve25519-dalek `CompressedRistretto::decompress` [ext], reached from
// /root/reference/src/toolbox/batch_verifier.rs:226, verifier.rs:90,164) is 254 squarings per point, so some warps of
// k_decompress run the exponentiation a^((p-5)/8) with the arithmetic below while the others run fe.cuh.
//
// Representation.  Twelve limbs in radix 2^21.25 (sizes 22,21,21,21 repeating; 12 * 21.25 = 255, so 2^255 = 19 wraps
// exactly at limb 12).  Limb i is stored as a double holding q_i * 2^(o_i), q_i a signed integer, o_i = ceil(21.25 i):
// the weight lives in the exponent, so products a_i * a_j carry the right weight with no per-pair factor, and every
// quantity in column k is an integer multiple of 2^(o_k).  All operations are EXACT:
//   * balanced limbs |q_i| <= 2^(s_i - 1) + 2^12 give |a_i a_j| <= 2^(o_i + o_j + 42.01); a column sums at most 12
//     products, the wrapped ones times 19 and cross terms of a squaring times 2: |c_k| < 12 * 38 * 2^42.01 * 2^(o_k)
//     < 2^(o_k + 50.9), below the 2^53 integer range of a double, so every DFMA result is exact;
//   * a carry rounds column k to the nearest multiple of 2^(o_(k+1)) by adding and subtracting 1.5 * 2^(52 + o_(k+1)).
// Results are therefore the same field elements fe.cuh computes; tests/test_host_emul.py checks both against the
// big-int oracle (this header also compiles for the host with -DZKP_HOST_EMUL).
#pragma once
#include <stdint.h>
#include "fe.cuh"

#if !ZKP_DEVICE_ASM
#include <math.h>
#endif

namespace zkp {

struct fe64 { double v[12]; };

// exact IEEE operations, immune to contraction / reassociation by the compiler
#if ZKP_DEVICE_ASM
ZKP_DEV double d_add(double a, double b) { return __dadd_rn(a, b); }
ZKP_DEV double d_sub(double a, double b) { return __dsub_rn(a, b); }
ZKP_DEV double d_mul(double a, double b) { return __dmul_rn(a, b); }
ZKP_DEV double d_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
#else
ZKP_DEV double d_add(double a, double b) { volatile double r = a + b; return r; }
ZKP_DEV double d_sub(double a, double b) { volatile double r = a - b; return r; }
ZKP_DEV double d_mul(double a, double b) { volatile double r = a * b; return r; }
ZKP_DEV double d_fma(double a, double b, double c) { return fma(a, b, c); }
#endif

// 2^e as a double, e a compile-time constant in the normal range
ZKP_DEV double d_pow2(int e) {
#if ZKP_DEVICE_ASM
  return __longlong_as_double((long long)(1023 + e) << 52);
#else
  return ldexp(1.0, e);
#endif
}

#define ZKP_FE64_OFF(i) ((85 * (i) + 3) / 4)   // o_i = ceil(21.25 i): 0,22,43,64,85,107,128,149,170,192,213,234,255

// carry out of column k into column k+1 (k = 11 wraps into column 0 with 2^255 = 19), rounding to nearest:
// afterwards |c_k| <= 2^(o_(k+1) - 1)
template <int K>
ZKP_DEV void fe64_carry(double* c) {
  const double M = 1.5 * d_pow2(52 + ZKP_FE64_OFF(K + 1));
  const double t = d_add(c[K], M);
  const double hi = d_sub(t, M);
  c[K] = d_sub(c[K], hi);
  if (K == 11) c[0] = d_fma(hi, 19.0 * d_pow2(-255), c[0]);
  else c[K + 1] = d_add(c[K + 1], hi);
}

// two interleaved carry chains 0->..->6 and 6->..->11->0, then 0->1 and 6->7 once more: 14 carries, critical path 8
ZKP_DEV void fe64_carry_all(double* c) {
  fe64_carry<0>(c); fe64_carry<6>(c);
  fe64_carry<1>(c); fe64_carry<7>(c);
  fe64_carry<2>(c); fe64_carry<8>(c);
  fe64_carry<3>(c); fe64_carry<9>(c);
  fe64_carry<4>(c); fe64_carry<10>(c);
  fe64_carry<5>(c); fe64_carry<11>(c);
  fe64_carry<0>(c); fe64_carry<6>(c);
}

// r = a^2: 78 products, 11 doublings, 11 wrap folds, 14 carries = 156 FP64 instructions
ZKP_DEV void fe64_sq(fe64& r, const fe64& a) {
  const double* A = a.v;
  double A2[12], c[12];
#pragma unroll
  for (int i = 0; i < 11; i++) A2[i] = d_add(A[i], A[i]);
  const double W19 = 19.0 * d_pow2(-255);
#pragma unroll
  for (int k = 0; k < 12; k++) {
    // direct terms i + j = k, i <= j
    double x = 0.0;
    bool xs = false;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      const int j = k - i;
      if (j < i || j > 11) continue;
      const double u = (i == j) ? A[i] : A2[i];
      x = xs ? d_fma(u, A[j], x) : d_mul(u, A[j]);
      xs = true;
    }
    // wrapped terms i + j = k + 12, i <= j
    double y = 0.0;
    bool ys = false;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      const int j = k + 12 - i;
      if (j < i || j > 11) continue;
      const double u = (i == j) ? A[i] : A2[i];
      y = ys ? d_fma(u, A[j], y) : d_mul(u, A[j]);
      ys = true;
    }
    c[k] = ys ? d_fma(y, W19, x) : x;
  }
  fe64_carry_all(c);
#pragma unroll
  for (int k = 0; k < 12; k++) r.v[k] = c[k];
}

// r = a * b: 144 products, 11 wrap folds, 14 carries = 169 FP64 instructions
ZKP_DEV void fe64_mul(fe64& r, const fe64& a, const fe64& b) {
  const double* A = a.v;
  const double* B = b.v;
  double c[12];
  const double W19 = 19.0 * d_pow2(-255);
#pragma unroll
  for (int k = 0; k < 12; k++) {
    double x = 0.0;
    bool xs = false;
#pragma unroll
    for (int i = 0; i <= k; i++) {
      x = xs ? d_fma(A[i], B[k - i], x) : d_mul(A[i], B[k - i]);
      xs = true;
    }
    double y = 0.0;
    bool ys = false;
#pragma unroll
    for (int i = k + 1; i < 12; i++) {
      y = ys ? d_fma(A[i], B[k + 12 - i], y) : d_mul(A[i], B[k + 12 - i]);
      ys = true;
    }
    c[k] = ys ? d_fma(y, W19, x) : x;
  }
  fe64_carry_all(c);
#pragma unroll
  for (int k = 0; k < 12; k++) r.v[k] = c[k];
}

// r = a^(2^n)
ZKP_DEV void fe64_sqn(fe64& r, const fe64& a, int n) {
  fe64 t = a;
#if ZKP_DEVICE_ASM
#pragma unroll 1
#endif
  for (int i = 0; i < n; i++) fe64_sq(t, t);
  r = t;
}

// ---- conversions -------------------------------------------------------------------------------------------------
// 8 x 32-bit words (any value < 2^256) -> balanced weighted limbs
ZKP_DEV void fe64_from_fe(fe64& r, const fe& a) {
  double c[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const int o = ZKP_FE64_OFF(i);
    const int bits = (i == 11) ? 22 : (ZKP_FE64_OFF(i + 1) - o);   // limb 11 also takes bit 255
    const int w = o >> 5, sh = o & 31;
    uint32_t q = a.v[w] >> sh;
    if (sh + bits > 32) q |= a.v[w + 1] << (32 - sh);
    q &= (1u << bits) - 1u;
    c[i] = d_mul((double)q, d_pow2(o));
  }
  fe64_carry_all(c);
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = c[i];
}

// floor(x / 2^e) * 2^e for 0 <= x < 2^(52 + e)
ZKP_DEV double d_floor_to(double x, int e) {
#if ZKP_DEVICE_ASM
  const double M = d_pow2(52 + e);
  return __dsub_rn(__dadd_rd(x, M), M);
#else
  return ldexp(floor(ldexp(x, -e)), e);
#endif
}

// balanced weighted limbs -> 8 x 32-bit words, value in [0, 2^256) (weakly reduced like every fe)
ZKP_DEV void fe64_to_fe(fe& r, const fe64& a) {
  double c[12];
  // add 8p limb-wise (p = 2^255 - 19: limb 0 = 2^22 - 19, the others all ones) so that every limb is positive
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const int o = ZKP_FE64_OFF(i), s = ZKP_FE64_OFF(i + 1) - o;
    const double bias = 8.0 * ((double)(1u << s) - (i == 0 ? 19.0 : 1.0));
    c[i] = d_fma(bias, d_pow2(o), a.v[i]);
  }
  // floor carries: limb i in [0, 2^(s_i)), top carry T (a small non-negative integer) has weight 2^255
  uint32_t q[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const int o = ZKP_FE64_OFF(i), o1 = ZKP_FE64_OFF(i + 1);
    const double hi = d_floor_to(c[i], o1);
    const double lo = d_sub(c[i], hi);
    q[i] = (uint32_t)(long long)d_mul(lo, d_pow2(-o));
    if (i < 11) c[i + 1] = d_add(c[i + 1], hi);
    else c[0] = d_mul(hi, d_pow2(-255));   // T
  }
  const uint32_t T = (uint32_t)(long long)c[0];
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const int o = ZKP_FE64_OFF(i), s = ZKP_FE64_OFF(i + 1) - o;
    const int k = o >> 5, sh = o & 31;
    w[k] |= q[i] << sh;
    if (sh + s > 32) w[k + 1] |= q[i] >> (32 - sh);
  }
  // + 19 T (T < 2^5): the packed value is < 2^255, so the sum stays below 2^256
  fe_tail_add<false>(w, 19u * T);
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = w[i];
}

// a^((p-5)/8) = a^(2^252 - 3) on the FP64 pipe: the same addition chain as fe_pow22523 (dalek field.rs pow22501 /
// pow_p58 [ext]); in and out in the integer representation
ZKP_DEV void fe_pow22523_fp64(fe& r, const fe& a_int) {
  fe64 a, t0, t1, t2, t3, t5, t7, t9, t13, t15, t;
  fe64_from_fe(a, a_int);
  fe64_sq(t0, a);            // 2
  fe64_sqn(t1, t0, 2);       // 8
  fe64_mul(t2, a, t1);       // 9
  fe64_mul(t3, t0, t2);      // 11
  fe64_sq(t, t3);            // 22
  fe64_mul(t5, t2, t);       // 2^5 - 1
  fe64_sqn(t, t5, 5);
  fe64_mul(t7, t, t5);       // 2^10 - 1
  fe64_sqn(t, t7, 10);
  fe64_mul(t9, t, t7);       // 2^20 - 1
  fe64_sqn(t, t9, 20);
  fe64_mul(t, t, t9);        // 2^40 - 1
  fe64_sqn(t, t, 10);
  fe64_mul(t13, t, t7);      // 2^50 - 1
  fe64_sqn(t, t13, 50);
  fe64_mul(t15, t, t13);     // 2^100 - 1
  fe64_sqn(t, t15, 100);
  fe64_mul(t, t, t15);       // 2^200 - 1
  fe64_sqn(t, t, 50);
  fe64_mul(t, t, t13);       // 2^250 - 1
  fe64_sqn(t, t, 2);
  fe64_mul(t, t, a);         // 2^252 - 3
  fe64_to_fe(r, t);
}

}  // namespace zkp
