// GF(2^255-19) arithmetic for sm_100a, entirely in registers.
//
// Replaces curve25519-dalek 2.x `backend/serial/u64/field.rs` (FieldElement51) [ext] -- SURVEY.md section 2.2
// row E6 -- underneath every hot-path call of the reference (decompress: toolbox/verifier.rs:90,164,
// toolbox/batch_verifier.rs:226; compress: toolbox/mod.rs:180,204; MSM: prover.rs:94, verifier.rs:97,162,
// batch_verifier.rs:219).
//
// Representation.  The radix-2^51 five-limb form (dalek's, and the one north_star names) stays the
// *interface* layout for limb-form points (include/zkp_b200.h).  Inside the kernels an element is eight
// saturated 32-bit limbs, any value in [0, 2^256) standing for its residue mod p: Blackwell's integer
// multiplier is a 32x32->64 IMAD.WIDE on the FMA pipe, and ptxas fuses each `mad.lo.cc / madc.hi.cc` pair
// below into ONE `IMAD.WIDE.U32.X` with predicate carry-in/out, so a multiply costs 64+8 wide IMADs versus
// ~100 for 5x51 (each 64x64->128 product = 4 IMAD.WIDE) or 10x25.5 limbs.  `bench_fe` measures all of them
// on the device ("field multiplier ablation").  Results are bit-exact whatever the internal radix
// because every output is canonicalised in fe_tobytes.
//
// Every carry chain is one self-contained asm block (no condition-code state crosses statements).
// With -DZKP_HOST_EMUL the same source compiles for the host with the asm blocks replaced by 64-bit C
// (tests/host_emul): the kernel *logic* is checked against the oracle on CPU; that build is test
// infrastructure, never a fallback.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(ZKP_HOST_EMUL)
#define ZKP_DEV __device__ __forceinline__
#define ZKP_DEVICE_ASM 1
#else
#define ZKP_DEV static inline
#define ZKP_DEVICE_ASM 0
#endif

namespace zkp {

struct fe { uint32_t v[8]; };

// ------------------------------------------------------------------------------------------------------
// carry-chain primitives
// ------------------------------------------------------------------------------------------------------

// acc[0..2n-1] += (a0,a1,..)*b laid out as non-overlapping 64-bit pairs; returns the carry out (0/1).
ZKP_DEV uint32_t mad4(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
#if ZKP_DEVICE_ASM
  uint32_t c;
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32 %8, 0, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
        "+r"(acc[7]), "=r"(c)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
  return c;
#else
  const uint32_t a[4] = {a0, a1, a2, a3};
  uint64_t c = 0;
  for (int k = 0; k < 4; k++) {
    uint64_t p = (uint64_t)a[k] * b;
    uint64_t lo = (uint64_t)acc[2 * k] + (uint32_t)p + c;
    acc[2 * k] = (uint32_t)lo;
    uint64_t hi = (uint64_t)acc[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
    acc[2 * k + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
  return (uint32_t)c;
#endif
}

ZKP_DEV uint32_t mad3(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b) {
#if ZKP_DEVICE_ASM
  uint32_t c;
  asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
      "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
      "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
      "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
      "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
      "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
      "addc.u32 %6, 0, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "=r"(c)
      : "r"(a0), "r"(a1), "r"(a2), "r"(b));
  return c;
#else
  const uint32_t a[3] = {a0, a1, a2};
  uint64_t c = 0;
  for (int k = 0; k < 3; k++) {
    uint64_t p = (uint64_t)a[k] * b;
    uint64_t lo = (uint64_t)acc[2 * k] + (uint32_t)p + c;
    acc[2 * k] = (uint32_t)lo;
    uint64_t hi = (uint64_t)acc[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
    acc[2 * k + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
  return (uint32_t)c;
#endif
}

ZKP_DEV uint32_t mad2(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t b) {
#if ZKP_DEVICE_ASM
  uint32_t c;
  asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
      "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
      "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
      "addc.u32 %4, 0, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "=r"(c)
      : "r"(a0), "r"(a1), "r"(b));
  return c;
#else
  const uint32_t a[2] = {a0, a1};
  uint64_t c = 0;
  for (int k = 0; k < 2; k++) {
    uint64_t p = (uint64_t)a[k] * b;
    uint64_t lo = (uint64_t)acc[2 * k] + (uint32_t)p + c;
    acc[2 * k] = (uint32_t)lo;
    uint64_t hi = (uint64_t)acc[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
    acc[2 * k + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
  return (uint32_t)c;
#endif
}

ZKP_DEV uint32_t mad1(uint32_t* acc, uint32_t a0, uint32_t b) {
#if ZKP_DEVICE_ASM
  uint32_t c;
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, 0, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "=r"(c)
      : "r"(a0), "r"(b));
  return c;
#else
  uint64_t p = (uint64_t)a0 * b;
  uint64_t lo = (uint64_t)acc[0] + (uint32_t)p;
  acc[0] = (uint32_t)lo;
  uint64_t hi = (uint64_t)acc[1] + (uint32_t)(p >> 32) + (lo >> 32);
  acc[1] = (uint32_t)hi;
  return (uint32_t)(hi >> 32);
#endif
}

// ---- chains whose carry never leaves the chain (round 2) ----------------------------------------------------------------
// A carry captured into a fresh limb (`c = addc 0, 0`) becomes, one row later, the low word of a 64-bit addend whose high
// word must be zeroed: ptxas emits SEL + `IMAD.MOV.U32 Rx, RZ, RZ, RZ`, and the zeroing lands on the FMA-heavy pipe that
// the wide multiplies saturate (profiles/r02_pipe_counters.md: 6 % of its cycles in a squaring).  The chains below end
// either in a FRESH pair computed in the chain itself -- (x * y) + carry, which cannot overflow 64 bits -- or in a two-limb
// ripple of the carry into the pair above, which already holds one product plus at most one earlier carry.
#if ZKP_DEVICE_ASM
#define ZKP_MADPAIR(lo, hi, a, b) "madc.lo.cc.u32 " lo ", " a ", " b ", " lo ";\n\tmadc.hi.cc.u32 " hi ", " a ", " b ", " hi ";\n\t"
#endif
// acc[0..2N-1] += (a0..a_{N-1}) * b as pairs; then acc[2N], acc[2N+1] = x * y + carry   (fresh pair, written not read)
ZKP_DEV void mad1x1(uint32_t* acc, uint32_t a0, uint32_t b, uint32_t x, uint32_t y) {
#if ZKP_DEVICE_ASM
  asm("mad.lo.cc.u32 %0, %4, %5, %0;\n\tmadc.hi.cc.u32 %1, %4, %5, %1;\n\t"
      "madc.lo.cc.u32 %2, %6, %7, 0;\n\tmadc.hi.u32 %3, %6, %7, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "=r"(acc[2]), "=r"(acc[3])
      : "r"(a0), "r"(b), "r"(x), "r"(y));
#else
  acc[2] = 0; acc[3] = 0;
  uint32_t c = mad1(acc, a0, b);
  uint64_t p = (uint64_t)x * y + c;
  acc[2] = (uint32_t)p; acc[3] = (uint32_t)(p >> 32);
#endif
}
ZKP_DEV void mad2x1(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t b, uint32_t x, uint32_t y) {
#if ZKP_DEVICE_ASM
  asm("mad.lo.cc.u32 %0, %6, %8, %0;\n\tmadc.hi.cc.u32 %1, %6, %8, %1;\n\t"
      "madc.lo.cc.u32 %2, %7, %8, %2;\n\tmadc.hi.cc.u32 %3, %7, %8, %3;\n\t"
      "madc.lo.cc.u32 %4, %9, %10, 0;\n\tmadc.hi.u32 %5, %9, %10, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "=r"(acc[4]), "=r"(acc[5])
      : "r"(a0), "r"(a1), "r"(b), "r"(x), "r"(y));
#else
  uint32_t c = mad2(acc, a0, a1, b);
  uint64_t p = (uint64_t)x * y + c;
  acc[4] = (uint32_t)p; acc[5] = (uint32_t)(p >> 32);
#endif
}
ZKP_DEV void mad3x1(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b, uint32_t x, uint32_t y) {
#if ZKP_DEVICE_ASM
  asm("mad.lo.cc.u32 %0, %8, %11, %0;\n\tmadc.hi.cc.u32 %1, %8, %11, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %11, %2;\n\tmadc.hi.cc.u32 %3, %9, %11, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %11, %4;\n\tmadc.hi.cc.u32 %5, %10, %11, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, 0;\n\tmadc.hi.u32 %7, %12, %13, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "=r"(acc[6]), "=r"(acc[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(b), "r"(x), "r"(y));
#else
  uint32_t c = mad3(acc, a0, a1, a2, b);
  uint64_t p = (uint64_t)x * y + c;
  acc[6] = (uint32_t)p; acc[7] = (uint32_t)(p >> 32);
#endif
}
ZKP_DEV void mad4x1(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b, uint32_t x, uint32_t y) {
#if ZKP_DEVICE_ASM
  asm("mad.lo.cc.u32 %0, %10, %14, %0;\n\tmadc.hi.cc.u32 %1, %10, %14, %1;\n\t"
      "madc.lo.cc.u32 %2, %11, %14, %2;\n\tmadc.hi.cc.u32 %3, %11, %14, %3;\n\t"
      "madc.lo.cc.u32 %4, %12, %14, %4;\n\tmadc.hi.cc.u32 %5, %12, %14, %5;\n\t"
      "madc.lo.cc.u32 %6, %13, %14, %6;\n\tmadc.hi.cc.u32 %7, %13, %14, %7;\n\t"
      "madc.lo.cc.u32 %8, %15, %16, 0;\n\tmadc.hi.u32 %9, %15, %16, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
        "=r"(acc[8]), "=r"(acc[9])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b), "r"(x), "r"(y));
#else
  uint32_t c = mad4(acc, a0, a1, a2, a3, b);
  uint64_t p = (uint64_t)x * y + c;
  acc[8] = (uint32_t)p; acc[9] = (uint32_t)(p >> 32);
#endif
}
// acc[0..2N-1] += (a0..a_{N-1}) * b as pairs; then the carry ripples into acc[2N], acc[2N+1] (no carry out: see above)
ZKP_DEV void mad1r(uint32_t* acc, uint32_t a0, uint32_t b) {
#if ZKP_DEVICE_ASM
  asm("mad.lo.cc.u32 %0, %4, %5, %0;\n\tmadc.hi.cc.u32 %1, %4, %5, %1;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\taddc.u32 %3, %3, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3])
      : "r"(a0), "r"(b));
#else
  uint32_t c = mad1(acc, a0, b);
  uint64_t s = (uint64_t)acc[2] + c;
  acc[2] = (uint32_t)s; acc[3] += (uint32_t)(s >> 32);
#endif
}
ZKP_DEV void mad2r(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t b) {
#if ZKP_DEVICE_ASM
  asm("mad.lo.cc.u32 %0, %6, %8, %0;\n\tmadc.hi.cc.u32 %1, %6, %8, %1;\n\t"
      "madc.lo.cc.u32 %2, %7, %8, %2;\n\tmadc.hi.cc.u32 %3, %7, %8, %3;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\taddc.u32 %5, %5, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5])
      : "r"(a0), "r"(a1), "r"(b));
#else
  uint32_t c = mad2(acc, a0, a1, b);
  uint64_t s = (uint64_t)acc[4] + c;
  acc[4] = (uint32_t)s; acc[5] += (uint32_t)(s >> 32);
#endif
}
ZKP_DEV void mad3r(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b) {
#if ZKP_DEVICE_ASM
  asm("mad.lo.cc.u32 %0, %8, %11, %0;\n\tmadc.hi.cc.u32 %1, %8, %11, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %11, %2;\n\tmadc.hi.cc.u32 %3, %9, %11, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %11, %4;\n\tmadc.hi.cc.u32 %5, %10, %11, %5;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\taddc.u32 %7, %7, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(b));
#else
  uint32_t c = mad3(acc, a0, a1, a2, b);
  uint64_t s = (uint64_t)acc[6] + c;
  acc[6] = (uint32_t)s; acc[7] += (uint32_t)(s >> 32);
#endif
}

// (lo,hi) = a*b  -- a single IMAD.WIDE.U32
ZKP_DEV void mulw(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t p = (uint64_t)a * b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
}

// acc[0..7] += (a0^2, a1^2, a2^2, a3^2) as pairs, + cin; returns carry out
ZKP_DEV uint32_t sqr4c(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t cin) {
#if ZKP_DEVICE_ASM
  uint32_t c;
  asm("{\n\t.reg .u32 t;\n\t"
      "add.cc.u32 t, %13, 0xffffffff;\n\t"
      "madc.lo.cc.u32 %0, %9, %9, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %9, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %10, %2;\n\t"
      "madc.hi.cc.u32 %3, %10, %10, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %11, %4;\n\t"
      "madc.hi.cc.u32 %5, %11, %11, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %12, %6;\n\t"
      "madc.hi.cc.u32 %7, %12, %12, %7;\n\t"
      "addc.u32 %8, 0, 0;\n\t}"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
        "+r"(acc[7]), "=r"(c)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(cin));
  return c;
#else
  const uint32_t a[4] = {a0, a1, a2, a3};
  uint64_t c = cin;
  for (int k = 0; k < 4; k++) {
    uint64_t p = (uint64_t)a[k] * a[k];
    uint64_t lo = (uint64_t)acc[2 * k] + (uint32_t)p + c;
    acc[2 * k] = (uint32_t)lo;
    uint64_t hi = (uint64_t)acc[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
    acc[2 * k + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
  return (uint32_t)c;
#endif
}

// r[0..7] = a[0..7] + b[0..7]; returns carry out
ZKP_DEV uint32_t add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if ZKP_DEVICE_ASM
  uint32_t c;
  asm("add.cc.u32 %0, %9, %17;\n\t"
      "addc.cc.u32 %1, %10, %18;\n\t"
      "addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\t"
      "addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t"
      "addc.cc.u32 %7, %16, %24;\n\t"
      "addc.u32 %8, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  return c;
#else
  uint64_t c = 0;
  uint32_t t[8];
  for (int i = 0; i < 8; i++) {
    uint64_t s = (uint64_t)a[i] + b[i] + c;
    t[i] = (uint32_t)s;
    c = s >> 32;
  }
  for (int i = 0; i < 8; i++) r[i] = t[i];
  return (uint32_t)c;
#endif
}

// r[0..7] = a[0..7] + b[0..7] + cin; returns carry out
ZKP_DEV uint32_t add8c(uint32_t* r, const uint32_t* a, const uint32_t* b, uint32_t cin) {
#if ZKP_DEVICE_ASM
  uint32_t c;
  asm("{\n\t.reg .u32 t;\n\t"
      "add.cc.u32 t, %25, 0xffffffff;\n\t"
      "addc.cc.u32 %0, %9, %17;\n\t"
      "addc.cc.u32 %1, %10, %18;\n\t"
      "addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\t"
      "addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t"
      "addc.cc.u32 %7, %16, %24;\n\t"
      "addc.u32 %8, 0, 0;\n\t}"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(cin));
  return c;
#else
  uint64_t c = cin;
  uint32_t t[8];
  for (int i = 0; i < 8; i++) {
    uint64_t s = (uint64_t)a[i] + b[i] + c;
    t[i] = (uint32_t)s;
    c = s >> 32;
  }
  for (int i = 0; i < 8; i++) r[i] = t[i];
  return (uint32_t)c;
#endif
}

// r[0..7] = a[0..7] - b[0..7]; returns borrow out (0/1)
ZKP_DEV uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if ZKP_DEVICE_ASM
  uint32_t c;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  return c & 1u;
#else
  uint64_t br = 0;
  uint32_t t[8];
  for (int i = 0; i < 8; i++) {
    uint64_t s = (uint64_t)a[i] - b[i] - br;
    t[i] = (uint32_t)s;
    br = (s >> 32) & 1;
  }
  for (int i = 0; i < 8; i++) r[i] = t[i];
  return (uint32_t)br;
#endif
}

// r[0..7] += k (a small word), returns carry out.
ZKP_DEV uint32_t add_word8(uint32_t* r, uint32_t k) {
#if ZKP_DEVICE_ASM
  uint32_t c;
  asm("add.cc.u32 %0, %0, %9;\n\t"
      "addc.cc.u32 %1, %1, 0;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t"
      "addc.cc.u32 %3, %3, 0;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.cc.u32 %7, %7, 0;\n\t"
      "addc.u32 %8, 0, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "=r"(c)
      : "r"(k));
  return c;
#else
  uint64_t c = k;
  for (int i = 0; i < 8; i++) {
    uint64_t s = (uint64_t)r[i] + c;
    r[i] = (uint32_t)s;
    c = s >> 32;
  }
  return (uint32_t)c;
#endif
}

// r[0..7] -= k, returns borrow out.
ZKP_DEV uint32_t sub_word8(uint32_t* r, uint32_t k) {
#if ZKP_DEVICE_ASM
  uint32_t c;
  asm("sub.cc.u32 %0, %0, %9;\n\t"
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.cc.u32 %2, %2, 0;\n\t"
      "subc.cc.u32 %3, %3, 0;\n\t"
      "subc.cc.u32 %4, %4, 0;\n\t"
      "subc.cc.u32 %5, %5, 0;\n\t"
      "subc.cc.u32 %6, %6, 0;\n\t"
      "subc.cc.u32 %7, %7, 0;\n\t"
      "subc.u32 %8, 0, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "=r"(c)
      : "r"(k));
  return c & 1u;
#else
  uint64_t br = k;
  for (int i = 0; i < 8; i++) {
    uint64_t s = (uint64_t)r[i] - br;
    r[i] = (uint32_t)s;
    br = (s >> 32) & 1;
  }
  return (uint32_t)br;
#endif
}

// r[0..14] = a[0..14] + b[0..14] as ONE carry chain (the caller knows the sum fits: no carry out)
ZKP_DEV void add15(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if ZKP_DEVICE_ASM
  asm("add.cc.u32 %0, %15, %30;\n\t"
      "addc.cc.u32 %1, %16, %31;\n\t"
      "addc.cc.u32 %2, %17, %32;\n\t"
      "addc.cc.u32 %3, %18, %33;\n\t"
      "addc.cc.u32 %4, %19, %34;\n\t"
      "addc.cc.u32 %5, %20, %35;\n\t"
      "addc.cc.u32 %6, %21, %36;\n\t"
      "addc.cc.u32 %7, %22, %37;\n\t"
      "addc.cc.u32 %8, %23, %38;\n\t"
      "addc.cc.u32 %9, %24, %39;\n\t"
      "addc.cc.u32 %10, %25, %40;\n\t"
      "addc.cc.u32 %11, %26, %41;\n\t"
      "addc.cc.u32 %12, %27, %42;\n\t"
      "addc.cc.u32 %13, %28, %43;\n\t"
      "addc.u32 %14, %29, %44;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]),
        "r"(b[8]), "r"(b[9]), "r"(b[10]), "r"(b[11]), "r"(b[12]), "r"(b[13]), "r"(b[14]));
#else
  uint64_t c = 0;
  uint32_t t[15];
  for (int i = 0; i < 15; i++) {
    uint64_t s = (uint64_t)a[i] + b[i] + c;
    t[i] = (uint32_t)s;
    c = s >> 32;
  }
  for (int i = 0; i < 15; i++) r[i] = t[i];
#endif
}

// r[0..7] = a[0..7] + b[0..7], carry out dropped (the caller knows there is none)
ZKP_DEV void add8_nc(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if ZKP_DEVICE_ASM
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
  uint64_t c = 0;
  uint32_t t[8];
  for (int i = 0; i < 8; i++) {
    uint64_t s = (uint64_t)a[i] + b[i] + c;
    t[i] = (uint32_t)s;
    c = s >> 32;
  }
  for (int i = 0; i < 8; i++) r[i] = t[i];
#endif
}

// acc[0..15] += (a0^2, ..., a7^2) laid out as eight 64-bit pairs, ONE carry chain (the sum fits in 512 bits)
ZKP_DEV void sqr8(uint32_t* acc, const uint32_t* a) {
#if ZKP_DEVICE_ASM
  asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
      "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
      "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
      "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
      "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
      "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
      "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
      "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
      "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
      "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
      "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
      "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
      "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
      "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
      "madc.hi.u32 %15, %23, %23, %15;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
        "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11]), "+r"(acc[12]), "+r"(acc[13]),
        "+r"(acc[14]), "+r"(acc[15])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#else
  uint64_t c = 0;
  for (int k = 0; k < 8; k++) {
    uint64_t p = (uint64_t)a[k] * a[k];
    uint64_t lo = (uint64_t)acc[2 * k] + (uint32_t)p + c;
    acc[2 * k] = (uint32_t)lo;
    uint64_t hi = (uint64_t)acc[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
    acc[2 * k + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
#endif
}

// ------------------------------------------------------------------------------------------------------
// tails: r[0..7] (+|-)= k with the 2^256 wrap folded back as 38.  k is small (< 2^12), so a carry out of limb 0
// has probability ~ k / 2^32.  The constant-time form always ripples through the eight limbs; the variable-time
// form (VT, public data only: decompression and bucket accumulation of the verifier's MSM) tests limb 0 and takes
// a cold branch.  Both produce the same value.
// ------------------------------------------------------------------------------------------------------
template <bool VT>
ZKP_DEV void fe_tail_add(uint32_t* r, uint32_t k) {
  if (VT) {
    r[0] += k;
    if (r[0] < k) {   // carry out of limb 0: cold
      uint32_t c = 1u;
#pragma unroll
      for (int i = 1; i < 8; i++) {
        r[i] += c;
        c = (r[i] < c) ? 1u : 0u;
      }
      r[0] += 38u * c;   // wrapped past 2^256: r is tiny, cannot carry again
    }
  } else {
    uint32_t c = add_word8(r, k);
    r[0] += 38u * c;
  }
}
template <bool VT>
ZKP_DEV void fe_tail_sub(uint32_t* r, uint32_t k) {
  if (VT) {
    uint32_t old = r[0];
    r[0] -= k;
    if (old < k) {   // borrow out of limb 0: cold
      uint32_t b = 1u;
#pragma unroll
      for (int i = 1; i < 8; i++) {
        uint32_t o = r[i];
        r[i] -= b;
        b = (o < b) ? 1u : 0u;
      }
      r[0] -= 38u * b;   // wrapped below 0: r is just under 2^256, cannot borrow again
    }
  } else {
    uint32_t b = sub_word8(r, k);
    r[0] -= 38u * b;
  }
}

// ------------------------------------------------------------------------------------------------------
// reduction of a 16-limb product t (value < 2^512) to 8 limbs (< 2^256): 2^256 = 38 (mod p)
// ------------------------------------------------------------------------------------------------------
template <bool VT>
ZKP_DEV void fe_reduce512_t(fe& r, const uint32_t* t) {
  uint32_t lo[9], od[8];
#pragma unroll
  for (int i = 0; i < 8; i++) lo[i] = t[i];
  // even high limbs: 38*t[8,10,12,14] lands on pairs (0,1)(2,3)(4,5)(6,7); carry has weight 2^256
  lo[8] = mad4(lo, t[8], t[10], t[12], t[14], 38u);
  // odd high limbs: 38*t[9,11,13,15] lands on limbs 1..8
  mulw(od[0], od[1], t[9], 38u);
  mulw(od[2], od[3], t[11], 38u);
  mulw(od[4], od[5], t[13], 38u);
  mulw(od[6], od[7], t[15], 38u);
  add8_nc(lo + 1, lo + 1, od);        // lo[8] <= 1 + 37 + 1: no carry out
  fe_tail_add<VT>(lo, lo[8] * 38u);
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = lo[i];
}
ZKP_DEV void fe_reduce512(fe& r, const uint32_t* t) { fe_reduce512_t<false>(r, t); }

// Variant of fe_reduce512 that multiplies the high half by 38 = 2^5 + 2^2 + 2^1 with funnel shifts and adds on the
// ALU pipe instead of eight wide IMADs on the (saturated) FMA pipe.  Selected with -DZKP_REDUCE_SHIFTADD.
ZKP_DEV uint32_t shl_limb(uint32_t lo, uint32_t hi, int s) {
#if ZKP_DEVICE_ASM
  return __funnelshift_l(lo, hi, s);
#else
  return (hi << s) | (lo >> (32 - s));
#endif
}
ZKP_DEV void fe_reduce512_shiftadd(fe& r, const uint32_t* t) {
  uint32_t a5[9], a2[9], a1[9], s1[9], s2[9], lo[9];
  const uint32_t* H = t + 8;
  a5[0] = H[0] << 5; a2[0] = H[0] << 2; a1[0] = H[0] << 1;
#pragma unroll
  for (int i = 1; i < 8; i++) {
    a5[i] = shl_limb(H[i - 1], H[i], 5);
    a2[i] = shl_limb(H[i - 1], H[i], 2);
    a1[i] = shl_limb(H[i - 1], H[i], 1);
  }
  a5[8] = H[7] >> 27; a2[8] = H[7] >> 30; a1[8] = H[7] >> 31;
  uint32_t c = add8(s1, a5, a2);
  s1[8] = a5[8] + a2[8] + c;
  c = add8(s2, s1, a1);
  s2[8] = s1[8] + a1[8] + c;           // 38*H as nine limbs, s2[8] <= 37
  c = add8(lo, t, s2);
  lo[8] = s2[8] + c;                   // weight 2^256, <= 38
  uint32_t c3 = add_word8(lo, lo[8] * 38u);
  lo[0] += 38u * c3;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = lo[i];
}
#ifndef ZKP_REDUCE_DEFAULT
#define ZKP_REDUCE_DEFAULT 0   // 0 = wide-IMAD reduction, 1 = shift-add reduction
#endif
template <int RED, bool VT>
ZKP_DEV void fe_reduce512_sel(fe& r, const uint32_t* t) {
  if (RED) fe_reduce512_shiftadd(r, t);
  else fe_reduce512_t<VT>(r, t);
}

// r = a*b mod p (weak: r < 2^256) -- round-1 schedule: every row's carry is captured into a fresh limb (kept for the
// micro-benchmark that measures the difference: zkp_bench_field kinds 17, 18)
template <int RED, bool VT = false>
ZKP_DEV void fe_mul_v1(fe& r, const fe& a, const fe& b) {
  const uint32_t* A = a.v;
  const uint32_t* B = b.v;
  uint32_t E[17], O[17];
#pragma unroll
  for (int i = 0; i < 17; i++) { E[i] = 0; O[i] = 0; }
  // row 0 (fresh accumulators: plain wide multiplies)
  mulw(E[0], E[1], A[0], B[0]); mulw(E[2], E[3], A[2], B[0]);
  mulw(E[4], E[5], A[4], B[0]); mulw(E[6], E[7], A[6], B[0]);
  mulw(O[0], O[1], A[1], B[0]); mulw(O[2], O[3], A[3], B[0]);
  mulw(O[4], O[5], A[5], B[0]); mulw(O[6], O[7], A[7], B[0]);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    // odd row i: even a -> O[i-1 ..], carry into the fresh limb above; odd a -> E[i+1 ..]
    O[i + 7] = mad4(O + i - 1, A[0], A[2], A[4], A[6], B[i]);
    mad4(E + i + 1, A[1], A[3], A[5], A[7], B[i]);
    if (i + 1 < 8) {
      // even row i+1: even a -> E[i+1 ..], carry into fresh limb; odd a -> O[i+1 ..]
      E[i + 9] = mad4(E + i + 1, A[0], A[2], A[4], A[6], B[i + 1]);
      mad4(O + i + 1, A[1], A[3], A[5], A[7], B[i + 1]);
    }
  }
  // t = E + (O << 32): one 15-limb chain (E[16] = O[15] = 0 and the product fits in 512 bits)
  uint32_t t[16];
  t[0] = E[0];
  add15(t + 1, E + 1, O);
  fe_reduce512_sel<RED, VT>(r, t);
}
// r = a*b mod p (weak: r < 2^256): the same 64 + 8 wide products, scheduled so that no carry is ever materialised
// between chains (see "chains whose carry never leaves the chain" above): the even-a part of an odd row also computes the
// top product of the following row into a fresh pair, and that row ripples its carry into the pair.
template <int RED, bool VT = false>
ZKP_DEV void fe_mul_t(fe& r, const fe& a, const fe& b) {
  const uint32_t* A = a.v;
  const uint32_t* B = b.v;
  uint32_t E[17], O[17];
#pragma unroll
  for (int i = 0; i < 17; i++) { E[i] = 0; O[i] = 0; }
  // row 0 (fresh accumulators: plain wide multiplies)
  mulw(E[0], E[1], A[0], B[0]); mulw(E[2], E[3], A[2], B[0]);
  mulw(E[4], E[5], A[4], B[0]); mulw(E[6], E[7], A[6], B[0]);
  mulw(O[0], O[1], A[1], B[0]); mulw(O[2], O[3], A[3], B[0]);
  mulw(O[4], O[5], A[5], B[0]); mulw(O[6], O[7], A[7], B[0]);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    // odd row i, even a -> O[i-1 .. i+6]; with it the top product a7 * b[i+1] of the next row into the fresh pair O[i+7, i+8]
    if (i + 1 < 8) mad4x1(O + i - 1, A[0], A[2], A[4], A[6], B[i], A[7], B[i + 1]);
    else O[i + 7] = mad4(O + i - 1, A[0], A[2], A[4], A[6], B[i]);   // last row: the carry is the top limb
    // odd row i, odd a -> E[i+1 .. i+8]: the top pair is fresh for i = 1 and was made by the previous even row afterwards
    if (i == 1) mad3x1(E + 2, A[1], A[3], A[5], B[1], A[7], B[1]);
    else mad3r(E + i + 1, A[1], A[3], A[5], B[i]);
    if (i + 1 < 8) {
      // even row i+1, even a -> E[i+1 .. i+8]; with it the top product a7 * b[i+2] of the next odd row into E[i+9, i+10]
      mad4x1(E + i + 1, A[0], A[2], A[4], A[6], B[i + 1], A[7], B[i + 2]);
      // even row i+1, odd a -> O[i+1 .. i+6], carry into the pair made above
      mad3r(O + i + 1, A[1], A[3], A[5], B[i + 1]);
    }
  }
  // t = E + (O << 32): one 15-limb chain (E[16] = O[15] = 0 and the product fits in 512 bits)
  uint32_t t[16];
  t[0] = E[0];
  add15(t + 1, E + 1, O);
  fe_reduce512_sel<RED, VT>(r, t);
}
ZKP_DEV void fe_mul(fe& r, const fe& a, const fe& b) { fe_mul_t<ZKP_REDUCE_DEFAULT, false>(r, a, b); }
ZKP_DEV void fe_mul_vt(fe& r, const fe& a, const fe& b) { fe_mul_t<ZKP_REDUCE_DEFAULT, true>(r, a, b); }

// r = a^2 mod p -- round-1 schedule (see fe_mul_v1)
template <int RED, bool VT = false>
ZKP_DEV void fe_sq_v1(fe& r, const fe& a) {
  const uint32_t* A = a.v;
  uint32_t E[17], O[17];
#pragma unroll
  for (int i = 0; i < 17; i++) { E[i] = 0; O[i] = 0; }
  // cross products a_i*a_j, i<j; (i+j) even -> E[i+j], odd -> O[i+j-1]
  // row 0
  mulw(E[2], E[3], A[2], A[0]); mulw(E[4], E[5], A[4], A[0]); mulw(E[6], E[7], A[6], A[0]);
  mulw(O[0], O[1], A[1], A[0]); mulw(O[2], O[3], A[3], A[0]);
  mulw(O[4], O[5], A[5], A[0]); mulw(O[6], O[7], A[7], A[0]);
  // row 1: j=3,5,7 -> E[4..9]; j=2,4,6 -> O[2..7]
  mad3(E + 4, A[3], A[5], A[7], A[1]);
  O[8] = mad3(O + 2, A[2], A[4], A[6], A[1]);
  // row 2: j=4,6 -> E[6..9]; j=3,5,7 -> O[4..9]
  E[10] = mad2(E + 6, A[4], A[6], A[2]);
  mad3(O + 4, A[3], A[5], A[7], A[2]);
  // row 3: j=5,7 -> E[8..11]; j=4,6 -> O[6..9]
  mad2(E + 8, A[5], A[7], A[3]);
  O[10] = mad2(O + 6, A[4], A[6], A[3]);
  // row 4: j=6 -> E[10,11]; j=5,7 -> O[8..11]
  E[12] = mad1(E + 10, A[6], A[4]);
  mad2(O + 8, A[5], A[7], A[4]);
  // row 5: j=7 -> E[12,13]; j=6 -> O[10,11]
  mad1(E + 12, A[7], A[5]);
  O[12] = mad1(O + 10, A[6], A[5]);
  // row 6: j=7 -> O[12,13]
  mad1(O + 12, A[7], A[6]);
  // cross = E + (O << 32): limbs 1..15
  uint32_t x[16];
  x[0] = 0;
  add15(x + 1, E + 1, O);          // one chain; x[15] picks up the last carry
  // double it (funnel shifts: independent, no carry chain)
  uint32_t t[16];
  t[0] = 0;
#pragma unroll
  for (int i = 1; i < 16; i++) t[i] = (x[i] << 1) | (x[i - 1] >> 31);
  // add the squares a_i^2 at limb 2i: one chain of eight fused lo/hi pairs
  sqr8(t, A);
  fe_reduce512_sel<RED, VT>(r, t);
}
// r = a^2 mod p: 28 cross products + 8 squares + 8 for the reduction, no carry materialised between chains
template <int RED, bool VT = false>
ZKP_DEV void fe_sq_t(fe& r, const fe& a) {
  const uint32_t* A = a.v;
  uint32_t E[17], O[17];
#pragma unroll
  for (int i = 0; i < 17; i++) { E[i] = 0; O[i] = 0; }
  // cross products a_i*a_j, i<j; (i+j) even -> pair E[i+j, i+j+1], odd -> pair O[i+j-1, i+j]
  // row 0: fresh pairs
  mulw(E[2], E[3], A[2], A[0]); mulw(E[4], E[5], A[4], A[0]); mulw(E[6], E[7], A[6], A[0]);
  mulw(O[0], O[1], A[1], A[0]); mulw(O[2], O[3], A[3], A[0]);
  mulw(O[4], O[5], A[5], A[0]); mulw(O[6], O[7], A[7], A[0]);
  // odd sums: (1,2)(1,4)(1,6) + fresh (2,7) | (2,3)(2,5) ripple | (3,4)(3,6) + fresh (4,7) | (4,5) ripple | (5,6) + fresh (6,7)
  mad3x1(O + 2, A[2], A[4], A[6], A[1], A[7], A[2]);
  mad2r(O + 4, A[3], A[5], A[2]);
  mad2x1(O + 6, A[4], A[6], A[3], A[7], A[4]);
  mad1r(O + 8, A[5], A[4]);
  mad1x1(O + 10, A[6], A[5], A[7], A[6]);
  // even sums: (1,3)(1,5) + fresh (1,7) | (2,4)(2,6) + fresh (3,7) | (3,5) ripple | (4,6) + fresh (5,7)
  mad2x1(E + 4, A[3], A[5], A[1], A[7], A[1]);
  mad2x1(E + 6, A[4], A[6], A[2], A[7], A[3]);
  mad1r(E + 8, A[5], A[3]);
  mad1x1(E + 10, A[6], A[4], A[7], A[5]);
  // cross = E + (O << 32): limbs 1..15
  uint32_t x[16];
  x[0] = 0;
  add15(x + 1, E + 1, O);          // one chain; x[15] picks up the last carry
  // double it (funnel shifts: independent, no carry chain)
  uint32_t t[16];
  t[0] = 0;
#pragma unroll
  for (int i = 1; i < 16; i++) t[i] = (x[i] << 1) | (x[i - 1] >> 31);
  // add the squares a_i^2 at limb 2i: one chain of eight fused lo/hi pairs
  sqr8(t, A);
  fe_reduce512_sel<RED, VT>(r, t);
}
ZKP_DEV void fe_sq(fe& r, const fe& a) { fe_sq_t<ZKP_REDUCE_DEFAULT, false>(r, a); }
ZKP_DEV void fe_sq_vt(fe& r, const fe& a) { fe_sq_t<ZKP_REDUCE_DEFAULT, true>(r, a); }

// r = a + b
template <bool VT>
ZKP_DEV void fe_add_t(fe& r, const fe& a, const fe& b) {
  uint32_t t[8];
  uint32_t c = add8(t, a.v, b.v);
  fe_tail_add<VT>(t, 38u * c);
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = t[i];
}
ZKP_DEV void fe_add(fe& r, const fe& a, const fe& b) { fe_add_t<false>(r, a, b); }
ZKP_DEV void fe_add_vt(fe& r, const fe& a, const fe& b) { fe_add_t<true>(r, a, b); }

// r = a - b
template <bool VT>
ZKP_DEV void fe_sub_t(fe& r, const fe& a, const fe& b) {
  uint32_t t[8];
  uint32_t br = sub8(t, a.v, b.v);
  fe_tail_sub<VT>(t, 38u * br);
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = t[i];
}
ZKP_DEV void fe_sub(fe& r, const fe& a, const fe& b) { fe_sub_t<false>(r, a, b); }
ZKP_DEV void fe_sub_vt(fe& r, const fe& a, const fe& b) { fe_sub_t<true>(r, a, b); }
// policy-dispatched spellings used by the templated chains below and in ge.cuh
template <bool VT> ZKP_DEV void fe_mulx(fe& r, const fe& a, const fe& b) { fe_mul_t<ZKP_REDUCE_DEFAULT, VT>(r, a, b); }
template <bool VT> ZKP_DEV void fe_sqx(fe& r, const fe& a) { fe_sq_t<ZKP_REDUCE_DEFAULT, VT>(r, a); }

ZKP_DEV void fe_zero(fe& r) {
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
}
ZKP_DEV void fe_one(fe& r) {
  fe_zero(r);
  r.v[0] = 1;
}
ZKP_DEV void fe_neg(fe& r, const fe& a) {
  fe z;
  fe_zero(z);
  fe_sub(r, z, a);
}
// r = 2a
ZKP_DEV void fe_dbl(fe& r, const fe& a) { fe_add(r, a, a); }

// r = flag ? b : a   (flag is 0/1; branch-free)
ZKP_DEV void fe_select(fe& r, const fe& a, const fe& b, uint32_t flag) {
  uint32_t m = 0u - flag;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = a.v[i] ^ (m & (a.v[i] ^ b.v[i]));
}
ZKP_DEV void fe_cswap(fe& a, fe& b, uint32_t flag) {
  uint32_t m = 0u - flag;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t x = m & (a.v[i] ^ b.v[i]);
    a.v[i] ^= x;
    b.v[i] ^= x;
  }
}

// canonical representative in [0, p)
ZKP_DEV void fe_canon(fe& r, const fe& a) {
  uint32_t t[8];
#pragma unroll
  for (int i = 0; i < 8; i++) t[i] = a.v[i];
  // fold bit 255 twice: afterwards t < 2^255 + 19 - ... then a conditional subtract of p
  uint32_t top = t[7] >> 31;
  t[7] &= 0x7fffffffu;
  add_word8(t, 19u * top);  // t < 2^255 + 19
  top = t[7] >> 31;
  t[7] &= 0x7fffffffu;
  add_word8(t, 19u * top);  // t < 2^255 now (if the first fold overflowed bit 255 the rest is tiny)
  // t >= p  <=>  t + 19 >= 2^255
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; i++) u[i] = t[i];
  add_word8(u, 19u);
  uint32_t ge = u[7] >> 31;
  u[7] &= 0x7fffffffu;
  uint32_t m = 0u - ge;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = t[i] ^ (m & (t[i] ^ u[i]));
}

ZKP_DEV uint32_t fe_is_zero(const fe& a) {
  fe c;
  fe_canon(c, a);
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= c.v[i];
  return o == 0 ? 1u : 0u;
}
ZKP_DEV uint32_t fe_is_negative(const fe& a) {
  fe c;
  fe_canon(c, a);
  return c.v[0] & 1u;
}
ZKP_DEV uint32_t fe_eq(const fe& a, const fe& b) {
  fe d;
  fe_sub(d, a, b);
  return fe_is_zero(d);
}
// r = neg ? -a : a
ZKP_DEV void fe_cneg(fe& r, const fe& a, uint32_t neg) {
  fe n;
  fe_neg(n, a);
  fe_select(r, a, n, neg);
}
ZKP_DEV void fe_abs(fe& r, const fe& a) { fe_cneg(r, a, fe_is_negative(a)); }

// r = a^(2^k)
template <bool VT = false>
ZKP_DEV void fe_sqn(fe& r, const fe& a, int k) {
  fe t = a;
#if ZKP_DEVICE_ASM
#pragma unroll 1
#endif
  for (int i = 0; i < k; i++) fe_sqx<VT>(t, t);
  r = t;
}

// a^(2^250 - 1) and a^11 (the shared head of invert and pow22523; dalek field.rs pow22501 [ext])
template <bool VT = false>
ZKP_DEV void fe_pow22501(fe& t19, fe& t3, const fe& a) {
  fe t0, t1, t2, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18;
  fe_sqx<VT>(t0, a);          // 2
  fe_sqn<VT>(t1, t0, 2);      // 8
  fe_mulx<VT>(t2, a, t1);     // 9
  fe_mulx<VT>(t3, t0, t2);    // 11
  fe_sqx<VT>(t4, t3);         // 22
  fe_mulx<VT>(t5, t2, t4);    // 31 = 2^5-1
  fe_sqn<VT>(t6, t5, 5);
  fe_mulx<VT>(t7, t6, t5);    // 2^10-1
  fe_sqn<VT>(t8, t7, 10);
  fe_mulx<VT>(t9, t8, t7);    // 2^20-1
  fe_sqn<VT>(t10, t9, 20);
  fe_mulx<VT>(t11, t10, t9);  // 2^40-1
  fe_sqn<VT>(t12, t11, 10);
  fe_mulx<VT>(t13, t12, t7);  // 2^50-1
  fe_sqn<VT>(t14, t13, 50);
  fe_mulx<VT>(t15, t14, t13); // 2^100-1
  fe_sqn<VT>(t16, t15, 100);
  fe_mulx<VT>(t17, t16, t15); // 2^200-1
  fe_sqn<VT>(t18, t17, 50);
  fe_mulx<VT>(t19, t18, t13); // 2^250-1
}
// a^(p-2)
ZKP_DEV void fe_invert(fe& r, const fe& a) {
  fe t19, t3, t20;
  fe_pow22501(t19, t3, a);
  fe_sqn(t20, t19, 5);
  fe_mul(r, t20, t3);
}
// a^((p-5)/8) = a^(2^252 - 3)
template <bool VT = false>
ZKP_DEV void fe_pow22523(fe& r, const fe& a) {
  fe t19, t3, t20;
  fe_pow22501<VT>(t19, t3, a);
  fe_sqn<VT>(t20, t19, 2);
  fe_mulx<VT>(r, a, t20);
}

// 32 little-endian bytes (as 8 words) -> element; bit 255 is ignored like dalek's from_bytes.
ZKP_DEV void fe_from_words(fe& r, const uint32_t* w) {
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = w[i];
  r.v[7] &= 0x7fffffffu;
}
ZKP_DEV void fe_to_words(uint32_t* w, const fe& a) {
  fe c;
  fe_canon(c, a);
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = c.v[i];
}

// dalek FieldElement51 limb form (5 x u64, 51 bits each, possibly unreduced up to 2^54) <-> fe
ZKP_DEV void fe_from_limbs51(fe& r, const uint64_t* l) {
  // value = sum l[i] 2^(51 i); limbs may exceed 51 bits slightly: accumulate with carries, fold 2^255 -> 19
  unsigned long long acc[5];
#pragma unroll
  for (int i = 0; i < 5; i++) acc[i] = l[i];
  // carry-propagate to 51 bits
#pragma unroll
  for (int rep = 0; rep < 2; rep++) {
    unsigned long long c = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
      acc[i] += c;
      c = acc[i] >> 51;
      acc[i] &= 0x7ffffffffffffULL;
    }
    acc[0] += 19ULL * c;
  }
  // after two passes every limb is < 2^51: pack 5x51 -> 256 bits
  unsigned long long w0 = acc[0] | (acc[1] << 51);
  unsigned long long w1 = (acc[1] >> 13) | (acc[2] << 38);
  unsigned long long w2 = (acc[2] >> 26) | (acc[3] << 25);
  unsigned long long w3 = (acc[3] >> 39) | (acc[4] << 12);
  r.v[0] = (uint32_t)w0; r.v[1] = (uint32_t)(w0 >> 32);
  r.v[2] = (uint32_t)w1; r.v[3] = (uint32_t)(w1 >> 32);
  r.v[4] = (uint32_t)w2; r.v[5] = (uint32_t)(w2 >> 32);
  r.v[6] = (uint32_t)w3; r.v[7] = (uint32_t)(w3 >> 32);
}
ZKP_DEV void fe_to_limbs51(uint64_t* l, const fe& a) {
  fe c;
  fe_canon(c, a);
  unsigned long long w0 = c.v[0] | ((unsigned long long)c.v[1] << 32);
  unsigned long long w1 = c.v[2] | ((unsigned long long)c.v[3] << 32);
  unsigned long long w2 = c.v[4] | ((unsigned long long)c.v[5] << 32);
  unsigned long long w3 = c.v[6] | ((unsigned long long)c.v[7] << 32);
  const unsigned long long M = 0x7ffffffffffffULL;
  l[0] = w0 & M;
  l[1] = ((w0 >> 51) | (w1 << 13)) & M;
  l[2] = ((w1 >> 38) | (w2 << 26)) & M;
  l[3] = ((w2 >> 25) | (w3 << 39)) & M;
  l[4] = (w3 >> 12) & M;
}

}  // namespace zkp
