// Instruction-throughput micro-benchmarks on the integer / FP pipes of one B200 (numbers for DESIGN.md section 3).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench tools/ubench.cu
// Every kernel runs 16 warps per SM sub-partition (8 blocks x 256 threads per SM) of independent register-resident
// chains; the figure of merit is cycles of one SM sub-partition per warp instruction (1.0 = the issue limit).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 65536
#define UNROLL 8
template <int KIND>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t a[UNROLL], b[UNROLL], c[UNROLL];
  uint64_t w[UNROLL];
  double d[UNROLL];
  float f[UNROLL];
  uint32_t x = seed + threadIdx.x, y = seed * 3 + blockIdx.x + 1;
  for (int i = 0; i < UNROLL; i++) {
    a[i] = x + i; b[i] = y ^ i; c[i] = x * 7 + i; w[i] = (uint64_t)x * (i + 3); d[i] = 1.0 + i + x; f[i] = 1.0f + i + x;
  }
  double dm = 1.0000001 + seed, da = 0.5;
  float fm = 1.0000001f + seed, fa = 0.5f;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNROLL; i++) {
#define WIDE asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(y))
#define IADDA asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(y))
#define IADDB asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(x))
#define IADDC asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(x))
#define FFMA asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fm), "f"(fa))
#define IMADLO asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c[i]) : "r"(x), "r"(y))
      if (KIND == 0) WIDE;
      if (KIND == 1) IMADLO;
      if (KIND == 2) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(x), "r"(y));
      if (KIND == 3) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(a[i]), "+r"(b[i]) : "r"(x), "r"(y));
      if (KIND == 4) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(da));
      if (KIND == 5) IADDA;
      if (KIND == 6) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
      if (KIND == 7) { WIDE; IADDA; IADDB; }
      if (KIND == 8) { WIDE; asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(da)); }
      if (KIND == 9) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(y));
      if (KIND == 10) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(a[i]), "+r"(b[i]) : "r"(y));
      if (KIND == 11) { WIDE; IADDA; }
      if (KIND == 12) { WIDE; IADDA; IADDB; IADDC; }
      if (KIND == 13) { IMADLO; IADDA; }
      if (KIND == 14) FFMA;
      if (KIND == 15) { FFMA; IADDA; }
      if (KIND == 16) { FFMA; WIDE; }
      if (KIND == 17) { FFMA; IMADLO; }
      if (KIND == 18) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(y));
      if (KIND == 19) { WIDE; asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i])); }
      if (KIND == 20) { FFMA; FFMA; IADDA; IADDB; }
      if (KIND == 21) { FFMA; IADDA; WIDE; }
      // three-input add (IADD3 with all three operands live)
      if (KIND == 22) asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t;}" : "+r"(a[i]) : "r"(b[i]), "r"(y));
      // 64-bit add = IADD3 + IADD3.X
      if (KIND == 23) asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) & (UNROLL - 1)]));
      // mul.wide into fresh registers then 64-bit add (the "lazy accumulation" shape)
      if (KIND == 24) { uint64_t p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a[i]), "r"(y));
                        asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(p)); }
      if (KIND == 25) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(a[i]) : "r"(b[i]));
      if (KIND == 26) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(a[i]) : "r"(b[i]), "r"(y));
      if (KIND == 27) asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(a[i]) : "r"(b[i]), "r"(y));
    }
    // carry-chain shapes: one chain per trip over the 8 registers
    if (KIND == 30) {   // 4 fused lo/hi pairs with carry in/out (the mad4 of fe_mul), twice
      asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                   "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                   "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                   "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.cc.u32 %7, %11, %12, %7;\n\t"
                   : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
                   : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(x));
      asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                   "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                   "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                   "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.cc.u32 %7, %11, %12, %7;\n\t"
                   : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7])
                   : "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(y));
    }
    if (KIND == 31) {   // 8-limb add with carry (IADD3.X chain), twice
      asm volatile("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t"
                   "addc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\taddc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"
                   : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
                   : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
      asm volatile("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t"
                   "addc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\taddc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"
                   : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7])
                   : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    }
    if (KIND == 32) {   // 8 independent mul.wide into fresh registers, merged by two 8-limb carry adds (16 ops of each)
      uint32_t lo[8], hi[8];
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("{.reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0,%1}, t;}" : "=r"(lo[i]), "=r"(hi[i]) : "r"(b[i]), "r"(x));
      asm volatile("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t"
                   "addc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\taddc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"
                   : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
                   : "r"(lo[0]), "r"(hi[0]), "r"(lo[2]), "r"(hi[2]), "r"(lo[4]), "r"(hi[4]), "r"(lo[6]), "r"(hi[6]));
      asm volatile("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t"
                   "addc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\taddc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"
                   : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7])
                   : "r"(lo[1]), "r"(hi[1]), "r"(lo[3]), "r"(hi[3]), "r"(lo[5]), "r"(hi[5]), "r"(lo[7]), "r"(hi[7]));
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < UNROLL; i++)
    acc ^= a[i] ^ b[i] ^ c[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)__double2loint(d[i]) ^ __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int KIND>
void run(const char* name, int ops_per_iter, int per_trip = 0) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* out; long long* cyc; cudaMalloc(&out, sms * 8 * 256 * 4); cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND><<<sms * 8, 256>>>(out, 1, cyc);
  cudaEventRecord(e0); k<KIND><<<sms * 8, 256>>>(out, 2, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  // warp instructions issued per SM sub-partition while block 0 ran: 16 warps x trips x instructions per trip
  double per_trip_insts = per_trip ? (double)per_trip : (double)UNROLL * ops_per_iter;
  double warp_insts = 16.0 * ITERS * per_trip_insts;
  // two views: block 0's own clock64 span, and the event time at the nominal 1965 MHz (they agree once clocks are up)
  printf("%-44s %8.3f ms  %6.3f cyc/warp-inst/SMSP by clock64, %6.3f by events@1965MHz  (block0 clk/ms => %.0f MHz)\n", name, ms,
         (double)hc / warp_insts, ms * 1e-3 * 1.965e9 / warp_insts, (double)hc / (ms * 1e-3) / 1e6);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int i = 0; i < 20; i++) run<14>("warm-up (FFMA)", 1);   // ~1 s of load so the SM clock is at its maximum
  run<0>("IMAD.WIDE.U32 (mad.wide.u32)", 1);
  run<9>("IMAD.WIDE.U32 (mul.wide.u32)", 1);
  run<1>("IMAD (mad.lo.u32)", 1);
  run<27>("IMAD (mul.lo.u32)", 1);
  run<2>("IMAD.HI (mad.hi.u32)", 1);
  run<26>("IMAD.HI (mul.hi.u32)", 1);
  run<3>("mad.lo.cc+madc.hi pair (1 fused inst)", 1);
  run<4>("DFMA (fma.rn.f64)", 1);
  run<14>("FFMA", 1);
  run<5>("IADD3 (add.u32)", 1);
  run<22>("IADD3 three live inputs", 1);
  run<10>("add.cc+addc pair", 2);
  run<23>("add.u64 (IADD3 + IADD3.X)", 2);
  run<6>("SHF (shf.l.wrap)", 1);
  run<18>("LOP3", 1);
  run<25>("PRMT", 1);
  run<11>("mix: 1 IMAD.WIDE + 1 IADD", 2);
  run<7>("mix: 1 IMAD.WIDE + 2 IADD", 3);
  run<12>("mix: 1 IMAD.WIDE + 3 IADD", 4);
  run<19>("mix: 1 IMAD.WIDE + 1 SHF", 2);
  run<13>("mix: 1 IMAD.lo + 1 IADD", 2);
  run<15>("mix: 1 FFMA + 1 IADD", 2);
  run<20>("mix: 2 FFMA + 2 IADD", 4);
  run<16>("mix: 1 FFMA + 1 IMAD.WIDE", 2);
  run<17>("mix: 1 FFMA + 1 IMAD.lo", 2);
  run<21>("mix: 1 FFMA + 1 IADD + 1 IMAD.WIDE", 3);
  run<8>("mix: 1 IMAD.WIDE + 1 DFMA", 2);
  run<24>("mul.wide + add.u64 (3 inst)", 3);
  run<30>("2 x mad4 carry chain (8 IMAD.WIDE.X)", 0, 8);
  run<31>("2 x 8-limb add.cc chain (16 IADD3.X)", 0, 16);
  run<32>("8 mul.wide + 2 x 8-limb carry add (24 inst)", 0, 24);
  return 0;
}
