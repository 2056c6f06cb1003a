// Instruction-throughput micro-benchmarks on the integer / FP64 pipes of one B200 (numbers for DESIGN.md).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
#define UNROLL 8
template <int KIND>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t a[UNROLL], b[UNROLL];
  uint64_t w[UNROLL];
  double d[UNROLL];
  uint32_t x = seed + threadIdx.x, y = seed * 3 + blockIdx.x + 1;
  for (int i = 0; i < UNROLL; i++) { a[i] = x + i; b[i] = y ^ i; w[i] = (uint64_t)x * (i + 3); d[i] = 1.0 + i + x; }
  double dm = 1.0000001 + seed, da = 0.5;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNROLL; i++) {
      if (KIND == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(y));
      if (KIND == 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(x), "r"(y));
      if (KIND == 2) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(x), "r"(y));
      if (KIND == 3) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(a[i]), "+r"(b[i]) : "r"(x), "r"(y));
      if (KIND == 4) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(da));
      if (KIND == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
      if (KIND == 6) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
      if (KIND == 7) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(y));
                       asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
                       asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(x)); }
      if (KIND == 8) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(y));
                       asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(da)); }
      if (KIND == 9) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(y));
      if (KIND == 10) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(a[i]), "+r"(b[i]) : "r"(y));
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < UNROLL; i++) acc ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)__double2loint(d[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int KIND>
void run(const char* name, int ops_per_iter) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* out; long long* cyc; cudaMalloc(&out, sms * 8 * 256 * 4); cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND><<<sms * 8, 256>>>(out, 1, cyc);
  cudaEventRecord(e0); k<KIND><<<sms * 8, 256>>>(out, 2, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  double ops = (double)sms * 8 * 256 * ITERS * UNROLL * ops_per_iter;
  // block 0's cycle count covers 8 resident blocks/SM sharing the SM: warp-instr per SMSP-cycle
  double per_sm_clk = ((double)8 * 256 * ITERS * UNROLL * ops_per_iter) / (double)hc;
  printf("%-34s %8.3f ms  %.3e lane-ops/s  %.1f lane-ops/clk/SM (block0 clk=%lld => %.0f MHz)\n", name, ms, ops / (ms * 1e-3),
         per_sm_clk, hc, (double)hc / (ms * 1e-3) / 1e6);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("IMAD.WIDE.U32 (mad.wide.u32)", 1);
  run<9>("IMAD.WIDE.U32 (mul.wide.u32)", 1);
  run<1>("IMAD (mad.lo.u32)", 1);
  run<2>("IMAD.HI (mad.hi.u32)", 1);
  run<3>("mad.lo.cc+madc.hi pair (fused)", 1);
  run<4>("DFMA (fma.rn.f64)", 1);
  run<5>("IADD3 (add.u32)", 1);
  run<10>("add.cc+addc pair", 2);
  run<6>("SHF (shf.l.wrap)", 1);
  run<7>("mix: 1 IMAD.WIDE + 2 IADD", 3);
  run<8>("mix: 1 IMAD.WIDE + 1 DFMA", 2);
  return 0;
}
