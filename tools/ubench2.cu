// Issue-cost model of the integer instructions the field arithmetic is made of (B200, sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench2.bin tools/ubench2.cu
// Each kernel runs 16 warps per SM sub-partition (8 blocks x 256 threads per SM); a loop trip issues NW wide multiplies
// (of one of three kinds) interleaved with NA ALU instructions (of one of three kinds) on independent registers whose
// operands are loop-carried, so ptxas can neither hoist nor fuse them (checked with cuobjdump -sass).
// Output: cycles of one SM sub-partition per loop trip and per instruction, at the clock measured in-kernel
// (clock64 against globaltimer on one thread).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define TRIPS 32768

// WK: 0 = IMAD.WIDE.U32 with 64-bit addend (mad.wide), 1 = IMAD.WIDE.U32 without addend (mul.wide),
//     2 = carry chain of fused lo/hi pairs (IMAD.WIDE.U32.X), 3 = 32-bit IMAD (mad.lo)
// AK: 0 = IADD3.X carry chain, 1 = SHF (funnel shift), 2 = LOP3, 3 = SEL-free plain IADD3 (three live inputs)
template <int NW, int WK, int NA, int AK>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, double* mhz) {
  uint32_t a[8], b[8], m[8];
  uint64_t w[8];
  double d[8], dm = 1.0000001 + seed * 1e-9, dc = 0.5;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = tid * 2654435761u + i * 40503u + seed;
    b[i] = tid * 2246822519u + i * 3266489917u + 7;
    m[i] = tid * 668265263u + i * 374761393u + 3;
    w[i] = ((uint64_t)a[i] << 32) | b[i];
    d[i] = 1.0 + i + tid * 1e-6;
  }
  long long c0 = 0; unsigned long long g0 = 0;
  if (tid == 0) { c0 = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0)); }
#pragma unroll 1
  for (int it = 0; it < TRIPS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (WK == 2 && NW > 0) {
        // NW/4 chains of 4 fused pairs; even chains on w[0..3], odd chains on w[4..7] (aligned register pairs)
#pragma unroll
        for (int ch = 0; ch < NW / 4; ch++) {
          uint64_t* q = w + 4 * (ch & 1);
          uint32_t l0 = (uint32_t)q[0], h0 = (uint32_t)(q[0] >> 32), l1 = (uint32_t)q[1], h1 = (uint32_t)(q[1] >> 32);
          uint32_t l2 = (uint32_t)q[2], h2 = (uint32_t)(q[2] >> 32), l3 = (uint32_t)q[3], h3 = (uint32_t)(q[3] >> 32);
          asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                       "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                       "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                       "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;"
                       : "+r"(l0), "+r"(h0), "+r"(l1), "+r"(h1), "+r"(l2), "+r"(h2), "+r"(l3), "+r"(h3)
                       : "r"(m[(ch * 4 + 0) & 7]), "r"(m[(ch * 4 + 1) & 7]), "r"(m[(ch * 4 + 2) & 7]), "r"(m[(ch * 4 + 3) & 7]),
                         "r"(m[(ch + r + 4) & 7]));
          q[0] = ((uint64_t)h0 << 32) | l0; q[1] = ((uint64_t)h1 << 32) | l1;
          q[2] = ((uint64_t)h2 << 32) | l2; q[3] = ((uint64_t)h3 << 32) | l3;
        }
      } else {
#pragma unroll
        for (int i = 0; i < NW; i++) {
          const int j = i & 7;
          // both multiplier operands come from other accumulators: nothing is loop-invariant, every bit is live
          if (WK == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"((uint32_t)w[(j + 1) & 7]), "r"((uint32_t)(w[(j + 3) & 7] >> 32)));
          if (WK == 1) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[j]) : "r"((uint32_t)w[(j + 1) & 7]), "r"((uint32_t)(w[(j + 3) & 7] >> 32)));
          if (WK == 3) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(m[j]) : "r"(m[(j + 1) & 7]), "r"(m[(j + 3) & 7]));
        }
      }
      if (AK == 0 && NA > 0) {
#pragma unroll
        for (int ch = 0; ch < NA / 8; ch++)
          asm volatile("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\t"
                       "addc.cc.u32 %3, %3, %11;\n\taddc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\t"
                       "addc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"
                       : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
                       : "r"(b[(ch + 0) & 7]), "r"(b[(ch + 1) & 7]), "r"(b[(ch + 2) & 7]), "r"(b[(ch + 3) & 7]),
                         "r"(b[(ch + 4) & 7]), "r"(b[(ch + 5) & 7]), "r"(b[(ch + 6) & 7]), "r"(b[(ch + 7) & 7]));
      } else {
#pragma unroll
        for (int i = 0; i < NA; i++) {
          const int j = i & 7;
          if (AK == 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[j]) : "r"(b[(j + r) & 7]));
          if (AK == 4) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[j]) : "d"(dm), "d"(d[(j + 3) & 7]));
          if (AK == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b[(j + r) & 7]), "r"(b[(j + 3) & 7]));
        }
      }
    }
    // keep the operands loop-carried
    b[0] ^= a[1];
    m[0] ^= (uint32_t)w[1];
    dc = d[2];
  }
  if (tid == 0) {
    long long c1 = clock64(); unsigned long long g1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
    *mhz = (double)(c1 - c0) / (double)(g1 - g0) * 1e3;
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc ^= a[i] ^ b[i] ^ m[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)__double2loint(d[i] + dc);
  out[tid] = acc;
}

static int g_sms;
static uint32_t* g_out;
static double* g_mhz;

template <int NW, int WK, int NA, int AK>
void run(const char* name) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NW, WK, NA, AK><<<g_sms * 8, 256>>>(g_out, 1, g_mhz);
  cudaEventRecord(e0); k<NW, WK, NA, AK><<<g_sms * 8, 256>>>(g_out, 2, g_mhz); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double mhz; cudaMemcpy(&mhz, g_mhz, 8, cudaMemcpyDeviceToHost);
  const double trips = 16.0 * TRIPS * 4.0;            // warp-trips per SM sub-partition (r-unrolled x4)
  const double cyc = ms * 1e-3 * mhz * 1e6 / trips;   // SMSP cycles per (NW wide + NA alu) group
  printf("%-52s %7.2f ms @%5.0f MHz  %6.2f cyc/group  %5.2f cyc/inst\n", name, ms, mhz, cyc, cyc / (NW + NA));
}

int main() {
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
  cudaMalloc(&g_out, (size_t)g_sms * 8 * 256 * 4); cudaMalloc(&g_mhz, 8);
  for (int i = 0; i < 12; i++) { k<8, 0, 8, 2><<<g_sms * 8, 256>>>(g_out, 0, g_mhz); }   // warm up the clocks
  cudaDeviceSynchronize();
  run<8, 0, 0, 0>("8 IMAD.WIDE (64-bit addend)");
  run<8, 1, 0, 0>("8 IMAD.WIDE (no addend)");
  run<8, 2, 0, 0>("8 IMAD.WIDE.X (2 carry chains of 4)");
  run<8, 3, 0, 0>("8 IMAD (32-bit)");
  run<0, 0, 8, 0>("8 IADD3.X (1 carry chain)");
  run<0, 0, 8, 1>("8 SHF");
  run<0, 0, 8, 2>("8 LOP3");
  run<8, 0, 8, 0>("8 IMAD.WIDE + 8 IADD3.X");
  run<8, 0, 16, 0>("8 IMAD.WIDE + 16 IADD3.X");
  run<8, 0, 8, 1>("8 IMAD.WIDE + 8 SHF");
  run<8, 0, 8, 2>("8 IMAD.WIDE + 8 LOP3");
  run<8, 0, 16, 2>("8 IMAD.WIDE + 16 LOP3");
  run<8, 0, 24, 2>("8 IMAD.WIDE + 24 LOP3");
  run<8, 1, 8, 0>("8 IMAD.WIDE(no addend) + 8 IADD3.X");
  run<8, 1, 16, 0>("8 IMAD.WIDE(no addend) + 16 IADD3.X");
  run<8, 2, 8, 0>("8 IMAD.WIDE.X + 8 IADD3.X");
  run<8, 2, 16, 0>("8 IMAD.WIDE.X + 16 IADD3.X");
  run<8, 2, 8, 2>("8 IMAD.WIDE.X + 8 LOP3");
  run<8, 2, 16, 2>("8 IMAD.WIDE.X + 16 LOP3");
  run<0, 0, 8, 4>("8 DFMA");
  run<8, 2, 8, 4>("8 IMAD.WIDE.X + 8 DFMA");
  run<8, 2, 16, 4>("8 IMAD.WIDE.X + 16 DFMA");
  run<8, 1, 8, 4>("8 IMAD.WIDE(no addend) + 8 DFMA");
  run<8, 3, 8, 4>("8 IMAD(32) + 8 DFMA");
  run<0, 0, 16, 4>("16 DFMA");
  run<8, 3, 8, 0>("8 IMAD(32) + 8 IADD3.X");
  run<8, 3, 8, 2>("8 IMAD(32) + 8 LOP3");
  run<8, 3, 16, 2>("8 IMAD(32) + 16 LOP3");
  return 0;
}
