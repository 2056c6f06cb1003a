// Field-multiplier micro-benchmarks (zkp_bench_field): the integer-pipe roofline calibration and the
// "51-bit radix vs saturated 32-bit" ablation that DESIGN.md reports.  Each thread runs a dependent chain of
// `iters` operations on register-resident operands; with 8 resident warps per scheduler the chain latency is
// hidden and the measured rate is the pipe's throughput for that instruction mix.
#pragma once
#include "fe.cuh"
#include "ge.cuh"
#ifdef ZKP_ABLATIONS
#include "fe64.cuh"   // FP64-pipe field arithmetic: a measured ablation, not in the product library
#endif

namespace zkp {

template <int RED, bool VT, int VER = 2>   // VER 1 = the round-1 schedule (carries captured into fresh limbs)
__global__ void __launch_bounds__(256) k_bench_mul32(uint32_t* out, int iters) {
  uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  fe a, b;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a.v[i] = gid * 2654435761u + i * 40503u + 1;
    b.v[i] = gid * 2246822519u + i * 3266489917u + 7;
  }
#pragma unroll 1
  for (int k = 0; k < iters; k += 2) {
    if (VER == 1) { fe_mul_v1<RED, VT>(a, a, b); fe_mul_v1<RED, VT>(b, b, a); }
    else { fe_mul_t<RED, VT>(a, a, b); fe_mul_t<RED, VT>(b, b, a); }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) out[(size_t)gid * 16 + i] = a.v[i] ^ b.v[i];
}

template <int RED, bool VT, int VER = 2>
__global__ void __launch_bounds__(256) k_bench_sq32(uint32_t* out, int iters) {
  uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  fe a;
#pragma unroll
  for (int i = 0; i < 8; i++) a.v[i] = gid * 2654435761u + i * 40503u + 1;
#pragma unroll 1
  for (int k = 0; k < iters; k += 2) {
    if (VER == 1) { fe_sq_v1<RED, VT>(a, a); fe_sq_v1<RED, VT>(a, a); }
    else { fe_sq_t<RED, VT>(a, a); fe_sq_t<RED, VT>(a, a); }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) out[(size_t)gid * 16 + i] = a.v[i];
}

// ---- signed mixed addition chain (the inner operation of bucket accumulation): acc += (+/-) q, register-resident
template <bool VT>
__global__ void __launch_bounds__(128, 4) k_bench_madd(uint32_t* out, int iters) {
  uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  ge_ext acc;
  ge_aniels q;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    acc.X.v[i] = gid * 2654435761u + i * 40503u + 1;
    acc.Y.v[i] = gid * 2246822519u + i * 3266489917u + 7;
    acc.Z.v[i] = gid * 668265263u + i * 374761393u + 3;
    acc.T.v[i] = gid * 2870177450u + i * 2147483647u + 5;
    q.yplusx.v[i] = gid * 1597334677u + i * 3812015801u + 11;
    q.yminusx.v[i] = gid * 958689277u + i * 1103515245u + 13;
    q.xy2d.v[i] = gid * 3323815723u + i * 12345u + 17;
  }
#pragma unroll 1
  for (int k = 0; k < iters; k++) ge_madd_signed<VT>(acc, acc, q, (uint32_t)k & 1u);
#pragma unroll
  for (int i = 0; i < 8; i++) out[(size_t)gid * 16 + i] = acc.X.v[i] ^ acc.Y.v[i] ^ acc.Z.v[i] ^ acc.T.v[i];
}

#ifdef ZKP_ABLATIONS
// ---- integer and FP64 squaring chains side by side: warps [0, NFP) of every 8-warp block square on the FP64 pipe
// (fe64_sq), the others on the integer pipes (fe_sq, variable-time tail).  Work is handed out dynamically: a warp
// takes the next unit (a chain of 256 squarings per lane) from a global counter until `units` are done, so slower
// warps simply take fewer units -- the schedule k_decompress uses.  Figure of merit: squarings per second over all
// warps (zkp_bench_field kinds 12..16 = NFP 0, 2, 3, 4, 8).
template <int NFP>
__global__ void __launch_bounds__(256, 2) k_bench_sq_mixed(uint32_t* out, int units, unsigned int* counter) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  const bool fp = (int)(threadIdx.x >> 5) < NFP;
  uint32_t x = 0;
  for (;;) {
    unsigned int u = 0;
    if (lane == 0) u = atomicAdd(counter, 1u);
    u = __shfl_sync(0xffffffffu, u, 0);
    if (u >= (unsigned int)units) break;
    fe a;
#pragma unroll
    for (int i = 0; i < 8; i++) a.v[i] = (u * 32 + lane) * 2654435761u + i * 40503u + 1;
    if (fp) {
      fe64 t;
      fe64_from_fe(t, a);
#pragma unroll 1
      for (int k = 0; k < 256; k++) fe64_sq(t, t);
      fe64_to_fe(a, t);
    } else {
#pragma unroll 1
      for (int k = 0; k < 256; k += 2) {
        fe_sq_vt(a, a);
        fe_sq_vt(a, a);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= a.v[i];
  }
  out[gid] = x;
}

#endif  // ZKP_ABLATIONS

// ---- 5 x 51-bit limbs, 64x64->128 products (the layout of dalek's u64 backend, FieldElement51::mul [ext]) --
struct fe51 { unsigned long long v[5]; };
__device__ __forceinline__ void fe51_mul(fe51& r, const fe51& a, const fe51& b) {
  typedef unsigned __int128 u128;
  const unsigned long long M = 0x7ffffffffffffULL;
  unsigned long long b1_19 = b.v[1] * 19, b2_19 = b.v[2] * 19, b3_19 = b.v[3] * 19, b4_19 = b.v[4] * 19;
  u128 c0 = (u128)a.v[0] * b.v[0] + (u128)a.v[4] * b1_19 + (u128)a.v[3] * b2_19 + (u128)a.v[2] * b3_19 + (u128)a.v[1] * b4_19;
  u128 c1 = (u128)a.v[1] * b.v[0] + (u128)a.v[0] * b.v[1] + (u128)a.v[4] * b2_19 + (u128)a.v[3] * b3_19 + (u128)a.v[2] * b4_19;
  u128 c2 = (u128)a.v[2] * b.v[0] + (u128)a.v[1] * b.v[1] + (u128)a.v[0] * b.v[2] + (u128)a.v[4] * b3_19 + (u128)a.v[3] * b4_19;
  u128 c3 = (u128)a.v[3] * b.v[0] + (u128)a.v[2] * b.v[1] + (u128)a.v[1] * b.v[2] + (u128)a.v[0] * b.v[3] + (u128)a.v[4] * b4_19;
  u128 c4 = (u128)a.v[4] * b.v[0] + (u128)a.v[3] * b.v[1] + (u128)a.v[2] * b.v[2] + (u128)a.v[1] * b.v[3] + (u128)a.v[0] * b.v[4];
  c1 += (unsigned long long)(c0 >> 51);
  unsigned long long o0 = (unsigned long long)c0 & M;
  c2 += (unsigned long long)(c1 >> 51);
  unsigned long long o1 = (unsigned long long)c1 & M;
  c3 += (unsigned long long)(c2 >> 51);
  unsigned long long o2 = (unsigned long long)c2 & M;
  c4 += (unsigned long long)(c3 >> 51);
  unsigned long long o3 = (unsigned long long)c3 & M;
  unsigned long long carry = (unsigned long long)(c4 >> 51);
  unsigned long long o4 = (unsigned long long)c4 & M;
  o0 += carry * 19;
  o1 += o0 >> 51;
  o0 &= M;
  r.v[0] = o0; r.v[1] = o1; r.v[2] = o2; r.v[3] = o3; r.v[4] = o4;
}

__global__ void __launch_bounds__(256) k_bench_mul51(unsigned long long* out, int iters) {
  uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  fe51 a, b;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    a.v[i] = ((unsigned long long)gid * 0x9E3779B97F4A7C15ULL + i * 0xD1B54A32D192ED03ULL) & 0x7ffffffffffffULL;
    b.v[i] = ((unsigned long long)gid * 0xC2B2AE3D27D4EB4FULL + i * 0x165667B19E3779F9ULL) & 0x7ffffffffffffULL;
  }
#pragma unroll 1
  for (int k = 0; k < iters; k += 2) {
    fe51_mul(a, a, b);
    fe51_mul(b, b, a);
  }
#pragma unroll
  for (int i = 0; i < 5; i++) out[(size_t)gid * 8 + i] = a.v[i] ^ b.v[i];
}

// ---- 10 x 25.5-bit limbs, 32x32->64 products accumulated lazily (dalek's u32 backend / ref10 layout) ------
struct fe25 { uint32_t v[10]; };
__device__ __forceinline__ void fe25_mul(fe25& r, const fe25& f, const fe25& g) {
  uint32_t g19[10], f2[10];
#pragma unroll
  for (int i = 0; i < 10; i++) { g19[i] = 19u * g.v[i]; f2[i] = (i & 1) ? 2u * f.v[i] : f.v[i]; }
  unsigned long long h[10];
#pragma unroll
  for (int k = 0; k < 10; k++) {
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
      int j = k - i;
      bool wrap = j < 0;
      if (wrap) j += 10;
      // odd*odd products carry a factor 2 (26/25-bit alternating radix)
      uint32_t fi = ((i & 1) && (j & 1)) ? f2[i] : f.v[i];
      uint32_t gj = wrap ? g19[j] : g.v[j];
      s += (unsigned long long)fi * gj;
    }
    h[k] = s;
  }
  // carry chain: limbs alternate 26 / 25 bits
  unsigned long long c;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    int bits = (i & 1) ? 25 : 26;
    c = h[i] >> bits;
    h[i] &= (1ull << bits) - 1;
    h[i + 1] += c;
  }
  c = h[9] >> 25;
  h[9] &= (1ull << 25) - 1;
  h[0] += 19 * c;
  c = h[0] >> 26;
  h[0] &= (1ull << 26) - 1;
  h[1] += c;
#pragma unroll
  for (int i = 0; i < 10; i++) r.v[i] = (uint32_t)h[i];
}

__global__ void __launch_bounds__(256) k_bench_mul25(uint32_t* out, int iters) {
  uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  fe25 a, b;
#pragma unroll
  for (int i = 0; i < 10; i++) {
    a.v[i] = (gid * 2654435761u + i * 40503u + 1) & 0x1ffffffu;
    b.v[i] = (gid * 2246822519u + i * 3266489917u + 7) & 0x1ffffffu;
  }
#pragma unroll 1
  for (int k = 0; k < iters; k += 2) {
    fe25_mul(a, a, b);
    fe25_mul(b, b, a);
  }
#pragma unroll
  for (int i = 0; i < 10; i++) out[(size_t)gid * 16 + i] = a.v[i] ^ b.v[i];
}


// ---- raw wide-multiply issue rate with realistic operand variety: 32 IMAD.WIDE per iteration --------------
// kind 4: plain mad.wide.u32 into eight independent 64-bit accumulators (no carries)
__global__ void __launch_bounds__(256) k_bench_wide_plain(uint32_t* out, int iters) {
  uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t a[8], b[4];
  unsigned long long acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = gid * 2654435761u + i * 40503u + 1; acc[i] = gid + i; }
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = gid * 2246822519u + i * 3266489917u + 7;
#pragma unroll 1
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] += (unsigned long long)a[(i + j) & 7] * b[j];
    b[0] ^= (uint32_t)acc[0];
  }
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) x ^= (uint32_t)acc[i] ^ (uint32_t)(acc[i] >> 32);
  out[gid] = x;
}
// kind 5: the same 32 products as eight carry chains of four (mad4), the shape fe_mul uses
__global__ void __launch_bounds__(256) k_bench_wide_carry(uint32_t* out, int iters) {
  uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t a[8], b[8], e[8], o[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = gid * 2654435761u + i * 40503u + 1;
    b[i] = gid * 2246822519u + i * 3266489917u + 7;
    e[i] = gid + i; o[i] = gid ^ i;
  }
#pragma unroll 1
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      e[0] += mad4(e, a[0], a[2], a[4], a[6], b[2 * j]);
      o[0] += mad4(o, a[1], a[3], a[5], a[7], b[2 * j + 1]);
    }
    b[0] ^= e[7];
  }
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) x ^= e[i] ^ o[i];
  out[gid] = x;
}


#ifdef ZKP_ABLATIONS
// ---- can the FP64 pipe run beside the integer pipe?  Odd warps run the fe_sq chain, even warps run a chain of
// DFMAs (16 independent accumulators, `fp64_per_iter` DFMAs per loop trip).  mode 0: both, 1: only integer warps
// (the others exit), 2: only FP64 warps.  Reported through zkp_bench_dual.
__global__ void __launch_bounds__(256) k_bench_dual(uint32_t* out, int iters, int mode) {
  uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  const bool int_warp = ((threadIdx.x >> 5) & 1) != 0;
  if (int_warp) {
    if (mode == 2) return;
    fe a;
#pragma unroll
    for (int i = 0; i < 8; i++) a.v[i] = gid * 2654435761u + i * 40503u + 1;
#pragma unroll 1
    for (int k = 0; k < iters; k += 2) {
      fe_sq(a, a);
      fe_sq(a, a);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) out[(size_t)gid * 16 + i] = a.v[i];
  } else {
    if (mode == 1) return;
    double acc[16], m = 1.0000001 + gid * 1e-9, c = 0.5;
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = 1.0 + i + gid;
#pragma unroll 1
    for (int k = 0; k < iters; k += 2) {
#pragma unroll
      for (int r = 0; r < 20; r++)   // 2 x 160 DFMAs per trip = the op budget of two FP64 squarings
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[(size_t)gid * 16] = (uint32_t)__double2loint(s);
  }
}

#endif  // ZKP_ABLATIONS

}  // namespace zkp
