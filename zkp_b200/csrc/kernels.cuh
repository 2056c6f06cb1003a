// CUDA kernels of the ristretto255 MSM engine (sm_100a).
//
// The variable-time path replaces curve25519-dalek 2.x `backend/serial/scalar_mul/pippenger.rs` +
// `ristretto.rs` decompress/compress [ext] as reached from
//   /root/reference/src/toolbox/batch_verifier.rs:219-230 (the north-star call), verifier.rs:97-106, :162-168.
// Pipeline for one MSM of n terms, window width c, W = ceil(253/c) windows, B = 2^(c-1) buckets per window:
//   k_decompress      32 B encoding -> 96 B affine-Niels point (y+x, y-x, 2dxy) in HBM, validity via atomicMin
//   k_recode_hist     scalar -> sign fold -> signed digits; histogram of bucket loads  (hist[W][B])
//   k_scan            exclusive scan per window                                         (offs[W][B+1], cursor)
//   k_scatter         counting-sort scatter of (sign | term index) by bucket            (sorted[W][n])
//   k_plan/k_items    cut buckets into work items of <= S entries (bounded work per thread for any digit mix)
//   k_accumulate      one item per thread: gather Niels points (software-prefetched), mixed additions (7M)
//   k_merge           add the partial sums of buckets that were cut
//   k_chunk_reduce    running-sum reduction of bucket rows in chunks of L (recursive levels)
//   k_tree_sum        per-window sums of the chunk partials
//   k_finish          per-window Horner over levels, Horner over windows, ristretto encode, identity/status
// All group arithmetic is in registers (fe.cuh / ge.cuh); HBM traffic is 64 B in per term plus the
// workspace streams listed in DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fe.cuh"
#include "ge.cuh"
#include "sc.cuh"

// An affine Niels point (y+x, y-x, 2dxy: 96 bytes) occupies ZKP_NIELS_U4 uint4 slots of the workspace.  With 8 (128 bytes,
// 128-byte aligned) the gather of the bucket accumulation touches exactly ONE line per point; with 6 (packed) a point
// straddles two lines half of the time and HBM delivers whole lines (measured: 196 B per 96-byte point).
#ifndef ZKP_NIELS_U4
#define ZKP_NIELS_U4 8
#endif

namespace zkp {

#define ZKP_CHUNK_L 32      // chunk length of the bucket running-sum reduction
#define ZKP_CHUNK_LOG 5

struct msm_result {          // layout documented in include/zkp_b200.h (zkp_msm_vartime_dev)
  uint32_t enc[8];
  int32_t status;
  int32_t is_identity;
  long long first_bad;
};

// flags[0] = first invalid point index (atomicMin, init 0x7fffffff), flags[1] = first bad scalar index
__device__ __forceinline__ void load_words8(uint32_t* w, const uint4* p) {
  uint4 a = __ldg(p), b = __ldg(p + 1);
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
  w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
__device__ __forceinline__ void store_fe(uint4* p, const fe& a) {
  p[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
  p[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
__device__ __forceinline__ void load_fe(fe& a, const uint4* p) {
  uint4 x = p[0], y = p[1];
  a.v[0] = x.x; a.v[1] = x.y; a.v[2] = x.z; a.v[3] = x.w;
  a.v[4] = y.x; a.v[5] = y.y; a.v[6] = y.z; a.v[7] = y.w;
}
__device__ __forceinline__ void load_fe_ldg(fe& a, const uint4* p) {
  uint4 x = __ldg(p), y = __ldg(p + 1);
  a.v[0] = x.x; a.v[1] = x.y; a.v[2] = x.z; a.v[3] = x.w;
  a.v[4] = y.x; a.v[5] = y.y; a.v[6] = y.z; a.v[7] = y.w;
}
__device__ __forceinline__ void store_ext(uint4* p, const ge_ext& q) {
  store_fe(p, q.X); store_fe(p + 2, q.Y); store_fe(p + 4, q.Z); store_fe(p + 6, q.T);
}
__device__ __forceinline__ void load_ext(ge_ext& q, const uint4* p) {
  load_fe(q.X, p); load_fe(q.Y, p + 2); load_fe(q.Z, p + 4); load_fe(q.T, p + 6);
}

// ---------------------------------------------------------------------------------------------------------
// K1: decompression.  One thread per point; ~254 squarings + ~25 multiplications, all in registers.
// ---------------------------------------------------------------------------------------------------------
// `base` = global index of element 0 (chunked ingestion reports global indices of bad points)
__global__ void __launch_bounds__(256) k_decompress(const uint4* __restrict__ enc, size_t n, uint4* __restrict__ niels,
                                                    int* __restrict__ flags, size_t base) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t w[8];
  load_words8(w, enc + 2 * i);
  fe x, y, t;
  uint32_t ok = ristretto_decode<true>(x, y, t, w);
  ge_aniels q;
  if (ok) {
    ge_aniels_from_affine(q, x, y, t);
  } else {
    ge_aniels_identity(q);
    atomicMin(&flags[0], (int)(base + i));
  }
  uint4* o = niels + ZKP_NIELS_U4 * i;
  store_fe(o, q.yplusx);
  store_fe(o + 2, q.yminusx);
  store_fe(o + 4, q.xy2d);
}

// decompress to FieldElement51 limb form (zkp_decompress_batch)
__global__ void __launch_bounds__(256, 3) k_decompress_limbs(const uint4* __restrict__ enc, size_t n,
                                                          unsigned long long* __restrict__ limbs,
                                                          uint8_t* __restrict__ valid) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t w[8];
  load_words8(w, enc + 2 * i);
  fe x, y, t, z;
  uint32_t ok = ristretto_decode(x, y, t, w);
  if (!ok) { fe_zero(x); fe_one(y); fe_zero(t); }
  fe_one(z);
  unsigned long long* o = limbs + 20 * i;
  fe_to_limbs51((uint64_t*)o, x);
  fe_to_limbs51((uint64_t*)o + 5, y);
  fe_to_limbs51((uint64_t*)o + 10, z);
  fe_to_limbs51((uint64_t*)o + 15, t);
  valid[i] = (uint8_t)ok;
}

__device__ __forceinline__ void load_ext_limbs51(ge_ext& p, const unsigned long long* l) {
  fe_from_limbs51(p.X, (const uint64_t*)l);
  fe_from_limbs51(p.Y, (const uint64_t*)l + 5);
  fe_from_limbs51(p.Z, (const uint64_t*)l + 10);
  fe_from_limbs51(p.T, (const uint64_t*)l + 15);
}

// compress from limb form (zkp_compress_batch)
__global__ void __launch_bounds__(256) k_compress_limbs(const unsigned long long* __restrict__ limbs, size_t n,
                                                        uint4* __restrict__ enc) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge_ext p;
  load_ext_limbs51(p, limbs + 20 * i);
  uint32_t w[8];
  ristretto_encode(w, p);
  enc[2 * i] = make_uint4(w[0], w[1], w[2], w[3]);
  enc[2 * i + 1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// ---------------------------------------------------------------------------------------------------------
// K2/K4: scalar recode + histogram / scatter (counting sort by bucket, per window)
// ---------------------------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_recode(const uint4* __restrict__ scalars, size_t n, int c, int W, uint32_t B,
                                                uint32_t* __restrict__ counters,  // hist[W][B] or cursor[W][B]
                                                uint32_t* __restrict__ sorted,    // [W][n] (SCATTER only)
                                                int* __restrict__ flags, size_t base) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[8], k[8];
  load_words8(s, scalars + 2 * i);
  uint32_t neg;
  uint32_t canonical = sc_fold_sign(k, neg, s);
  if (!canonical) {
    if (!SCATTER) atomicMin(&flags[1], (int)(base + i));
    return;  // contributes nothing; the call fails with ZKP_ERR_SCALAR
  }
  uint32_t carry = 0;
  for (int w = 0; w < W; w++) {
    uint32_t mag, dneg;
    sc_next_digit(mag, dneg, carry, k, c);
    if (mag != 0) {
      uint32_t* ctr = counters + (size_t)w * B + (mag - 1);
      if (SCATTER) {
        uint32_t pos = atomicAdd(ctr, 1u);
        sorted[(size_t)w * n + pos] = (uint32_t)i | ((neg ^ dneg) << 31);
      } else {
        atomicAdd(ctr, 1u);
      }
    }
  }
}

#ifdef ZKP_ABLATIONS   // single-phase predecessor of k_ingest2 (api.cu: fused_sort = 0 outside profiling mode)
// K1+K2 fused: the digit histogram of term i is issued (fire-and-forget L2 reductions) before the ~31k-instruction
// decompression of point i, whose LSU pipe is otherwise idle.  Used outside profiling mode.
__global__ void __launch_bounds__(256) k_ingest(const uint4* __restrict__ enc, const uint4* __restrict__ scalars, size_t n,
                                                uint4* __restrict__ niels, int c, int W, uint32_t B,
                                                uint32_t* __restrict__ hist, int* __restrict__ flags, size_t base) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  {
    uint32_t s[8], k[8];
    load_words8(s, scalars + 2 * i);
    uint32_t neg;
    uint32_t canonical = sc_fold_sign(k, neg, s);
    if (!canonical) {
      atomicMin(&flags[1], (int)(base + i));
    } else {
      uint32_t carry = 0;
      for (int w = 0; w < W; w++) {
        uint32_t mag, dneg;
        sc_next_digit(mag, dneg, carry, k, c);
        if (mag != 0) atomicAdd(hist + (size_t)w * B + (mag - 1), 1u);
      }
    }
  }
  uint32_t w8[8];
  load_words8(w8, enc + 2 * i);
  fe x, y, t;
  uint32_t ok = ristretto_decode<true>(x, y, t, w8);
  ge_aniels q;
  if (ok) {
    ge_aniels_from_affine(q, x, y, t);
  } else {
    ge_aniels_identity(q);
    atomicMin(&flags[0], (int)(base + i));
  }
  uint4* o = niels + ZKP_NIELS_U4 * i;
  store_fe(o, q.yplusx);
  store_fe(o + 2, q.yminusx);
  store_fe(o + 4, q.xy2d);
}

#endif  // ZKP_ABLATIONS

// Two-phase ingestion: every thread first issues the digit work of up to two terms -- MODE 0: histogram reductions,
// MODE 1: counting-sort scatter (atomic cursor + 4-byte store) -- and then decompresses one point.  The L2 atomics and
// scattered stores ride under the ~31k-instruction decode whose LSU pipe is otherwise idle, so with
//   phase 1 = first half of the points + histogram of ALL scalars,  scan,  phase 2 = second half + scatter of ALL scalars
// neither the histogram nor the scatter costs a separate pass.  Thread t takes the terms sA_lo + t (t < sA_cnt) and
// sB_lo + t (t < sB_cnt) and the point p_lo + t (t < p_cnt); all indices are global term indices.
template <int MODE>
__device__ __forceinline__ void ingest_digits(const uint4* __restrict__ scalars, size_t i, size_t n, int c, int W, uint32_t B,
                                              uint32_t* __restrict__ counters, uint32_t* __restrict__ sorted,
                                              int* __restrict__ flags, bool batched) {
  uint32_t s[8], k[8];
  load_words8(s, scalars + 2 * i);
  uint32_t neg;
  uint32_t canonical = sc_fold_sign(k, neg, s);
  if (!canonical) {
    if (MODE == 0) atomicMin(&flags[1], (int)i);
    return;   // contributes nothing; the call fails with ZKP_ERR_SCALAR
  }
  uint32_t carry = 0;
  if (MODE == 1 && batched) {
    // scatter, four windows at a time: the four cursor atomics are in flight together and the four stores wait for
    // them afterwards (one round trip to L2 per group instead of one per window)
    for (int w0 = 0; w0 < W && !sc_digits_done(k, carry); w0 += 4) {   // (the higher windows of a 128-bit weight are empty)
      uint32_t pos[4], val[4];
      bool has[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int w = w0 + u;
        has[u] = false;
        if (w < W) {
          uint32_t mag, dneg;
          sc_next_digit(mag, dneg, carry, k, c);
          if (mag != 0) {
            has[u] = true;
            val[u] = (uint32_t)i | ((neg ^ dneg) << 31);
            pos[u] = atomicAdd(counters + (size_t)w * B + (mag - 1), 1u);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (has[u]) sorted[(size_t)(w0 + u) * n + pos[u]] = val[u];
    }
    return;
  }
  for (int w = 0; w < W && !sc_digits_done(k, carry); w++) {
    uint32_t mag, dneg;
    sc_next_digit(mag, dneg, carry, k, c);
    if (mag != 0) {
      uint32_t* ctr = counters + (size_t)w * B + (mag - 1);
      if (MODE == 0) {
        atomicAdd(ctr, 1u);
      } else {
        uint32_t pos = atomicAdd(ctr, 1u);
        sorted[(size_t)w * n + pos] = (uint32_t)i | ((neg ^ dneg) << 31);
      }
    }
  }
}

// VAR picks the occupancy point (block size, minimum resident blocks): 0 = 256 threads, 106 registers, 16 warps/SM;
// 1 = (128, 5): <= 96 registers, 20 warps/SM; 2 = (256, 3) and 3 = (128, 6): <= 80 registers, 24 warps/SM;
// 4 = (128, 7): 72 registers, 28 warps/SM (default: the spills it adds, 180 bytes per thread, lie outside the squaring loops);
// 5 = (128, 8): 64 registers, 32 warps/SM.  Measured, both ingestion launches of a step: 0: 91.0 ms (round 1 arithmetic),
// 2: 85.8, 3: 85.7, 4: 84.2, 5: 84.2 ms.
#define ZKP_INGEST_THREADS(VAR) ((VAR) == 0 || (VAR) == 2 ? 256 : 128)
#define ZKP_INGEST_MINBLK(VAR) ((VAR) == 0 ? 1 : (VAR) == 1 ? 5 : (VAR) == 2 ? 3 : (VAR) == 3 ? 6 : (VAR) == 4 ? 7 : 8)
// One launch = up to three term ranges (A, B, C) and one point range; thread t takes the terms s_lo[j] + t (t < s_cnt[j])
// and the point p_lo + t (t < p_cnt).  blockIdx.y = one of several equally shaped launches whose starts are y_p / y_s[j]
// apart: the rows of a batch-verification slab go out as ONE grid (one partial last wave instead of one per row).
struct ingest_args {
  size_t p_lo, p_cnt, y_p;
  size_t s_lo[3], s_cnt[3], y_s[3];
  int batched;   // scatter: the cursor atomics of four windows in flight together
};
template <int MODE, int VAR>
__global__ void __launch_bounds__(ZKP_INGEST_THREADS(VAR), ZKP_INGEST_MINBLK(VAR))
    k_ingest2(const uint4* __restrict__ enc, uint4* __restrict__ niels, const uint4* __restrict__ scalars,
              const __grid_constant__ ingest_args a, size_t n, int c, int W, uint32_t B, uint32_t* __restrict__ counters,
              uint32_t* __restrict__ sorted, int* __restrict__ flags) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 1
  for (int j = 0; j < 3; j++)
    if (t < a.s_cnt[j])
      ingest_digits<MODE>(scalars, a.s_lo[j] + (size_t)blockIdx.y * a.y_s[j] + t, n, c, W, B, counters, sorted, flags,
                          a.batched != 0);
  const size_t p_cnt = a.p_cnt, p_lo = a.p_lo + (size_t)blockIdx.y * a.y_p;
  if (t >= p_cnt) return;
  const size_t i = p_lo + t;
  uint32_t w8[8];
  load_words8(w8, enc + 2 * i);
  fe x, y, tt;
  uint32_t ok = ristretto_decode<true>(x, y, tt, w8);
  ge_aniels q;
  if (ok) {
    ge_aniels_from_affine(q, x, y, tt);
  } else {
    ge_aniels_identity(q);
    atomicMin(&flags[0], (int)i);
  }
  uint4* o = niels + ZKP_NIELS_U4 * i;
  store_fe(o, q.yplusx);
  store_fe(o + 2, q.yminusx);
  store_fe(o + 4, q.xy2d);
}

// K3: per-window exclusive scan.  One block (1024 threads) per window.
__global__ void __launch_bounds__(1024) k_scan(const uint32_t* __restrict__ hist, uint32_t B,
                                               uint32_t* __restrict__ offs,    // [W][B+1]
                                               uint32_t* __restrict__ cursor)  // [W][B]
{
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t base_s;
  const int w = blockIdx.x;
  const uint32_t* h = hist + (size_t)w * B;
  uint32_t* o = offs + (size_t)w * (B + 1);
  uint32_t* cu = cursor + (size_t)w * B;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (uint32_t start = 0; start < B; start += 1024) {
    uint32_t idx = start + threadIdx.x;
    uint32_t v = idx < B ? h[idx] : 0u;
    // inclusive warp scan
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if ((threadIdx.x & 31) >= d) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t t = warp_tot[threadIdx.x];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, t, d);
        if (threadIdx.x >= d) t += y;
      }
      warp_tot[threadIdx.x] = t;  // inclusive totals of warps
    }
    __syncthreads();
    uint32_t warp_base = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0u;
    uint32_t excl = base_s + warp_base + x - v;
    if (idx < B) { o[idx] = excl; cu[idx] = excl; }
    __syncthreads();
    if (threadIdx.x == 1023) base_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) o[B] = base_s;
}

// Multi-block exclusive scan of a long uint32 array (the item offsets over all W * B buckets): each block of 1024
// threads scans 4096 elements and publishes its total, one block scans the totals, a third pass adds them back.
// out has n + 1 entries (out[n] = grand total).
#define ZKP_SCAN_TILE 4096
__global__ void __launch_bounds__(1024) k_scan_tiles(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out,
                                                     uint32_t* __restrict__ tile_tot) {
  __shared__ uint32_t warp_tot[32];
  const uint32_t base = blockIdx.x * ZKP_SCAN_TILE + threadIdx.x * 4;
  uint32_t v[4], sum = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    v[i] = base + i < n ? in[base + i] : 0u;
    sum += v[i];
  }
  uint32_t x = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if ((threadIdx.x & 31) >= d) x += y;
  }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t t = warp_tot[threadIdx.x];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, t, d);
      if (threadIdx.x >= d) t += y;
    }
    warp_tot[threadIdx.x] = t;
  }
  __syncthreads();
  uint32_t excl = ((threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0u) + x - sum;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (base + i < n) out[base + i] = excl;
    excl += v[i];
  }
  if (threadIdx.x == 1023) tile_tot[blockIdx.x] = excl;
}
__global__ void __launch_bounds__(1024) k_scan_add(uint32_t* __restrict__ out, uint32_t n, const uint32_t* __restrict__ tile_offs) {
  const uint32_t i = blockIdx.x * 1024 + threadIdx.x;
  if (i < n) out[i] += tile_offs[i / ZKP_SCAN_TILE];
  if (i == 0) out[n] = tile_offs[(n + ZKP_SCAN_TILE - 1) / ZKP_SCAN_TILE];
}

// ---------------------------------------------------------------------------------------------------------
// K5: bucket accumulation.  G lanes cooperate on one bucket.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_aniels(ge_aniels& q, const uint4* __restrict__ niels, uint32_t e) {
  const uint4* p = niels + (size_t)(e & 0x7fffffffu) * ZKP_NIELS_U4;
  load_fe_ldg(q.yplusx, p);
  load_fe_ldg(q.yminusx, p + 2);
  load_fe_ldg(q.xy2d, p + 4);
}

__device__ __forceinline__ void shfl_down_ext(ge_ext& r, const ge_ext& p, int off, int width) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.X.v[i] = __shfl_down_sync(0xffffffffu, p.X.v[i], off, width);
    r.Y.v[i] = __shfl_down_sync(0xffffffffu, p.Y.v[i], off, width);
    r.Z.v[i] = __shfl_down_sync(0xffffffffu, p.Z.v[i], off, width);
    r.T.v[i] = __shfl_down_sync(0xffffffffu, p.T.v[i], off, width);
  }
}

// Work items: every bucket is cut into chunks of at most S sorted entries, so the work per thread is bounded
// whatever the digit distribution is (uniform scalars already overload the top window: after the sign fold its
// digits span only a fraction of the buckets).  k_plan counts chunks per bucket, an exclusive scan gives each
// bucket its first item, k_items writes the descriptors, k_accumulate sums one item per thread, k_merge adds
// the partial sums of multi-chunk buckets.
struct work_item { uint32_t start, end, out, wflag; };  // wflag = window | (partial ? 0x80000000 : 0)

__global__ void __launch_bounds__(256) k_plan(const uint32_t* __restrict__ offs, uint32_t B, uint32_t total_buckets,
                                              uint32_t S, uint32_t* __restrict__ chunk_cnt) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_buckets) return;
  uint32_t w = g / B, b = g % B;
  const uint32_t* o = offs + (size_t)w * (B + 1) + b;
  uint32_t cnt = o[1] - o[0];
  chunk_cnt[g] = cnt == 0 ? 1u : (cnt + S - 1) / S;   // an empty bucket keeps one (empty) item: it writes the identity
}

__global__ void __launch_bounds__(256) k_items(const uint32_t* __restrict__ offs, const uint32_t* __restrict__ item_offs,
                                               uint32_t B, uint32_t total_buckets, uint32_t S,
                                               work_item* __restrict__ items, uint32_t* __restrict__ multi,
                                               uint32_t* __restrict__ n_multi) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_buckets) return;
  uint32_t w = g / B, b = g % B;
  const uint32_t* o = offs + (size_t)w * (B + 1) + b;
  uint32_t start = o[0], end = o[1];
  uint32_t first = item_offs[g], cnt = item_offs[g + 1] - first;
  if (cnt == 1) {
    work_item it = {start, end, g, w};
    items[first] = it;
  } else {
    multi[atomicAdd(n_multi, 1u)] = g;   // buckets whose partial sums k_merge has to add
    for (uint32_t k = 0; k < cnt; k++) {
      uint32_t s0 = start + k * S, e0 = min(s0 + S, end);
      work_item it = {s0, e0, first + k, w | 0x80000000u};
      items[first + k] = it;
    }
  }
}

// Lane balance: the items of a warp should be equally long (a warp runs until its longest item is done; bucket loads
// are Poisson, 768 +- 28 at the bench size).  A counting sort of the item ids by length, longest first, gives every
// warp 32 items of (almost) the same length.  len_hist[S + 1] must be zero on entry.
__global__ void __launch_bounds__(256) k_len_hist(const work_item* __restrict__ items, const uint32_t* __restrict__ n_items,
                                                  uint32_t S, uint32_t* __restrict__ len_hist) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= *n_items) return;
  const uint4 raw = __ldg((const uint4*)items + id);
  const uint32_t len = raw.y - raw.x;
  // warp-aggregated: lanes with the same length elect one lane to add the group's size
  const uint32_t key = S - (len > S ? S : len);
  const uint32_t peers = __match_any_sync(__activemask(), key);
  if ((uint32_t)(__ffs(peers) - 1) == (threadIdx.x & 31)) atomicAdd(len_hist + key, (uint32_t)__popc(peers));
}
__global__ void __launch_bounds__(256) k_len_scatter(const work_item* __restrict__ items, const uint32_t* __restrict__ n_items,
                                                     uint32_t S, uint32_t* __restrict__ len_cursor,
                                                     uint32_t* __restrict__ order) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= *n_items) return;
  const uint4 raw = __ldg((const uint4*)items + id);
  const uint32_t len = raw.y - raw.x;
  const uint32_t key = S - (len > S ? S : len);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t peers = __match_any_sync(__activemask(), key);
  const uint32_t leader = (uint32_t)(__ffs(peers) - 1);
  uint32_t base = 0;
  if (leader == lane) base = atomicAdd(len_cursor + key, (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  order[base + __popc(peers & ((1u << lane) - 1u))] = id;
}

// MB = minimum resident blocks per SM (the occupancy point).  4: 122 registers, 16 warps/SM, the next point's gather prefetched
// in software (24 registers); 5 (default): 96 registers, 20 warps/SM, no prefetch -- the fifth warp per scheduler hides the gather
// better than the prefetch did: 35.9 -> 35.3 ms at the bench size (6 = 80 registers needs spills inside the loop: 35.9 ms;
// asking for the next point's line with prefetch.global.L2 / .L1, no register held: 39.5 ms -- session 29).
template <int MB>
__global__ void __launch_bounds__(128, MB) k_accumulate(const uint4* __restrict__ niels, const uint32_t* __restrict__ sorted,
                                                       const work_item* __restrict__ items,
                                                       const uint32_t* __restrict__ order,
                                                       const uint32_t* __restrict__ n_items, size_t n,
                                                       uint4* __restrict__ buckets, uint4* __restrict__ partials) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *n_items) return;
  const uint32_t id = order ? __ldg(order + slot) : slot;
  const uint4 raw = __ldg((const uint4*)items + id);
  const uint32_t start = raw.x, end = raw.y, out = raw.z, wflag = raw.w;
  const uint32_t* base = sorted + (size_t)(wflag & 0x7fffffffu) * n;
  ge_ext acc;
  ge_identity(acc);
  uint32_t i = start;
  if (MB >= 5) {   // more resident warps instead of the software prefetch (its 24 registers)
    while (i < end) {
      const uint32_t e = __ldg(base + i);
      ge_aniels q;
      load_aniels(q, niels, e);
      ge_madd_signed<true>(acc, acc, q, e >> 31);
      i++;
    }
  } else {
  ge_aniels cur;
  uint32_t e_cur = 0;
  if (i < end) {
    e_cur = __ldg(base + i);
    load_aniels(cur, niels, e_cur);
  }
  while (i < end) {
    uint32_t inext = i + 1;
    ge_aniels nxt;
    uint32_t e_nxt = 0;
    if (inext < end) {  // software prefetch of the next gather while this addition runs
      e_nxt = __ldg(base + inext);
      load_aniels(nxt, niels, e_nxt);
    }
    ge_madd_signed<true>(acc, acc, cur, e_cur >> 31);
    cur = nxt;
    e_cur = e_nxt;
    i = inext;
  }
  }
  uint4* dst = (wflag >> 31) ? partials + (size_t)out * 8 : buckets + (size_t)out * 8;
  store_ext(dst, acc);
}

// buckets with more than one chunk: one warp per such bucket adds the partial sums (lanes stride over the
// partials, then a warp-shuffle tree), so even a bucket that received every term is reduced in parallel.
__global__ void __launch_bounds__(128) k_merge(const uint32_t* __restrict__ item_offs, const uint32_t* __restrict__ multi,
                                               const uint32_t* __restrict__ n_multi, const uint4* __restrict__ partials,
                                               uint4* __restrict__ buckets) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t total = *n_multi;
  for (uint32_t idx = warp; idx < total; idx += nwarps) {
    const uint32_t g = multi[idx];
    const uint32_t first = item_offs[g], cnt = item_offs[g + 1] - first;
    ge_ext acc, p;
    ge_identity(acc);
    for (uint32_t k = lane; k < cnt; k += 32) {
      load_ext(p, partials + (size_t)(first + k) * 8);
      ge_add(acc, acc, p);
    }
    // shuffle tree over the lanes that hold a partial sum: a bucket cut into two or three items (the usual case) needs one
    // or two levels, not five (lanes >= cnt hold the identity; cnt is the same for the whole warp).  One 2^16-term MSM:
    // k_merge 0.56 -> 0.2 ms, the call 1.68 -> 1.33 ms.
    int top = 16;
    while (top >= 1 && (uint32_t)top >= cnt) top >>= 1;   // largest power of two below cnt (0 when cnt <= 1)
#pragma unroll 1
    for (int off = top; off >= 1; off >>= 1) {
      shfl_down_ext(p, acc, off, 32);
      ge_add(acc, acc, p);
    }
    if (lane == 0) store_ext(buckets + (size_t)g * 8, acc);
  }
}

// ---------------------------------------------------------------------------------------------------------
// K6: bucket reduction  S = sum_{i=0}^{m-1} (i+1) x_i  by chunked running sums (zero-based weights inside).
//   chunk k (items kL .. kL+L-1):  T_k = sum x,  U_k = sum_j j * x_{kL+j}
//   V(x) = sum_k U_k + L * V(T),  Tot(x) = Tot(T)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_chunk_reduce(const uint4* __restrict__ in, uint32_t m, int W,
                                                      uint4* __restrict__ outT, uint4* __restrict__ outU) {
  const uint32_t chunks = m / ZKP_CHUNK_L;
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= chunks * (uint32_t)W) return;
  uint32_t w = tid / chunks, k = tid % chunks;
  const uint4* x = in + ((size_t)w * m + (size_t)k * ZKP_CHUNK_L) * 8;
  ge_ext run, acc, p;
  ge_identity(run);
  ge_identity(acc);
#pragma unroll 1
  for (int j = ZKP_CHUNK_L - 1; j >= 1; j--) {
    load_ext(p, x + (size_t)j * 8);
    ge_add(run, run, p);
    ge_add(acc, acc, run);
  }
  load_ext(p, x);
  ge_add(run, run, p);
  store_ext(outT + ((size_t)w * chunks + k) * 8, run);
  store_ext(outU + ((size_t)w * chunks + k) * 8, acc);
}

// per-window sum of `cnt` points: one block of 256 threads per window
__global__ void __launch_bounds__(256) k_tree_sum(const uint4* __restrict__ in, uint32_t cnt, uint4* __restrict__ out) {
  __shared__ uint4 sm[8 * 8];  // one ext per warp
  const int w = blockIdx.x;
  const uint4* x = in + (size_t)w * cnt * 8;
  ge_ext acc, p;
  ge_identity(acc);
  for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
    load_ext(p, x + (size_t)i * 8);
    ge_add(acc, acc, p);
  }
#pragma unroll 1
  for (int off = 16; off >= 1; off >>= 1) {
    shfl_down_ext(p, acc, off, 32);
    ge_add(acc, acc, p);
  }
  if ((threadIdx.x & 31) == 0) store_ext(sm + (threadIdx.x >> 5) * 8, acc);
  __syncthreads();
  if (threadIdx.x == 0) {
    load_ext(acc, sm);
    for (int k = 1; k < (int)(blockDim.x >> 5); k++) {
      load_ext(p, sm + k * 8);
      ge_add(acc, acc, p);
    }
    store_ext(out + (size_t)w * 8, acc);
  }
}

// ---------------------------------------------------------------------------------------------------------
// K7: finish.  usum[level][W] (nl levels), last[W][m_last] = the items of the final (short) level.
// One block, W threads for the per-window part, thread 0 for the Horner over windows + encode.
// ---------------------------------------------------------------------------------------------------------
// Point operations shared by a group of four adjacent lanes (all four hold the same point before and after):
//   doubling   lane i squares one of X, Y, Z, X+Y; the results are exchanged by shuffles; the four lanes form E, F, G, H
//              redundantly; lane i computes one of the four products
//   mixed add  lanes 0..2 compute (Y1-X1)(y2-x2), (Y1+X1)(y2+x2), T1 * 2d x2 y2; exchange; then the same four products
// Per operation a lane runs two field multiplications (plus additions and 56-64 shuffles) instead of seven or eight: the
// serial chains that bound the LATENCY of a call (k_finish: ~253 doublings in a row; a single proof's MSMs) shrink ~2-3x.
// The instruction stream is the same on every lane (operands are picked with selects), so the shuffles are convergent.
// gmask / gbase: the lane mask of the group and its first lane within the warp.
__device__ __forceinline__ void shflg_fe(fe& out, const fe& mine, int src, unsigned gmask, int gbase) {
#pragma unroll
  for (int i = 0; i < 8; i++) out.v[i] = __shfl_sync(gmask, mine.v[i], gbase + src);
}
__device__ __forceinline__ void ge_efgh_products_coop(ge_ext& p, const fe& e, const fe& f, const fe& g, const fe& h, int sub,
                                                      unsigned gmask, int gbase) {
  // lane 0: E F (X), lane 1: G H (Y), lane 2: F G (Z), lane 3: E H (T)
  const uint32_t l1 = sub == 1, l2 = sub == 2, l3 = sub == 3;
  fe a, b, r;
  fe_select(a, e, g, l1);
  fe_select(a, a, f, l2);
  fe_select(b, f, h, l1);
  fe_select(b, b, g, l2);
  fe_select(b, b, h, l3);
  fe_mul(r, a, b);
  shflg_fe(p.X, r, 0, gmask, gbase);
  shflg_fe(p.Y, r, 1, gmask, gbase);
  shflg_fe(p.Z, r, 2, gmask, gbase);
  shflg_fe(p.T, r, 3, gmask, gbase);
}
__device__ __forceinline__ void ge_double_coop4(ge_ext& p, int sub, unsigned gmask = 0xfu, int gbase = 0) {
  const uint32_t l1 = sub == 1, l2 = sub == 2, l3 = sub == 3;
  fe in, xpy, sq, dbl;
  fe_add(xpy, p.X, p.Y);
  fe_select(in, p.X, p.Y, l1);
  fe_select(in, in, p.Z, l2);
  fe_select(in, in, xpy, l3);
  fe_sq(sq, in);
  fe_add(dbl, sq, sq);
  fe_select(sq, sq, dbl, l2);          // lane 2 carries 2 Z^2
  fe xx, yy, zz2, xy2, e, f, g, h;
  shflg_fe(xx, sq, 0, gmask, gbase);
  shflg_fe(yy, sq, 1, gmask, gbase);
  shflg_fe(zz2, sq, 2, gmask, gbase);
  shflg_fe(xy2, sq, 3, gmask, gbase);
  fe_add(h, yy, xx);                   // H = YY + XX
  fe_sub(g, yy, xx);                   // G = YY - XX
  fe_sub(e, xy2, h);                   // E = (X+Y)^2 - YY - XX
  fe_sub(f, zz2, g);                   // F = 2ZZ - G
  // doubling: X' = E F, Y' = G H, Z' = F G, T' = E H   (ge_double: r.X = e f, r.Y = h g, r.Z = g f, r.T = e h)
  ge_efgh_products_coop(p, e, f, g, h, sub, gmask, gbase);
}
__device__ __forceinline__ void ge_madd_coop4(ge_ext& p, const ge_aniels& q, int sub, unsigned gmask, int gbase) {
  const uint32_t l1 = sub == 1, l2 = sub == 2;
  fe ymx, ypx, a, b, r;
  fe_sub(ymx, p.Y, p.X);
  fe_add(ypx, p.Y, p.X);
  fe_select(a, ymx, ypx, l1);
  fe_select(a, a, p.T, l2);
  fe_select(b, q.yminusx, q.yplusx, l1);
  fe_select(b, b, q.xy2d, l2);
  fe_mul(r, a, b);                     // lane 0: A, lane 1: B, lane 2: C (lane 3 repeats A)
  fe A, B, C, d, e, f, g, h;
  shflg_fe(A, r, 0, gmask, gbase);
  shflg_fe(B, r, 1, gmask, gbase);
  shflg_fe(C, r, 2, gmask, gbase);
  fe_add(d, p.Z, p.Z);
  fe_sub(e, B, A);
  fe_sub(f, d, C);
  fe_add(g, d, C);
  fe_add(h, B, A);
  // ge_madd: r.X = e f, r.Y = g h, r.Z = f g, r.T = e h
  ge_efgh_products_coop(p, e, f, g, h, sub, gmask, gbase);
}

// p += q (projective Niels) shared by four lanes: the four products of the first round take one lane each
__device__ __forceinline__ void ge_add_pniels_coop4(ge_ext& p, const ge_pniels& q, int sub, unsigned gmask, int gbase) {
  const uint32_t l1 = sub == 1, l2 = sub == 2, l3 = sub == 3;
  fe ymx, ypx, a, b, r;
  fe_sub(ymx, p.Y, p.X);
  fe_add(ypx, p.Y, p.X);
  fe_select(a, ymx, ypx, l1);
  fe_select(a, a, p.T, l2);
  fe_select(a, a, p.Z, l3);
  fe_select(b, q.YminusX, q.YplusX, l1);
  fe_select(b, b, q.T2d, l2);
  fe_select(b, b, q.Z, l3);
  fe_mul(r, a, b);                     // lane 0: A, lane 1: B, lane 2: C, lane 3: Z1 Z2
  fe A, B, C, d, e, f, g, h;
  shflg_fe(A, r, 0, gmask, gbase);
  shflg_fe(B, r, 1, gmask, gbase);
  shflg_fe(C, r, 2, gmask, gbase);
  shflg_fe(d, r, 3, gmask, gbase);
  fe_add(d, d, d);
  fe_sub(e, B, A);
  fe_sub(f, d, C);
  fe_add(g, d, C);
  fe_add(h, B, A);
  ge_efgh_products_coop(p, e, f, g, h, sub, gmask, gbase);
}

__global__ void __launch_bounds__(64) k_finish(const uint4* __restrict__ usum, int nl, const uint4* __restrict__ last,
                                               uint32_t m_last, int W, int c, size_t n, const int* __restrict__ flags,
                                               msm_result* __restrict__ res, uint4* __restrict__ partial_out) {
  __shared__ uint4 sw[64 * 8];
  const int w = threadIdx.x;
  if (w < W) {
    // final level: V = sum k*x_k (zero-based), Tot = sum x_k, by running sums from the top
    ge_ext run, acc, p;
    ge_identity(run);
    ge_identity(acc);
    const uint4* x = last + (size_t)w * m_last * 8;
    for (int j = (int)m_last - 1; j >= 1; j--) {
      load_ext(p, x + (size_t)j * 8);
      ge_add(run, run, p);
      ge_add(acc, acc, run);
    }
    load_ext(p, x);
    ge_add(run, run, p);  // Tot
    // Horner over the chunk levels: V = U[l] + L * V
    for (int l = nl - 1; l >= 0; l--) {
      for (int d = 0; d < ZKP_CHUNK_LOG; d++) ge_double(acc, acc);
      load_ext(p, usum + ((size_t)l * W + w) * 8);
      ge_add(acc, acc, p);
    }
    ge_add(acc, acc, run);  // weights are one-based: S = V + Tot
    store_ext(sw + w * 8, acc);
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    // Horner over the windows: ~253 doublings in a row, shared by lanes 0..3 (ge_double_coop4); the W - 1 additions are
    // done redundantly by the four lanes so that they keep holding the same point
    ge_ext tot, p;
    load_ext(tot, sw + (W - 1) * 8);
    for (int ww = W - 2; ww >= 0; ww--) {
#pragma unroll 1
      for (int d = 0; d < c; d++) ge_double_coop4(tot, (int)threadIdx.x);
      load_ext(p, sw + ww * 8);
      ge_add(tot, tot, p);
    }
    if (threadIdx.x != 0) return;
    if (partial_out) store_ext(partial_out, tot);
    uint32_t enc[8];
    ristretto_encode(enc, tot);
    uint32_t z = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { res->enc[i] = enc[i]; z |= enc[i]; }
    int bad_pt = flags[0], bad_sc = flags[1];
    int status = 0;
    long long first_bad = -1;
    if (bad_pt != 0x7fffffff) { status = 1; first_bad = bad_pt; }
    else if (bad_sc != 0x7fffffff) { status = 3; first_bad = bad_sc; }
    res->status = status;
    res->is_identity = (z == 0) ? 1 : 0;
    res->first_bad = first_bad;
  }
}

// Single-verdict mode across shards: a shard's MSM result leaves the device as an extended point in
// FieldElement51 limb form (X, Y, Z, T: 20 x u64, the layout of include/zkp_b200.h), the verdict is taken on the sum.
__global__ void k_ext_to_limbs(const uint4* __restrict__ ext, unsigned long long* __restrict__ limbs) {
  ge_ext p;
  load_ext(p, ext);
  fe_to_limbs51((uint64_t*)limbs, p.X);
  fe_to_limbs51((uint64_t*)limbs + 5, p.Y);
  fe_to_limbs51((uint64_t*)limbs + 10, p.Z);
  fe_to_limbs51((uint64_t*)limbs + 15, p.T);
}
// sum of `count` limb-form points, ristretto encode, coset-aware identity test (one thread: count is the shard count)
__global__ void k_sum_partials(const unsigned long long* __restrict__ limbs, size_t count, msm_result* __restrict__ res) {
  ge_ext acc, p;
  ge_identity(acc);
  for (size_t i = 0; i < count; i++) {
    load_ext_limbs51(p, limbs + 20 * i);
    ge_add(acc, acc, p);
  }
  uint32_t enc[8];
  ristretto_encode(enc, acc);
  uint32_t z = 0;
  for (int i = 0; i < 8; i++) { res->enc[i] = enc[i]; z |= enc[i]; }
  res->status = 0;
  res->is_identity = (z == 0) ? 1 : 0;
  res->first_bad = -1;
}

// n == 0: the empty sum
__global__ void k_empty_result(msm_result* res) {
  for (int i = 0; i < 8; i++) res->enc[i] = 0;
  res->status = 0;
  res->is_identity = 1;
  res->first_bad = -1;
}

__global__ void k_init_flags(int* flags) {
  flags[0] = 0x7fffffff;
  flags[1] = 0x7fffffff;
}

}  // namespace zkp
