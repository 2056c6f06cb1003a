// Host stand-ins for the CUDA built-ins the kernel headers use, so that the kernels can be compiled with g++ and run
// on the CPU.  TEST INFRASTRUCTURE ONLY: it checks indexing, layouts and the arithmetic logic of the kernels against
// the oracle; it is not a fallback and nothing in zkp_b200/ includes it.
//   emul_launch      kernels WITHOUT inter-thread communication: every thread of every block in turn, on this thread
//   emul_launch_mt   kernels that use __syncthreads / warp shuffles / shared memory: one block at a time, one OS thread
//                    per CUDA thread, __syncthreads and shuffles implemented with std::barrier
#pragma once
#include <cuda_runtime.h>   // uint4, make_uint4, dim3 (host-side headers of the toolkit)
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __global__
#define __global__
#define __grid_constant__
#undef __device__
#define __device__
#undef __shared__
#define __shared__ static
#undef __forceinline__
#define __forceinline__ inline
#include <algorithm>
#include <barrier>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>
using std::max;
using std::min;

static thread_local uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;

// ---- block context of emul_launch_mt -------------------------------------------------------------------------------
// Barriers are std::barrier (futex based: no mutex to fight over when a thousand threads wake up).  A thread that
// returns drops out of its block's and its warp's barrier.  A shuffle is: publish, warp barrier, read, warp barrier --
// all live lanes of a warp execute the same shuffle (the kernels' shuffles are convergent), whatever the mask says.
struct emul_warp {
  std::barrier<> bar;
  unsigned long long slot[32];
  explicit emul_warp(unsigned lanes) : bar((std::ptrdiff_t)lanes) {}
};
struct emul_block {
  std::barrier<> bar;
  std::vector<std::unique_ptr<emul_warp>> warps;
  explicit emul_block(unsigned threads) : bar((std::ptrdiff_t)threads) {
    for (unsigned w = 0; w * 32 < threads; w++)
      warps.emplace_back(new emul_warp(threads - w * 32 < 32 ? threads - w * 32 : 32));
  }
  void syncthreads() { bar.arrive_and_wait(); }
  void thread_exit(unsigned tid) {
    warps[tid >> 5]->bar.arrive_and_drop();
    bar.arrive_and_drop();
  }
};
static emul_block* g_emul_block = nullptr;   // non-null while emul_launch_mt runs a block

template <class T> static inline T __ldg(const T* p) { return *p; }
// atomics: real ones (several OS threads under emul_launch_mt)
static inline int atomicMin(int* a, int v) {
  int o = __atomic_load_n(a, __ATOMIC_RELAXED);
  while (v < o && !__atomic_compare_exchange_n(a, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return o;
}
static inline unsigned atomicAdd(unsigned* a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
static inline int atomicExch(int* a, int v) { return __atomic_exchange_n(a, v, __ATOMIC_RELAXED); }
static inline int atomicOr(int* a, int v) { return __atomic_fetch_or(a, v, __ATOMIC_RELAXED); }
static inline unsigned atomicOr(unsigned* a, unsigned v) { return __atomic_fetch_or(a, v, __ATOMIC_RELAXED); }

static inline int __double2loint(double d) { unsigned long long b; memcpy(&b, &d, 8); return (int)(unsigned)b; }
static inline int __double2hiint(double d) { unsigned long long b; memcpy(&b, &d, 8); return (int)(unsigned)(b >> 32); }
static inline double __hiloint2double(int hi, int lo) {
  unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
  double d; memcpy(&d, &b, 8); return d;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }

// shuffles: under emul_launch_mt the lanes named in `mask` exchange through the warp's slots; under emul_launch a thread
// is alone and reads its own value
template <class T>
static inline T emul_shfl(unsigned, T v, int src_lane, bool src_valid) {
  if (!g_emul_block) return v;
  static_assert(sizeof(T) <= 8, "shuffle of a wide type");
  emul_warp& w = *g_emul_block->warps[threadIdx.x >> 5];
  const int lane = (int)(threadIdx.x & 31);
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  w.slot[lane] = bits;
  w.bar.arrive_and_wait();
  T r = v;
  if (src_valid) memcpy(&r, &w.slot[src_lane], sizeof(T));
  w.bar.arrive_and_wait();
  return r;
}
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  const int lane = (int)(threadIdx.x & 31);
  return emul_shfl(mask, v, (lane & ~(width - 1)) | (src & (width - 1)), true);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, int d, int width = 32) {
  const int lane = (int)(threadIdx.x & 31), src = lane - d;
  return emul_shfl(mask, v, src, src >= (lane & ~(width - 1)));
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, int d, int width = 32) {
  const int lane = (int)(threadIdx.x & 31), src = lane + d;
  return emul_shfl(mask, v, src, src <= (lane | (width - 1)));
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
  const int lane = (int)(threadIdx.x & 31);
  return emul_shfl(mask, v, lane ^ x, true);
}
static inline void __syncthreads() {
  if (g_emul_block) g_emul_block->syncthreads();
}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
// warp votes / matches: a thread is its own (and only) peer -- equivalent for the kernels here, which use them to
// aggregate atomics over equal keys
static inline unsigned __activemask() { return 1u << (threadIdx.x & 31); }
template <class T> static inline unsigned __match_any_sync(unsigned, T) { return 1u << (threadIdx.x & 31); }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u << (threadIdx.x & 31) : 0u; }
static inline int __any_sync(unsigned, int p) { return p; }

// run a communication-free kernel: every thread of every block in turn
template <class K, class... A>
static void emul_launch(dim3 grid, unsigned block, K kernel, A... args) {
  gridDim = grid;
  blockDim = dim3(block, 1, 1);
  for (unsigned by = 0; by < grid.y; by++)
    for (unsigned b = 0; b < grid.x; b++)
      for (unsigned t = 0; t < block; t++) {
        blockIdx.x = b; blockIdx.y = by; blockIdx.z = 0;
        threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
        kernel(args...);
      }
}

// run a kernel whose threads cooperate: block after block, one OS thread per CUDA thread
template <class K, class... A>
static void emul_launch_mt(unsigned grid, unsigned block, K kernel, A... args) {
  gridDim = dim3(grid, 1, 1);
  blockDim = dim3(block, 1, 1);
  for (unsigned b = 0; b < grid; b++) {
    emul_block ctx(block);
    g_emul_block = &ctx;
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; t++)
      th.emplace_back([&, t]() {
        blockIdx.x = b; blockIdx.y = 0; blockIdx.z = 0;
        threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
        kernel(args...);
        ctx.thread_exit(t);
      });
    for (auto& x : th) x.join();
    g_emul_block = nullptr;
  }
}
