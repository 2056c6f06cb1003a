// Host stand-ins for the CUDA built-ins the kernel headers use, so that kernels WITHOUT inter-thread communication
// (no shared memory exchange, shuffles or barriers on their data path) can be compiled with g++ and run thread by
// thread on the CPU (emul_launch).  TEST INFRASTRUCTURE ONLY: it checks indexing, layouts and the arithmetic logic of
// such kernels against the oracle; it is not a fallback and nothing in zkp_b200/ includes it.
#pragma once
#include <cuda_runtime.h>   // uint4, make_uint4, dim3 (host-side headers of the toolkit)
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __global__
#define __global__
#undef __device__
#define __device__
#undef __shared__
#define __shared__ static
#undef __forceinline__
#define __forceinline__ inline
#include <algorithm>
#include <cstring>
using std::max;
using std::min;

static uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;

template <class T> static inline T __ldg(const T* p) { return *p; }
// single-threaded emulation: "atomics" are plain read-modify-writes
static inline int atomicMin(int* a, int v) { int o = *a; if (v < o) *a = v; return o; }
static inline unsigned atomicAdd(unsigned* a, unsigned v) { unsigned o = *a; *a += v; return o; }
static inline int atomicExch(int* a, int v) { int o = *a; *a = v; return o; }
static inline int atomicOr(int* a, int v) { int o = *a; *a |= v; return o; }
static inline unsigned atomicOr(unsigned* a, unsigned v) { unsigned o = *a; *a |= v; return o; }
// warp / block primitives: present so the headers compile; kernels that rely on them must not be run through emul_launch
template <class T> static inline T __shfl_up_sync(unsigned, T v, int, int = 32) { return v; }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int, int = 32) { return v; }
template <class T> static inline T __shfl_sync(unsigned, T v, int, int = 32) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int, int = 32) { return v; }
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline int __any_sync(unsigned, int p) { return p; }
// a thread is alone in its warp here: it is its own (and only) peer
static inline unsigned __activemask() { return 1u << (threadIdx.x & 31); }
template <class T> static inline unsigned __match_any_sync(unsigned, T) { return 1u << (threadIdx.x & 31); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }

// run a communication-free kernel: every thread of every block in turn
template <class K, class... A>
static void emul_launch(unsigned grid, unsigned block, K kernel, A... args) {
  gridDim = dim3(grid, 1, 1);
  blockDim = dim3(block, 1, 1);
  for (unsigned b = 0; b < grid; b++)
    for (unsigned t = 0; t < block; t++) {
      blockIdx.x = b; blockIdx.y = 0; blockIdx.z = 0;
      threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
      kernel(args...);
    }
}
