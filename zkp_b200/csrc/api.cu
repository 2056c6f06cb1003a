// C ABI of the engine (include/zkp_b200.h): context, workspace, launch sequencing.
// Everything that computes is in kernels.cuh / small_msm.cuh; this file only allocates, copies and launches.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <sys/random.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "../../include/zkp_b200.h"
#include "host/merlin.hpp"
#include "kernels.cuh"
#include "small_msm.cuh"
#include "bench_fe.cuh"
#include "bv_kernels.cuh"
#include "pv_kernels.cuh"
#include "pv_plan.hpp"
#include "bv_plan.hpp"

using namespace zkp;

// Paths that were measured and lost stay in the source as ablations but out of the product
// library: built only with -DZKP_ABLATIONS (the host-emulation tests do; `python -m zkp_b200.build --ablations`).
//   overlap        digit sort on a second stream beside the decompression      (159.2 vs 159.9 ms per step: neutral)
//   ramp_chunks    ramped H2D chunk schedule                                   (138.6 vs 135.8 ms per step: slower)
//   bv_compiled=0  byte-wise STROBE front end k_bv_prepare                     (~40 vs 18 ms per 2^21 proofs)
//   k_ingest       single-phase fused ingestion (histogram only)               (superseded by k_ingest2)
//   fe64 / bench   FP64-pipe field arithmetic and the mixed-pipe micro-benchmarks (no gain: section 3)
#ifdef ZKP_ABLATIONS
#define ZKP_ABL(x) (x)
#else
#define ZKP_ABL(x) (0)
#endif

struct devbuf {
  void* p = nullptr;
  size_t cap = 0;
};

// workspace + stream of one slice of the batch prover (two of them alternate: zkp_prove_batch)
struct pv_set {
  devbuf pv_misc, pv_limbs, pv_enc, pv_sec, pv_ent, pv_state, pv_blind, pv_resp, in_scalars, niels, tables, sk0, aux0, aux1,
      multi, flags, pv_static;
  cudaStream_t stream = nullptr;       // set 0: the context's stream; set 1: own_stream
  cudaStream_t own_stream = nullptr;
  int* h_flags = nullptr;              // pinned, 64 bytes
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

static bool os_random_bytes(uint8_t* out, size_t n) {
  while (n) {
    ssize_t got = getrandom(out, n > 256 ? 256 : n, 0);
    if (got <= 0) return false;
    out += got;
    n -= (size_t)got;
  }
  return true;
}

struct zkp_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  int window = 0;      // 0 = auto
  int lanes = 0;       // 0 = auto
  int window_cap = 19; // upper bound for the automatic choice (measured optimum at 5*10^7 terms: 19 ~ 20 < 18 < 17)
  uint64_t launches = 0;
  std::string err;
  // workspace (grown on demand, reused across calls)
  devbuf in_scalars, in_points, niels, hist, offs, cursor, sorted, buckets, lvlT[2], lvlU, usum, flags, result,
      aux0, aux1, aux2, sk0, sk1, tables, items, partials, multi, bv_com, bv_resp, bv_part, bv_misc, len_hist, order,
      pv_limbs, pv_enc, pv_sec, pv_ent, pv_state, pv_blind, pv_resp, pv_misc, scan_tmp, partial_buf, pv_static;
  void* h_result = nullptr;  // pinned, 64 bytes
  // optional per-stage timing of the vartime MSM ("profile" option): events around each stage
  int profile = 0;
  cudaEvent_t ev[10] = {};
  float stage_ms[9] = {};
  // always-on timing of the two ingestion launches of the device-resident fused path (stages 7, 8) and of the
  // bucket accumulation kernel (stage 9): three event pairs per call, read back lazily by zkp_ctx_stage_ms
  cudaEvent_t ev_live[6] = {};
  bool live_valid = false;
  int last_window = 0, last_lanes = 0;
  // host-input pipeline: copies on copy_stream overlap the per-chunk kernels on `stream`
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> chunk_ev;
  // terms per H2D chunk of the host-input MSM.  Phase 1 of the ingestion is copy-bound, so a step ends one chunk's copy
  // plus one chunk's kernel after the ideal: small chunks, with their kernels alternating between two streams so that the
  // per-launch partial waves overlap.  Measured at the bench size (e2e per step): 2^21 on one stream 135.8 ms,
  // 2^20 on two 132.3 ms, 2^19 on two 130.6 ms (device-resident inputs: 129.5 ms)
  // (round 2, with the faster ingestion kernel and 54 % of the points in phase 1: 2^19 125.5 ms, 2^18 124.2 ms, 2^17 125.3-125.8,
  // 2^16 137.8 ms)
  size_t chunk_terms = (size_t)1 << 18;
  // the same for the slabs of zkp_batch_verify_proofs (x 4 / rows proofs per slab: 65 536 CMZ proofs); measured at HEAD: 2^18
  // 140.9 ms, 3 * 2^17 139.5, 2^19 140.3, 3 * 2^18 142.2 ms per 2^21 CMZ proofs
  size_t bv_chunk_terms = (size_t)3 << 17;
  // the digit sort (histogram, scan, scatter: L2-atomic bound) runs on a second, higher-priority stream
  // concurrently with decompression (integer-multiply bound); joined before bucket accumulation
  int overlap = 0;   // measured: no gain on B200 (159.2 vs 159.9 ms per step), kept as an option
  int balance = 1;     // size-ordered work items in the bucket accumulation (equal-length items share a warp)
  void* partial_out = nullptr;   // when set, k_finish also stores the MSM result as an extended point (single-verdict mode)
  size_t prove_chunk = (size_t)1 << 17;   // proofs per slice of zkp_prove_batch at most (workspace bound)
  // batches of at least two such slices are pipelined: slices alternate between two buffer sets / streams, so the copies
  // of one slice run under the kernels of the other (0 = off)
  size_t prove_pipe_chunk = (size_t)1 << 14;
  pv_set pvs[2];
  int coop_max_msms = 8192;  // batched small vartime MSMs: up to this many run with four lanes per MSM (latency)
  // share of the host-path point chunks decompressed under the histogram (first phase).  Phase 1 also moves all the scalars, so
  // at 55 GB/s it is copy-bound: a point more in it costs 0.58 ns of copy and saves 1.67 ns of phase 2, until the kernels of
  // phase 1 catch up with its copies.  Measured e2e at the bench size: 50 % 125.9 ms, 52 % 125.5, 54 % 125.1, 56 % 125.9, 58 % 126.6
  int phase1_percent = 54;
  int share_static_tables = 1;   // batch proving: one constant-time table per batch-static point
  // batch proving: signed four-tooth combs, one per base (comb.cuh): 64 doublings per constraint MSM instead of 256.
  // 0 = Straus tables (k_small_msm_ct), 1 = combs scanned from global memory (k_small_msm_comb), 2 = combs staged in
  // shared memory by one CTA per 32 proofs (k_comb_msm_cta; falls back to 1 when a group's combs do not fit one SM).
  // Measured, 2^16 CMZ proofs from pinned buffers: 47.2 ms (0), 35.0 ms (1) -- profiles/configs_r02_s1_*.json
  int prove_comb = 2;
  // one small MSM (zkp_msm_vartime / zkp_batch_verify / zkp_msm_vartime_dev with few terms): up to this many terms skip the
  // sort pipeline and run as groups of four lanes over the terms (k_single_msm_vt) -- the reference's own size dispatch
  // (n < 190 -> Straus, /root/reference/src/toolbox/verifier.rs:162-166 -> dalek edwards.rs [ext]); 0 = always Pippenger
  int small_max = 4096;
  int small_groups = 4096;    // at most this many four-lane groups share the terms (measured: profiles/r02_small_msm.json)
  int prove_piece = 2;        // CTA-staged comb kernel (prove_comb = 2): terms per unit (pv_make_units)
  int smem_optin = -1;        // cudaDevAttrMaxSharedMemoryPerBlockOptin, read once
  int sm_count = -1;          // cudaDevAttrMultiProcessorCount, read once
  size_t cta_smem_set = 0;    // dynamic shared memory k_comb_msm_cta has been allowed so far
  // batch verification from proofs (zkp_batch_verify_proofs)
  int bv_phase1_rows = 0;   // rows of a slab decompressed in phase 1: 0 = chosen per slab (half, or all while the host link is the limit)
  int bv_prep_stream = 1;   // front-end kernel of slab i + 1 on its own high-priority stream next to the decompression of slab i
  int bv_prep_smem_kb = 64; // ... its residency cap: unused dynamic shared memory per 128-thread block (64 KB = 2 blocks per SM)
  size_t bv_prep_smem_set = 0;
  int bv_prep_blocks = 2;   // ... and the resident grid it runs as: blocks per SM (0 = one block per 128 proofs)
  cudaStream_t prep_stream = nullptr;
  std::vector<cudaEvent_t> prep_ev, ing_ev;
  int bv_compiled = 1;      // batch-verification front end: host-compiled transcript script (k_bv_prepare2)
  int ingest_variant = 4;   // occupancy point of k_ingest2 (kernels.cuh ZKP_INGEST_*): 4 = 72 registers, 28 warps/SM; 2 = 80, 24
  int accumulate_variant = 5;   // occupancy point of k_accumulate: 5 = 96 registers, 20 warps/SM, no prefetch; 4 = 122, 16, prefetch
  int scatter_batch = 1;    // k_ingest2, scatter phase: cursor atomics of four windows in flight together
  int fused_sort = 1;  // histogram and scatter ride under the two halves of the decompression (k_ingest2)
  // host-input pipeline: ramped chunk sizes (from chunk_terms / 8 up to 2 * chunk_terms, down again at the end of phase 1).
  // Measured at the bench size: 138.6 ms per step against 135.8 ms with uniform chunks -- phase 1 is copy-bound, so every
  // chunk larger than its predecessor leaves the SMs idle while it arrives.  Off; kept as an option.
  int ramp_chunks = 0;
  // host-input pipeline: the chunk kernels alternate between two streams, so the last partial wave of one chunk's kernel
  // overlaps the first blocks of the next (the digit work is order-independent); joined at the scan and at the end
  int dual_stream = 1;
  cudaStream_t aux_stream = nullptr;
  cudaStream_t sort_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

#define CUDA_TRY(ctx, call)                                                            \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
      return (e_ == cudaErrorMemoryAllocation) ? ZKP_ERR_NOMEM : ZKP_ERR_CUDA;         \
    }                                                                                  \
  } while (0)

static int32_t ensure(zkp_ctx* ctx, devbuf& b, size_t bytes) {
  if (bytes <= b.cap) return ZKP_OK;
  if (b.p) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&b.p, want);
  }
  if (e != cudaSuccess) {
    b.p = nullptr;
    ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return ZKP_ERR_NOMEM;
  }
  b.cap = want;
  return ZKP_OK;
}
#define ENSURE(ctx, buf, bytes)                       \
  do {                                                \
    int32_t r_ = ensure((ctx), (buf), (bytes));       \
    if (r_ != ZKP_OK) return r_;                      \
  } while (0)

#define LAUNCH_CHECK(ctx)                                                  \
  do {                                                                     \
    (ctx)->launches++;                                                     \
    cudaError_t e_ = cudaGetLastError();                                   \
    if (e_ != cudaSuccess) {                                               \
      (ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(e_); \
      return ZKP_ERR_CUDA;                                                 \
    }                                                                      \
  } while (0)

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

extern "C" int32_t zkp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int32_t zkp_ctx_create(zkp_ctx** out, int32_t device) {
  if (!out) return ZKP_ERR_SIZE;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) {
    cudaGetLastError();
    return ZKP_ERR_NOGPU;
  }
  zkp_ctx* ctx = new (std::nothrow) zkp_ctx();
  if (!ctx) return ZKP_ERR_NOMEM;
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost(&ctx->h_result, 64) != cudaSuccess) {
    cudaGetLastError();
    delete ctx;
    return ZKP_ERR_NOGPU;
  }
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return ZKP_OK;
}

extern "C" void zkp_ctx_destroy(zkp_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  devbuf* bufs[] = {&ctx->in_scalars, &ctx->in_points, &ctx->niels, &ctx->hist, &ctx->offs, &ctx->cursor,
                    &ctx->sorted, &ctx->buckets, &ctx->lvlT[0], &ctx->lvlT[1], &ctx->lvlU, &ctx->usum,
                    &ctx->flags, &ctx->result, &ctx->aux0, &ctx->aux1, &ctx->aux2, &ctx->sk0, &ctx->sk1,
                    &ctx->tables, &ctx->items, &ctx->partials, &ctx->multi, &ctx->bv_com, &ctx->bv_resp, &ctx->bv_part,
                    &ctx->bv_misc, &ctx->len_hist, &ctx->order, &ctx->pv_limbs, &ctx->pv_enc, &ctx->pv_sec, &ctx->pv_ent,
                    &ctx->pv_state, &ctx->pv_blind, &ctx->pv_resp, &ctx->pv_misc, &ctx->scan_tmp, &ctx->partial_buf, &ctx->pv_static};
  for (devbuf* b : bufs)
    if (b->p) cudaFree(b->p);
  for (pv_set& S : ctx->pvs) {
    devbuf* sb[] = {&S.pv_misc, &S.pv_limbs, &S.pv_enc, &S.pv_sec, &S.pv_ent, &S.pv_state, &S.pv_blind, &S.pv_resp,
                    &S.in_scalars, &S.niels, &S.tables, &S.sk0, &S.aux0, &S.aux1, &S.multi, &S.flags, &S.pv_static};
    for (devbuf* b : sb)
      if (b->p) cudaFree(b->p);
    if (S.h_flags) cudaFreeHost(S.h_flags);
    if (S.ev_fork) cudaEventDestroy(S.ev_fork);
    if (S.ev_join) cudaEventDestroy(S.ev_join);
    if (S.own_stream) cudaStreamDestroy(S.own_stream);
  }
  if (ctx->h_result) cudaFreeHost(ctx->h_result);
  for (cudaEvent_t e : ctx->chunk_ev) cudaEventDestroy(e);
  for (int i = 0; i < 10; i++)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 6; i++)
    if (ctx->ev_live[i]) cudaEventDestroy(ctx->ev_live[i]);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->sort_stream) cudaStreamDestroy(ctx->sort_stream);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->prep_stream) cudaStreamDestroy(ctx->prep_stream);
  for (cudaEvent_t e : ctx->prep_ev) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->ing_ev) cudaEventDestroy(e);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

extern "C" int32_t zkp_ctx_set_stream(zkp_ctx* ctx, void* cuda_stream) {
  if (!ctx) return ZKP_ERR_SIZE;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return ZKP_OK;
}

extern "C" int32_t zkp_ctx_set_option(zkp_ctx* ctx, const char* key, int64_t value) {
  if (!ctx || !key) return ZKP_ERR_SIZE;
  if (!strcmp(key, "window")) {
    if (value != 0 && (value < 4 || value > 24)) return ZKP_ERR_SIZE;
    ctx->window = (int)value;
  } else if (!strcmp(key, "window_cap")) {
    if (value < 4 || value > 24) return ZKP_ERR_SIZE;
    ctx->window_cap = (int)value;
  } else if (!strcmp(key, "overlap")) {
    if (value && !ZKP_ABL(1)) return ZKP_ERR_SIZE;   // ablation: not in the product build
    ctx->overlap = value ? 1 : 0;
  } else if (!strcmp(key, "balance")) {
    ctx->balance = value ? 1 : 0;
  } else if (!strcmp(key, "coop_max_msms")) {
    if (value < 0) return ZKP_ERR_SIZE;
    ctx->coop_max_msms = (int)value;
  } else if (!strcmp(key, "phase1_percent")) {
    if (value < 1 || value > 100) return ZKP_ERR_SIZE;
    ctx->phase1_percent = (int)value;
  } else if (!strcmp(key, "prove_chunk")) {
    if (value < 1) return ZKP_ERR_SIZE;
    ctx->prove_chunk = (size_t)value;
  } else if (!strcmp(key, "prove_pipe_chunk")) {
    if (value < 0) return ZKP_ERR_SIZE;
    ctx->prove_pipe_chunk = (size_t)value;
  } else if (!strcmp(key, "share_static_tables")) {
    ctx->share_static_tables = value ? 1 : 0;
  } else if (!strcmp(key, "prove_comb")) {
    if (value < 0 || value > 2) return ZKP_ERR_SIZE;
    ctx->prove_comb = (int)value;
  } else if (!strcmp(key, "small_max")) {
    if (value < 0 || value > (1 << 20)) return ZKP_ERR_SIZE;
    ctx->small_max = (int)value;
  } else if (!strcmp(key, "small_groups")) {
    if (value < 1 || value > (1 << 16)) return ZKP_ERR_SIZE;
    ctx->small_groups = (int)value;
  } else if (!strcmp(key, "prove_piece")) {
    if (value < 1 || value > 64) return ZKP_ERR_SIZE;
    ctx->prove_piece = (int)value;
  } else if (!strcmp(key, "bv_phase1_rows")) {
    if (value < -1 || value > 4096) return ZKP_ERR_SIZE;
    ctx->bv_phase1_rows = (int)value;
  } else if (!strcmp(key, "bv_prep_stream")) {
    ctx->bv_prep_stream = value ? 1 : 0;
  } else if (!strcmp(key, "bv_prep_blocks")) {
    if (value < 0 || value > 16) return ZKP_ERR_SIZE;
    ctx->bv_prep_blocks = (int)value;
  } else if (!strcmp(key, "bv_prep_smem_kb")) {
    if (value < 0 || value > 160) return ZKP_ERR_SIZE;
    ctx->bv_prep_smem_kb = (int)value;
  } else if (!strcmp(key, "bv_compiled")) {
    if (!value && !ZKP_ABL(1)) return ZKP_ERR_SIZE;   // ablation: not in the product build
    ctx->bv_compiled = value ? 1 : 0;
  } else if (!strcmp(key, "ingest_variant")) {
    if (value < 0 || value > 5 || (!ZKP_ABL(1) && value != 2 && value != 4)) return ZKP_ERR_SIZE;
    ctx->ingest_variant = (int)value;
  } else if (!strcmp(key, "fused_sort")) {
    ctx->fused_sort = value ? 1 : 0;
  } else if (!strcmp(key, "ramp_chunks")) {
    if (value && !ZKP_ABL(1)) return ZKP_ERR_SIZE;   // ablation: not in the product build
    ctx->ramp_chunks = value ? 1 : 0;
  } else if (!strcmp(key, "dual_stream")) {
    ctx->dual_stream = value ? 1 : 0;
  } else if (!strcmp(key, "chunk_terms")) {
    if (value < 1024) return ZKP_ERR_SIZE;
    ctx->chunk_terms = (size_t)value;
  } else if (!strcmp(key, "bv_chunk_terms")) {
    if (value < 1024) return ZKP_ERR_SIZE;
    ctx->bv_chunk_terms = (size_t)value;
  } else if (!strcmp(key, "accumulate_variant")) {
    if (value < 4 || value > 5) return ZKP_ERR_SIZE;
    ctx->accumulate_variant = (int)value;
  } else if (!strcmp(key, "scatter_batch")) {
    ctx->scatter_batch = value ? 1 : 0;
  } else if (!strcmp(key, "profile")) {
    ctx->profile = value ? 1 : 0;
    if (ctx->profile && !ctx->ev[0])
      for (int i = 0; i < 10; i++)
        if (cudaEventCreate(&ctx->ev[i]) != cudaSuccess) return ZKP_ERR_CUDA;
  } else if (!strcmp(key, "lanes") || !strcmp(key, "chunk")) {
    if (value < 0 || value > (1 << 20)) return ZKP_ERR_SIZE;
    ctx->lanes = (int)value;
  } else {
    return ZKP_ERR_SIZE;
  }
  return ZKP_OK;
}

extern "C" int32_t zkp_ctx_synchronize(zkp_ctx* ctx) {
  if (!ctx) return ZKP_ERR_SIZE;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKP_OK;
}

// diagnostic: stage 0..6 = decompress, recode+hist, scan, scatter, accumulate, bucket-reduce, finish;
// 100 = window width used, 101 = lanes per bucket used
extern "C" double zkp_ctx_stage_ms(zkp_ctx* ctx, int32_t stage) {
  if (!ctx) return -1.0;
  if (stage == 100) return ctx->last_window;
  if (stage == 101) return ctx->last_lanes;
  if (stage >= 7 && stage <= 9) {   // live events of the last fused device-resident call (the caller has synchronised)
    if (!ctx->live_valid) return -1.0;
    float ms = 0;
    const int i = 2 * (stage - 7);
    if (cudaEventElapsedTime(&ms, ctx->ev_live[i], ctx->ev_live[i + 1]) != cudaSuccess) {
      cudaGetLastError();
      return -1.0;
    }
    return ms;
  }
  if (stage < 0 || stage > 6) return -1.0;
  return ctx->stage_ms[stage];
}

extern "C" const char* zkp_last_error(zkp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" uint64_t zkp_ctx_launch_count(zkp_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------------------------
// window / lane heuristics
// ---------------------------------------------------------------------------------------------------------
static int choose_window(const zkp_ctx* ctx, size_t n) {
  if (ctx->window) return ctx->window;
  int best = 4;
  double best_cost = 1e300;
  for (int c = 4; c <= ctx->window_cap; c++) {
    double W = (253 + c - 1) / c;
    double B = (double)(1ull << (c - 1));
    double cost = W * ((double)n + 3.0 * B);  // bucket adds + ~2 full adds per bucket in the reduction
    if (cost < best_cost) {
      best_cost = cost;
      best = c;
    }
  }
  return best;
}

// ---------------------------------------------------------------------------------------------------------
// the variable-time MSM, device-resident and asynchronous
// ---------------------------------------------------------------------------------------------------------
struct msm_plan {
  int c, W;
  uint32_t B, total_buckets;
  cudaStream_t sort;   // stream of the digit sort: ctx->sort_stream when overlapping, else ctx->stream
};

// phase A: choose the window, size the workspace, reset flags and histogram (on the compute stream)
static int32_t msm_prepare(zkp_ctx* ctx, size_t n, msm_plan* pl) {
  cudaStream_t st = ctx->stream;
  if (n >= 0x7fffffffull) {
    ctx->err = "n too large (>= 2^31 terms per call)";
    return ZKP_ERR_SIZE;
  }
  ctx->live_valid = false;
  const int c = choose_window(ctx, n);
  const int W = (253 + c - 1) / c;
  const uint32_t B = 1u << (c - 1);
  const uint32_t total_buckets = (uint32_t)W * B;
  pl->c = c; pl->W = W; pl->B = B; pl->total_buckets = total_buckets;
  ENSURE(ctx, ctx->niels, n * (16 * ZKP_NIELS_U4));
  ENSURE(ctx, ctx->hist, (size_t)W * B * 4);
  ENSURE(ctx, ctx->offs, (size_t)W * (B + 1) * 4);
  ENSURE(ctx, ctx->cursor, (size_t)W * B * 4);
  ENSURE(ctx, ctx->sorted, (size_t)W * n * 4);
  ENSURE(ctx, ctx->buckets, (size_t)total_buckets * 128);
  ENSURE(ctx, ctx->flags, 16);
  const uint32_t chunks0 = B / ZKP_CHUNK_L;
  if (chunks0 > 0) {
    ENSURE(ctx, ctx->lvlT[0], (size_t)W * chunks0 * 128);
    ENSURE(ctx, ctx->lvlT[1], (size_t)W * (chunks0 / ZKP_CHUNK_L + 1) * 128);
    ENSURE(ctx, ctx->lvlU, (size_t)W * chunks0 * 128);
  }
  ENSURE(ctx, ctx->usum, (size_t)8 * W * 128);
  ctx->last_window = c;
  k_init_flags<<<1, 1, 0, st>>>((int*)ctx->flags.p);
  LAUNCH_CHECK(ctx);
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->hist.p, 0, (size_t)W * B * 4, st));
  pl->sort = st;
  if (ZKP_ABL(ctx->overlap) && !ctx->profile) {
    if (!ctx->sort_stream) {
      int lo = 0, hi = 0;
      CUDA_TRY(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->sort_stream, cudaStreamNonBlocking, hi));
    }
    if (!ctx->ev_fork) {
      CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
      CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, st));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->sort_stream, ctx->ev_fork, 0));
    pl->sort = ctx->sort_stream;
  }
  return ZKP_OK;
}

#define STAGE(i) do { if (ctx->profile) cudaEventRecord(ctx->ev[i], st); } while (0)

// phase B: ingest terms [base, base+cnt): decompress the points, histogram the scalars' digits
static int32_t msm_ingest(zkp_ctx* ctx, const msm_plan& pl, const void* d_scalars, const void* d_points, size_t base,
                          size_t cnt, bool whole) {
  cudaStream_t st = ctx->stream;
  int* flags = (int*)ctx->flags.p;
  const unsigned nb = (unsigned)((cnt + 255) / 256);
#ifdef ZKP_ABLATIONS
  if (!ctx->profile && pl.sort == st) {   // fused: histogram reductions ride along with the decompression
    k_ingest<<<nb, 256, 0, st>>>((const uint4*)d_points + 2 * base, (const uint4*)d_scalars + 2 * base, cnt,
                                 (uint4*)ctx->niels.p + ZKP_NIELS_U4 * base, pl.c, pl.W, pl.B, (uint32_t*)ctx->hist.p, flags, base);
    LAUNCH_CHECK(ctx);
    return ZKP_OK;
  }
#endif
  if (whole) STAGE(0);
  k_decompress<<<nb, 256, 0, st>>>((const uint4*)d_points + 2 * base, cnt, (uint4*)ctx->niels.p + ZKP_NIELS_U4 * base, flags, base);
  LAUNCH_CHECK(ctx);
  if (whole) STAGE(1);
  k_recode<false><<<nb, 256, 0, pl.sort>>>((const uint4*)d_scalars + 2 * base, cnt, pl.c, pl.W, pl.B,
                                           (uint32_t*)ctx->hist.p, nullptr, flags, base);
  LAUNCH_CHECK(ctx);
  if (whole) STAGE(2);
  return ZKP_OK;
}

static int32_t msm_finish(zkp_ctx* ctx, const msm_plan& pl, const void* d_scalars, size_t n, msm_result* d_result,
                          bool whole, bool sorted_done = false);

// one launch of the two-phase ingestion (kernels.cuh k_ingest2): the points and the (up to three) term ranges of `a`,
// ny equally shaped launches as the y dimension of one grid; MODE 0 histograms the terms' digits, MODE 1 scatters them
template <int MODE>
static int32_t launch_ingest2(zkp_ctx* ctx, const msm_plan& pl, const void* d_scalars, const void* d_points, size_t n,
                              const ingest_args& a_in, unsigned ny = 1, cudaStream_t on = nullptr) {
  const cudaStream_t launch_stream = on ? on : ctx->stream;
  ingest_args a = a_in;
  a.batched = ctx->scatter_batch;
  size_t threads = a.p_cnt;
  for (int j = 0; j < 3; j++)
    if (a.s_cnt[j] > threads) threads = a.s_cnt[j];
  if (!threads || !ny) return ZKP_OK;
#define ZKP_LAUNCH_INGEST(VAR)                                                                                              \
  k_ingest2<MODE, VAR><<<dim3((unsigned)((threads + ZKP_INGEST_THREADS(VAR) - 1) / ZKP_INGEST_THREADS(VAR)), ny),           \
                         ZKP_INGEST_THREADS(VAR), 0, launch_stream>>>(                                                      \
      (const uint4*)d_points, (uint4*)ctx->niels.p, (const uint4*)d_scalars, a, n, pl.c, pl.W, pl.B,                        \
      MODE == 0 ? (uint32_t*)ctx->hist.p : (uint32_t*)ctx->cursor.p, (uint32_t*)ctx->sorted.p, (int*)ctx->flags.p)
  switch (ctx->ingest_variant) {
    case 2: ZKP_LAUNCH_INGEST(2); break;
#ifdef ZKP_ABLATIONS   // the other occupancy points (measured: DESIGN.md section 4)
    case 0: ZKP_LAUNCH_INGEST(0); break;
    case 1: ZKP_LAUNCH_INGEST(1); break;
    case 3: ZKP_LAUNCH_INGEST(3); break;
    case 5: ZKP_LAUNCH_INGEST(5); break;
#endif
    default: ZKP_LAUNCH_INGEST(4); break;
  }
#undef ZKP_LAUNCH_INGEST
  LAUNCH_CHECK(ctx);
  return ZKP_OK;
}
// points [p_lo, p_lo + p_cnt) and ONE contiguous term range [s_lo, s_lo + s_cnt) cut into two halves
template <int MODE>
static int32_t launch_ingest2_range(zkp_ctx* ctx, const msm_plan& pl, const void* d_scalars, const void* d_points, size_t n,
                                    size_t p_lo, size_t p_cnt, size_t s_lo, size_t s_cnt, cudaStream_t on = nullptr) {
  const size_t half = (s_cnt + 1) / 2;
  ingest_args a = {};
  a.p_lo = p_lo; a.p_cnt = p_cnt;
  a.s_lo[0] = s_lo; a.s_cnt[0] = half;
  a.s_lo[1] = s_lo + half; a.s_cnt[1] = s_cnt - half;
  return launch_ingest2<MODE>(ctx, pl, d_scalars, d_points, n, a, 1, on);
}
static bool use_fused_sort(const zkp_ctx* ctx, const msm_plan& pl) {
  return !ctx->profile && pl.sort == ctx->stream && ctx->fused_sort;
}

// few terms: decompress, fold the scalars, groups of four lanes over the terms, one block adds and encodes (4 launches)
static bool use_small_path(const zkp_ctx* ctx, size_t n) {
  return n > 0 && n <= (size_t)ctx->small_max && !ctx->profile && !ctx->window;   // a forced window asks for the pipeline
}
static int32_t msm_small_launch(zkp_ctx* ctx, const void* d_scalars, const void* d_points, size_t n, msm_result* d_result) {
  cudaStream_t st = ctx->stream;
  const uint32_t G = (uint32_t)(n < (size_t)ctx->small_groups ? n : (size_t)ctx->small_groups);
  ENSURE(ctx, ctx->niels, (n + 1) * (16 * ZKP_NIELS_U4));
  ENSURE(ctx, ctx->sk0, n * 32 + 32);
  ENSURE(ctx, ctx->sk1, n * 32 + 32);
  ENSURE(ctx, ctx->partials, (size_t)G * 128);
  ENSURE(ctx, ctx->flags, 16);
  ctx->live_valid = false;
  ctx->last_window = 0;
  ctx->last_lanes = 0;
  k_init_flags<<<1, 1, 0, st>>>((int*)ctx->flags.p);
  LAUNCH_CHECK(ctx);
  const unsigned nb = (unsigned)((n + 255) / 256);
  k_decompress_valid<<<nb, 256, 0, st>>>((const uint4*)d_points, n, (uint4*)ctx->niels.p);
  LAUNCH_CHECK(ctx);
  k_prep_scalars_vt<<<nb, 256, 0, st>>>((const uint4*)d_scalars, n, (uint4*)ctx->sk0.p, (uint4*)ctx->sk1.p);
  LAUNCH_CHECK(ctx);
  k_single_msm_vt<<<(4 * G + 63) / 64, 64, 0, st>>>((const uint32_t*)ctx->sk0.p, (const uint32_t*)ctx->sk1.p,
                                                    (const uint4*)ctx->niels.p, n, G, (uint4*)ctx->partials.p,
                                                    (int*)ctx->flags.p);
  LAUNCH_CHECK(ctx);
  k_single_finish<<<1, 128, 0, st>>>((const uint4*)ctx->partials.p, G, (const int*)ctx->flags.p, d_result,
                                     (uint4*)ctx->partial_out);
  LAUNCH_CHECK(ctx);
  return ZKP_OK;
}

static int32_t msm_vartime_launch(zkp_ctx* ctx, const void* d_scalars, const void* d_points, size_t n,
                                  msm_result* d_result) {
  cudaStream_t st = ctx->stream;
  if (n == 0) {
    k_empty_result<<<1, 1, 0, st>>>(d_result);
    LAUNCH_CHECK(ctx);
    return ZKP_OK;
  }
  if (use_small_path(ctx, n)) return msm_small_launch(ctx, d_scalars, d_points, n, d_result);
  msm_plan pl;
  int32_t r = msm_prepare(ctx, n, &pl);
  if (r != ZKP_OK) return r;
  if (use_fused_sort(ctx, pl)) {
    // phase 1: first half of the points + histogram of all scalars; scan; phase 2: second half + scatter of all scalars
    const size_t half_p = (n + 1) / 2;
    if (!ctx->ev_live[0])
      for (int i = 0; i < 6; i++) CUDA_TRY(ctx, cudaEventCreate(&ctx->ev_live[i]));
    ctx->live_valid = false;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_live[0], st));
    r = launch_ingest2_range<0>(ctx, pl, d_scalars, d_points, n, 0, half_p, 0, n);
    if (r != ZKP_OK) return r;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_live[1], st));
    k_scan<<<pl.W, 1024, 0, st>>>((const uint32_t*)ctx->hist.p, pl.B, (uint32_t*)ctx->offs.p, (uint32_t*)ctx->cursor.p);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_live[2], st));
    r = launch_ingest2_range<1>(ctx, pl, d_scalars, d_points, n, half_p, n - half_p, 0, n);
    if (r != ZKP_OK) return r;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_live[3], st));
    ctx->live_valid = true;
    return msm_finish(ctx, pl, d_scalars, n, d_result, true, true);
  }
  r = msm_ingest(ctx, pl, d_scalars, d_points, 0, n, true);
  if (r != ZKP_OK) return r;
  return msm_finish(ctx, pl, d_scalars, n, d_result, true);
}

// phase C: counting sort by bucket, bucket accumulation, bucket reduction, Horner, encode
static int32_t msm_finish(zkp_ctx* ctx, const msm_plan& pl, const void* d_scalars, size_t n, msm_result* d_result,
                          bool whole, bool sorted_done) {
  cudaStream_t st = ctx->stream;
  const int c = pl.c, W = pl.W;
  const uint32_t B = pl.B, total_buckets = pl.total_buckets;
  int* flags = (int*)ctx->flags.p;
  const unsigned nb = (unsigned)((n + 255) / 256);
  if (!sorted_done) {
    k_scan<<<W, 1024, 0, pl.sort>>>((const uint32_t*)ctx->hist.p, B, (uint32_t*)ctx->offs.p, (uint32_t*)ctx->cursor.p);
    LAUNCH_CHECK(ctx);
    if (whole) STAGE(3);
    k_recode<true><<<nb, 256, 0, pl.sort>>>((const uint4*)d_scalars, n, c, W, B, (uint32_t*)ctx->cursor.p,
                                            (uint32_t*)ctx->sorted.p, flags, 0);
    LAUNCH_CHECK(ctx);
  }
  if (pl.sort != st) {   // join: bucket accumulation needs both the Niels points and the sorted digits
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, pl.sort));
    CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
  }

  STAGE(4);
  {
    // chunk length: about 2x the mean bucket load, but small enough to leave >= ~200k work items
    const double mean = (double)n / (double)B;
    uint32_t S = ctx->lanes ? (uint32_t)ctx->lanes : 0;   // "lanes" option doubles as an explicit chunk length
    if (!S) {
      double want = 2.0 * mean, cap = (double)n * W / 200000.0;
      double v = want < cap ? want : cap;
      if (v < 16.0) v = 16.0;
      if (v > 4096.0) v = 4096.0;
      S = (uint32_t)v;
    }
    ctx->last_lanes = (int)S;
    const size_t max_items = (size_t)total_buckets + ((size_t)W * n) / S + 1;
    ENSURE(ctx, ctx->aux0, (size_t)total_buckets * 4);          // chunk counts
    ENSURE(ctx, ctx->aux1, ((size_t)total_buckets + 1) * 4);    // item offsets
    ENSURE(ctx, ctx->aux2, (size_t)total_buckets * 4);          // scan scratch (cursor output, unused)
    ENSURE(ctx, ctx->items, max_items * 16);
    ENSURE(ctx, ctx->partials, max_items * 128);
    const unsigned tb = (total_buckets + 255) / 256;
    k_plan<<<tb, 256, 0, st>>>((const uint32_t*)ctx->offs.p, B, total_buckets, S, (uint32_t*)ctx->aux0.p);
    LAUNCH_CHECK(ctx);
    if (total_buckets <= 65536) {
      k_scan<<<1, 1024, 0, st>>>((const uint32_t*)ctx->aux0.p, total_buckets, (uint32_t*)ctx->aux1.p,
                                 (uint32_t*)ctx->aux2.p);
      LAUNCH_CHECK(ctx);
    } else {
      // long arrays: tiles of 4096, scan of the tile totals (<= 2^20 tiles would still fit one block's loop), add back
      const uint32_t tiles = (total_buckets + ZKP_SCAN_TILE - 1) / ZKP_SCAN_TILE;
      ENSURE(ctx, ctx->scan_tmp, ((size_t)tiles * 3 + 8) * 4);
      uint32_t* tt = (uint32_t*)ctx->scan_tmp.p;            // totals [tiles] | offsets [tiles + 1] | cursor (unused) [tiles]
      k_scan_tiles<<<tiles, 1024, 0, st>>>((const uint32_t*)ctx->aux0.p, total_buckets, (uint32_t*)ctx->aux1.p, tt);
      LAUNCH_CHECK(ctx);
      k_scan<<<1, 1024, 0, st>>>(tt, tiles, tt + tiles, tt + 2 * tiles + 1);
      LAUNCH_CHECK(ctx);
      k_scan_add<<<(total_buckets + 1023) / 1024, 1024, 0, st>>>((uint32_t*)ctx->aux1.p, total_buckets, tt + tiles);
      LAUNCH_CHECK(ctx);
    }
    ENSURE(ctx, ctx->multi, ((size_t)total_buckets + 4) * 4);
    uint32_t* n_multi = (uint32_t*)ctx->multi.p;          // word 0 = counter, list starts at word 4
    CUDA_TRY(ctx, cudaMemsetAsync(n_multi, 0, 16, st));
    k_items<<<tb, 256, 0, st>>>((const uint32_t*)ctx->offs.p, (const uint32_t*)ctx->aux1.p, B, total_buckets, S,
                                (work_item*)ctx->items.p, n_multi + 4, n_multi);
    LAUNCH_CHECK(ctx);
    const unsigned blocks = (unsigned)((max_items + 127) / 128);
    const uint32_t* n_items = (const uint32_t*)ctx->aux1.p + total_buckets;
    const uint32_t* order = nullptr;
    if (ctx->balance && max_items >= 4096) {
      // counting sort of the item ids by length (longest first): equal-length items share a warp
      ENSURE(ctx, ctx->len_hist, (3 * (size_t)S + 8) * 4);                // histogram [S+1] | offsets [S+2] | cursor [S+1]
      ENSURE(ctx, ctx->order, max_items * 4);
      uint32_t* lh = (uint32_t*)ctx->len_hist.p;
      CUDA_TRY(ctx, cudaMemsetAsync(lh, 0, ((size_t)S + 1) * 4, st));
      const unsigned ib = (unsigned)((max_items + 255) / 256);
      k_len_hist<<<ib, 256, 0, st>>>((const work_item*)ctx->items.p, n_items, S, lh);
      LAUNCH_CHECK(ctx);
      k_scan<<<1, 1024, 0, st>>>(lh, S + 1, lh + S + 1, lh + 2 * S + 3);
      LAUNCH_CHECK(ctx);
      k_len_scatter<<<ib, 256, 0, st>>>((const work_item*)ctx->items.p, n_items, S, lh + 2 * S + 3, (uint32_t*)ctx->order.p);
      LAUNCH_CHECK(ctx);
      order = (const uint32_t*)ctx->order.p;
    }
    if (ctx->live_valid) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_live[4], st));
#define ZKP_LAUNCH_ACC(MB)                                                                                       \
  k_accumulate<MB><<<blocks, 128, 0, st>>>((const uint4*)ctx->niels.p, (const uint32_t*)ctx->sorted.p,           \
                                           (const work_item*)ctx->items.p, order, n_items, n, (uint4*)ctx->buckets.p, \
                                           (uint4*)ctx->partials.p)
    switch (ctx->accumulate_variant) {
      case 4: ZKP_LAUNCH_ACC(4); break;
      default: ZKP_LAUNCH_ACC(5); break;
    }
#undef ZKP_LAUNCH_ACC
    LAUNCH_CHECK(ctx);
    if (ctx->live_valid) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_live[5], st));
    k_merge<<<148 * 8, 128, 0, st>>>((const uint32_t*)ctx->aux1.p, n_multi + 4, n_multi,
                                     (const uint4*)ctx->partials.p, (uint4*)ctx->buckets.p);
    LAUNCH_CHECK(ctx);
  }

  STAGE(5);
  // bucket reduction levels
  const uint4* cur = (const uint4*)ctx->buckets.p;
  uint32_t m = B;
  int nl = 0;
  while (m > ZKP_CHUNK_L) {
    const uint32_t chunks = m / ZKP_CHUNK_L;
    uint4* T = (uint4*)ctx->lvlT[nl & 1].p;
    uint4* U = (uint4*)ctx->lvlU.p;
    const unsigned threads = chunks * (unsigned)W;
    k_chunk_reduce<<<(threads + 127) / 128, 128, 0, st>>>(cur, m, W, T, U);
    LAUNCH_CHECK(ctx);
    k_tree_sum<<<W, 256, 0, st>>>(U, chunks, (uint4*)ctx->usum.p + (size_t)nl * W * 8);
    LAUNCH_CHECK(ctx);
    cur = T;
    m = chunks;
    nl++;
  }
  STAGE(6);
  k_finish<<<1, 64, 0, st>>>((const uint4*)ctx->usum.p, nl, cur, m, W, c, n, flags, d_result, (uint4*)ctx->partial_out);
  LAUNCH_CHECK(ctx);
  STAGE(7);
  if (ctx->profile && whole) {
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev[7]));
    for (int i = 0; i < 7; i++) cudaEventElapsedTime(&ctx->stage_ms[i], ctx->ev[i], ctx->ev[i + 1]);
  }
#undef STAGE
  return ZKP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// host inputs: chunked H2D copies on a second stream overlap decompression / digit histograms of earlier chunks
// ---------------------------------------------------------------------------------------------------------
struct hseg {
  const uint8_t* p;
  size_t cnt;
};

static int32_t copy_range(zkp_ctx* ctx, const hseg* segs, int nseg, size_t lo, size_t hi, uint8_t* dst) {
  size_t off = 0;
  for (int k = 0; k < nseg; k++) {
    size_t s0 = off, s1 = off + segs[k].cnt;
    size_t a = lo > s0 ? lo : s0, b = hi < s1 ? hi : s1;
    if (a < b)
      CUDA_TRY(ctx, cudaMemcpyAsync(dst + a * 32, segs[k].p + (a - s0) * 32, (b - a) * 32, cudaMemcpyHostToDevice,
                                    ctx->copy_stream));
    off = s1;
  }
  return ZKP_OK;
}

// Chunk boundaries of one ingestion phase over the points [lo, hi): sizes double from `small` up to `big`, and with
// `mirror` they shrink again towards the end.  A small first chunk starts the kernels early; a small last chunk matters
// when the phase is copy-bound (phase 1 moves all the scalars: its H2D time equals its compute time at the bench size),
// because the kernel of the last chunk can only start when its last byte has arrived; large chunks in between keep the
// number of launches (one partial wave each) low.
static void phase_bounds(size_t lo, size_t hi, size_t small, size_t big, bool mirror, std::vector<size_t>* out) {
  const size_t len = hi - lo;
  if (!len) return;
  std::vector<size_t> up;
  const size_t target = mirror ? (len + 1) / 2 : len;
  size_t c = small, sum = 0;
  while (sum < target) {
    const size_t step = c < target - sum ? c : target - sum;
    up.push_back(step);
    sum += step;
    if (c < big) c = 2 * c < big ? 2 * c : big;
  }
  size_t pos = lo;
  for (size_t v : up) {
    pos += v;
    out->push_back(pos);
  }
  if (mirror) {
    size_t rest = len - sum;   // <= sum, short by at most one term: taken off the first mirrored chunk(s)
    for (size_t i = up.size(); i-- > 0 && rest;) {
      const size_t v = up[i] < rest ? up[i] : rest;
      pos += v;
      rest -= v;
      out->push_back(pos);
    }
  }
}

// The chunk boundaries of one host-input MSM of n > 0 terms and the number K1 of phase-1 chunks.
static void chunk_schedule(size_t n, size_t chunk, int phase1_percent, bool ramp, std::vector<size_t>* bnd, size_t* K1) {
  bnd->clear();
  bnd->push_back(0);
  if (ramp && n > chunk) {
    const size_t pct = (size_t)phase1_percent;
    size_t P1 = n / 100 * pct + (n % 100) * pct / 100;
    if (P1 < 1) P1 = 1;
    const size_t small = chunk / 8 > 1024 ? chunk / 8 : 1024;
    phase_bounds(0, P1, small, 2 * chunk, true, bnd);
    *K1 = bnd->size() - 1;
    phase_bounds(P1, n, small, 2 * chunk, false, bnd);
  } else {
    for (size_t lo = chunk; lo < n; lo += chunk) bnd->push_back(lo);
    bnd->push_back(n);
    const size_t nch = bnd->size() - 1;
    size_t k1 = (nch * (size_t)phase1_percent + 99) / 100;
    if (k1 < 1) k1 = 1;
    if (k1 > nch) k1 = nch;
    *K1 = k1;
  }
}
// test hook (no GPU needed): the schedule msm_from_host would use; returns the number of chunks (or -1 if cap is too small)
extern "C" int64_t zkp_selftest_chunk_schedule(size_t n, size_t chunk, int32_t phase1_percent, int32_t ramp,
                                               size_t* bounds_out, size_t cap, size_t* k1_out) {
  if (!n || chunk < 1024 || phase1_percent < 1 || phase1_percent > 100 || !bounds_out || !k1_out) return -1;
  try {
    std::vector<size_t> bnd;
    size_t K1 = 0;
    chunk_schedule(n, chunk, phase1_percent, ramp != 0, &bnd, &K1);
    if (bnd.size() > cap) return -1;
    for (size_t i = 0; i < bnd.size(); i++) bounds_out[i] = bnd[i];
    *k1_out = K1;
    return (int64_t)bnd.size() - 1;
  } catch (...) {
    return -1;
  }
}

static int32_t msm_from_host(zkp_ctx* ctx, const hseg* sc_segs, const hseg* pt_segs, int nseg, size_t n,
                             msm_result* d_result) {
  cudaStream_t st = ctx->stream;
  if (n == 0) {
    k_empty_result<<<1, 1, 0, st>>>(d_result);
    LAUNCH_CHECK(ctx);
    return ZKP_OK;
  }
  if (n >= 0x7fffffffull) {   // refuse before staging anything (msm_prepare has the same bound)
    ctx->err = "n too large (>= 2^31 - 1 terms per call)";
    return ZKP_ERR_SIZE;
  }
  ENSURE(ctx, ctx->in_scalars, n * 32);
  ENSURE(ctx, ctx->in_points, n * 32);
  if (use_small_path(ctx, n)) {   // few terms: copies and kernels on the one stream
    size_t off = 0;
    for (int k = 0; k < nseg; k++) {
      if (sc_segs[k].cnt) {
        CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t*)ctx->in_scalars.p + off * 32, sc_segs[k].p, sc_segs[k].cnt * 32,
                                      cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t*)ctx->in_points.p + off * 32, pt_segs[k].p, pt_segs[k].cnt * 32,
                                      cudaMemcpyHostToDevice, st));
      }
      off += sc_segs[k].cnt;
    }
    return msm_small_launch(ctx, ctx->in_scalars.p, ctx->in_points.p, n, d_result);
  }
  if (!ctx->copy_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  if (ctx->profile) {   // per-stage timing wants the stages back to back: copy everything first, then one pass
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    int32_t rc = copy_range(ctx, pt_segs, nseg, 0, n, (uint8_t*)ctx->in_points.p);
    if (rc != ZKP_OK) return rc;
    rc = copy_range(ctx, sc_segs, nseg, 0, n, (uint8_t*)ctx->in_scalars.p);
    if (rc != ZKP_OK) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
    return msm_vartime_launch(ctx, ctx->in_scalars.p, ctx->in_points.p, n, d_result);
  }
  msm_plan pl;
  int32_t r = msm_prepare(ctx, n, &pl);
  if (r != ZKP_OK) return r;
  const size_t chunk = ctx->chunk_terms;
  const bool fused = use_fused_sort(ctx, pl);
  // chunk k = points [bnd[k], bnd[k + 1]); the first K1 chunks form phase 1 of the two-phase ingestion
  std::vector<size_t> bnd;
  size_t K1 = 0;
  try {
    chunk_schedule(n, chunk, ctx->phase1_percent, fused && ZKP_ABL(ctx->ramp_chunks), &bnd, &K1);
  } catch (...) {
    ctx->err = "host allocation failed";
    return ZKP_ERR_NOMEM;
  }
  const size_t nchunks = bnd.size() - 1;
  while (ctx->chunk_ev.size() < nchunks + 1) {
    cudaEvent_t e;
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->chunk_ev.push_back(e);
  }
  // the copy stream must not overwrite the staging buffers before earlier work on `st` (a previous call) is done
  CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[nchunks], st));
  CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[nchunks], 0));
  if (fused) {
    // two-phase ingestion over the chunk pipeline: the point chunks of phase 1 travel with ALL the scalars (cut into
    // slices in proportion to the chunks) and carry the digit histogram; after the scan the remaining point chunks carry
    // the scatter.  Share of the points in phase 1 (which also waits for all the scalars): measured end to end at the
    // bench size, 50 % -> 136.3 ms, 60 % -> 137.6 ms, 70 % -> 141.6 ms per step, so the halves stay equal
    const size_t K2 = nchunks - K1, P1 = bnd[K1];
    // two compute streams: odd chunks run on the auxiliary stream (fork after the preparation, join around the scan
    // and before the bucket phase)
    const bool dual = ctx->dual_stream && nchunks > 2;
    cudaStream_t aux = st;
    if (dual) {
      if (!ctx->aux_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
      if (!ctx->ev_fork) {
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
      }
      aux = ctx->aux_stream;
      CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, st));
      CUDA_TRY(ctx, cudaStreamWaitEvent(aux, ctx->ev_fork, 0));
    }
    for (size_t k = 0; k < nchunks; k++) {
      const size_t lo = bnd[k], hi = bnd[k + 1];
      const bool first = k < K1;
      const size_t ps = first ? 0 : P1, plen = first ? P1 : n - P1;
      const size_t s_lo = n * (lo - ps) / plen, s_hi = n * (hi - ps) / plen;   // n < 2^31: no overflow
      const cudaStream_t sk = (k & 1) ? aux : st;
      r = copy_range(ctx, pt_segs, nseg, lo, hi, (uint8_t*)ctx->in_points.p);
      if (r != ZKP_OK) return r;
      if (first) {
        r = copy_range(ctx, sc_segs, nseg, s_lo, s_hi, (uint8_t*)ctx->in_scalars.p);
        if (r != ZKP_OK) return r;
      }
      CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[k], ctx->copy_stream));
      if (k == K1) {   // the scan needs every histogram launch of both streams; the scatter launches need the scan
        if (dual) {
          CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, aux));
          CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
        }
        k_scan<<<pl.W, 1024, 0, st>>>((const uint32_t*)ctx->hist.p, pl.B, (uint32_t*)ctx->offs.p, (uint32_t*)ctx->cursor.p);
        LAUNCH_CHECK(ctx);
        if (dual) {
          CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, st));
          CUDA_TRY(ctx, cudaStreamWaitEvent(aux, ctx->ev_fork, 0));
        }
      }
      CUDA_TRY(ctx, cudaStreamWaitEvent(sk, ctx->chunk_ev[k], 0));
      if (first) r = launch_ingest2_range<0>(ctx, pl, ctx->in_scalars.p, ctx->in_points.p, n, lo, hi - lo, s_lo, s_hi - s_lo, sk);
      else r = launch_ingest2_range<1>(ctx, pl, ctx->in_scalars.p, ctx->in_points.p, n, lo, hi - lo, s_lo, s_hi - s_lo, sk);
      if (r != ZKP_OK) return r;
    }
    if (dual) {   // everything the auxiliary stream did is ordered before what follows on `st`
      CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, aux));
      CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
    }
    if (K2 == 0) {   // every point went through phase 1: the scatter has no points to ride under
      k_scan<<<pl.W, 1024, 0, st>>>((const uint32_t*)ctx->hist.p, pl.B, (uint32_t*)ctx->offs.p, (uint32_t*)ctx->cursor.p);
      LAUNCH_CHECK(ctx);
      r = launch_ingest2_range<1>(ctx, pl, ctx->in_scalars.p, ctx->in_points.p, n, 0, 0, 0, n);
      if (r != ZKP_OK) return r;
    }
    return msm_finish(ctx, pl, ctx->in_scalars.p, n, d_result, false, true);
  }
  for (size_t k = 0; k < nchunks; k++) {
    const size_t lo = bnd[k], hi = bnd[k + 1];
    r = copy_range(ctx, pt_segs, nseg, lo, hi, (uint8_t*)ctx->in_points.p);
    if (r != ZKP_OK) return r;
    r = copy_range(ctx, sc_segs, nseg, lo, hi, (uint8_t*)ctx->in_scalars.p);
    if (r != ZKP_OK) return r;
    CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[k], ctx->copy_stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->chunk_ev[k], 0));
    if (pl.sort != st) CUDA_TRY(ctx, cudaStreamWaitEvent(pl.sort, ctx->chunk_ev[k], 0));
    r = msm_ingest(ctx, pl, ctx->in_scalars.p, ctx->in_points.p, lo, hi - lo, false);
    if (r != ZKP_OK) return r;
  }
  return msm_finish(ctx, pl, ctx->in_scalars.p, n, d_result, false);
}

extern "C" int32_t zkp_msm_vartime_dev(zkp_ctx* ctx, const void* d_scalars, const void* d_points, size_t n,
                                       void* d_result) {
  if (!ctx || !d_result || (n && (!d_scalars || !d_points))) return ZKP_ERR_SIZE;
  if (!aligned16(d_scalars) || !aligned16(d_points) || !aligned16(d_result)) {
    ctx->err = "device pointers must be 16-byte aligned";
    return ZKP_ERR_SIZE;
  }
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  return msm_vartime_launch(ctx, d_scalars, d_points, n, (msm_result*)d_result);
}

static int32_t fetch_result(zkp_ctx* ctx, uint8_t* out32, int32_t* is_identity, int64_t* first_bad) {
  msm_result* h = (msm_result*)ctx->h_result;
  CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->result.p, sizeof(msm_result), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (first_bad) *first_bad = h->first_bad;
  if (h->status != ZKP_OK) {
    if (is_identity) *is_identity = 0;
    return h->status;
  }
  if (out32) memcpy(out32, h->enc, 32);
  if (is_identity) *is_identity = h->is_identity;
  return ZKP_OK;
}

extern "C" int32_t zkp_msm_vartime(zkp_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n,
                                   uint8_t* out32, int32_t* is_identity, int64_t* first_bad) {
  if (!ctx || (n && (!scalars || !points))) return ZKP_ERR_SIZE;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ENSURE(ctx, ctx->result, 64);
  hseg ss = {scalars, n}, ps = {points, n};
  int32_t r = msm_from_host(ctx, &ss, &ps, 1, n, (msm_result*)ctx->result.p);
  if (r != ZKP_OK) return r;
  return fetch_result(ctx, out32, is_identity, first_bad);
}

extern "C" int32_t zkp_batch_verify(zkp_ctx* ctx, const uint8_t* static_coeffs, const uint8_t* static_points,
                                    size_t num_s, const uint8_t* instance_coeffs, const uint8_t* instance_points,
                                    size_t rows, size_t batch, int32_t* accept, int64_t* first_bad) {
  if (!ctx || !accept) return ZKP_ERR_SIZE;
  *accept = 0;
  if ((num_s && (!static_coeffs || !static_points)) || (rows && batch && (!instance_coeffs || !instance_points)))
    return ZKP_ERR_SIZE;
  if (batch && rows > ((size_t)1 << 40) / batch) return ZKP_ERR_SIZE;
  const size_t n_inst = rows * batch, n = num_s + n_inst;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ENSURE(ctx, ctx->result, 64);
  // scalar order: static ++ row-major instance matrix; point order: static ++ rows   (batch_verifier.rs:219-226)
  hseg ss[2] = {{static_coeffs, num_s}, {instance_coeffs, n_inst}};
  hseg ps[2] = {{static_points, num_s}, {instance_points, n_inst}};
  int32_t r = msm_from_host(ctx, ss, ps, 2, n, (msm_result*)ctx->result.p);
  if (r != ZKP_OK) return r;
  int32_t ident = 0;
  r = fetch_result(ctx, nullptr, &ident, first_bad);
  if (r != ZKP_OK) return r;
  *accept = ident;
  return ZKP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// single-verdict mode over shards: partial sums out, verdict on their sum
// ---------------------------------------------------------------------------------------------------------
extern "C" int32_t zkp_batch_verify_partial(zkp_ctx* ctx, const uint8_t* static_coeffs, const uint8_t* static_points,
                                            size_t num_s, const uint8_t* instance_coeffs, const uint8_t* instance_points,
                                            size_t rows, size_t batch, uint64_t* partial_limbs_out, int64_t* first_bad) {
  if (!ctx || !partial_limbs_out) return ZKP_ERR_SIZE;
  if ((num_s && (!static_coeffs || !static_points)) || (rows && batch && (!instance_coeffs || !instance_points)))
    return ZKP_ERR_SIZE;
  if (batch && rows > ((size_t)1 << 40) / batch) return ZKP_ERR_SIZE;
  const size_t n_inst = rows * batch, n = num_s + n_inst;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ENSURE(ctx, ctx->result, 64);
  ENSURE(ctx, ctx->partial_buf, 128 + 160);
  cudaStream_t st = ctx->stream;
  if (n == 0) {   // the empty sum: the identity (0 : 1 : 1 : 0)
    memset(partial_limbs_out, 0, 160);
    partial_limbs_out[5] = 1;
    partial_limbs_out[10] = 1;
    if (first_bad) *first_bad = -1;
    return ZKP_OK;
  }
  hseg ss[2] = {{static_coeffs, num_s}, {instance_coeffs, n_inst}};
  hseg ps[2] = {{static_points, num_s}, {instance_points, n_inst}};
  ctx->partial_out = ctx->partial_buf.p;
  int32_t r = msm_from_host(ctx, ss, ps, 2, n, (msm_result*)ctx->result.p);
  ctx->partial_out = nullptr;
  if (r != ZKP_OK) return r;
  k_ext_to_limbs<<<1, 1, 0, st>>>((const uint4*)ctx->partial_buf.p, (unsigned long long*)((uint8_t*)ctx->partial_buf.p + 128));
  LAUNCH_CHECK(ctx);
  CUDA_TRY(ctx, cudaMemcpyAsync(partial_limbs_out, (uint8_t*)ctx->partial_buf.p + 128, 160, cudaMemcpyDeviceToHost, st));
  return fetch_result(ctx, nullptr, nullptr, first_bad);   // reports invalid points / scalars of this shard
}

extern "C" int32_t zkp_partials_verdict(zkp_ctx* ctx, const uint64_t* partial_limbs, size_t count, int32_t* accept,
                                        uint8_t* enc_out32) {
  if (!ctx || !accept || (count && !partial_limbs)) return ZKP_ERR_SIZE;
  *accept = 0;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ENSURE(ctx, ctx->result, 64);
  ENSURE(ctx, ctx->aux0, count * 160 + 160);
  cudaStream_t st = ctx->stream;
  if (count) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->aux0.p, partial_limbs, count * 160, cudaMemcpyHostToDevice, st));
  k_sum_partials<<<1, 1, 0, st>>>((const unsigned long long*)ctx->aux0.p, count, (msm_result*)ctx->result.p);
  LAUNCH_CHECK(ctx);
  int32_t ident = 0;
  int32_t r = fetch_result(ctx, enc_out32, &ident, nullptr);
  if (r != ZKP_OK) return r;
  *accept = ident;
  return ZKP_OK;
}

// device-resident forms: the partial sum stays on the device (160 bytes of limbs), so the all-gather of the shards' sums
// (NCCL, on the same stream) and the verdict follow without a host round trip
extern "C" int32_t zkp_msm_vartime_partial_dev(zkp_ctx* ctx, const void* d_scalars, const void* d_points, size_t n,
                                               void* d_result, void* d_partial_limbs) {
  if (!ctx || !d_result || !d_partial_limbs || (n && (!d_scalars || !d_points))) return ZKP_ERR_SIZE;
  if (!aligned16(d_scalars) || !aligned16(d_points) || !aligned16(d_result) || !aligned16(d_partial_limbs)) {
    ctx->err = "device pointers must be 16-byte aligned";
    return ZKP_ERR_SIZE;
  }
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ENSURE(ctx, ctx->partial_buf, 128 + 160);
  cudaStream_t st = ctx->stream;
  if (n == 0) {   // the empty sum: the identity (0 : 1 : 1 : 0)
    uint64_t ident[20] = {0};
    ident[5] = 1;
    ident[10] = 1;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_partial_limbs, ident, 160, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));   // `ident` is on this stack frame
    k_empty_result<<<1, 1, 0, st>>>((msm_result*)d_result);
    LAUNCH_CHECK(ctx);
    return ZKP_OK;
  }
  ctx->partial_out = ctx->partial_buf.p;
  int32_t r = msm_vartime_launch(ctx, d_scalars, d_points, n, (msm_result*)d_result);
  ctx->partial_out = nullptr;
  if (r != ZKP_OK) return r;
  k_ext_to_limbs<<<1, 1, 0, st>>>((const uint4*)ctx->partial_buf.p, (unsigned long long*)d_partial_limbs);
  LAUNCH_CHECK(ctx);
  return ZKP_OK;
}

extern "C" int32_t zkp_partials_verdict_dev(zkp_ctx* ctx, const void* d_partial_limbs, size_t count, void* d_result) {
  if (!ctx || !d_result || (count && !d_partial_limbs)) return ZKP_ERR_SIZE;
  if (!aligned16(d_partial_limbs) || !aligned16(d_result)) {
    ctx->err = "device pointers must be 16-byte aligned";
    return ZKP_ERR_SIZE;
  }
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  k_sum_partials<<<1, 1, 0, ctx->stream>>>((const unsigned long long*)d_partial_limbs, count, (msm_result*)d_result);
  LAUNCH_CHECK(ctx);
  return ZKP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// entry points that build host-side containers: nothing may unwind through the C ABI
// ---------------------------------------------------------------------------------------------------------
template <class F>
static int32_t guarded(zkp_ctx* ctx, F&& body) {
  try {
    return body();
  } catch (const std::bad_alloc&) {
    if (ctx) ctx->err = "host allocation failed";
    return ZKP_ERR_NOMEM;
  } catch (const std::exception& e) {
    if (ctx) ctx->err = std::string("host exception: ") + e.what();
    return ZKP_ERR_SIZE;
  } catch (...) {
    return ZKP_ERR_SIZE;
  }
}

// a flattened statement is used as sizes and indices: check it once, before anything is derived from it
static bool statement_ok(const zkp_statement_desc* sd) {
  const int m = sd->m, ni = sd->ni, nc = sd->nc, k = sd->k;
  if (m < 0 || ni < 0 || nc < 0 || k < 0) return false;
  if (ni > 2 * ZKP_BV_MAX_VARS || nc > 2 * ZKP_BV_MAX_VARS || k > ZKP_BV_MAX_CONS) return false;
  const int p = ni + nc;
  if (p && !sd->labels) return false;
  if (!k) return true;
  if (!sd->lhs || !sd->cons_off || sd->cons_off[0] != 0) return false;
  for (int c = 0; c < k; c++)
    if (sd->lhs[c] < 0 || sd->lhs[c] >= p || sd->cons_off[c + 1] < sd->cons_off[c]) return false;
  const int n_terms = sd->cons_off[k];
  if (n_terms > (1 << 20)) return false;
  if (n_terms && (!sd->term_scalar || !sd->term_point)) return false;
  for (int q = 0; q < n_terms; q++)
    if (sd->term_point[q] < 0 || sd->term_point[q] >= p || sd->term_scalar[q] < 0 || sd->term_scalar[q] >= m) return false;
  return true;
}

// Host execution of a compiled script for ONE proof (what a thread of k_bv_prepare2 does): the 64 challenge bytes.
// Test hook: lets the CPU suite check the script compiler against the byte-wise Merlin of host/merlin.cpp.
extern "C" int32_t zkp_selftest_bv_script(const zkp_statement_desc* sd, const uint32_t* prefix_state,
                                          const uint8_t* instance_enc, const uint8_t* common_enc,
                                          const uint8_t* commitments, uint8_t* challenge_out64, int32_t* n_blocks_out) {
  if (!sd || !prefix_state || !challenge_out64 || !statement_ok(sd)) return ZKP_ERR_SIZE;
  if ((sd->ni && !instance_enc) || (sd->nc && !common_enc) || (sd->k && !commitments)) return ZKP_ERR_SIZE;
  return guarded(nullptr, [&]() -> int32_t {
  bv_script script(prefix_state[50], prefix_state[51]);
  bv_compile_batch_verify(script, sd, common_enc);
  uint64_t st[25];
  for (int i = 0; i < 25; i++) st[i] = (uint64_t)prefix_state[2 * i] | ((uint64_t)prefix_state[2 * i + 1] << 32);
  const int nb = (int)(script.tmpl.size() / 21);
  for (int b = 0; b < nb; b++) {
    uint8_t blk[168];
    for (int l = 0; l < 21; l++)
      for (int x = 0; x < 8; x++) blk[8 * l + x] = (uint8_t)(script.tmpl[(size_t)b * 21 + l] >> (8 * x));
    for (uint32_t si = script.seg_start[b]; si < script.seg_start[b + 1]; si++) {
      const bv_seg& sg = script.segs[si];
      const uint8_t* src = sg.kind == 0 ? instance_enc + 32 * (size_t)sg.idx : commitments + 32 * (size_t)sg.idx;
      for (uint32_t x = 0; x < sg.len; x++) {
        const int pos = sg.shift + (int)(sg.src_off + x);
        if (pos < 0 || pos >= 166) return ZKP_ERR_CUDA;   // a segment must stay inside the rate
        blk[pos] ^= src[sg.src_off + x];
      }
    }
    for (int l = 0; l < 21; l++) {
      uint64_t v = 0;
      for (int x = 7; x >= 0; x--) v = (v << 8) | blk[8 * l + x];
      st[l] ^= v;
    }
    zkp_host::keccak_f1600(st);
  }
  for (int i = 0; i < 64; i++) challenge_out64[i] = (uint8_t)(st[i >> 3] >> (8 * (i & 7)));
  if (n_blocks_out) *n_blocks_out = nb;
  return ZKP_OK;
  });
}

// ---------------------------------------------------------------------------------------------------------
// batch verification from proofs: transcripts, challenges, weights and coefficient fold on the device
// ---------------------------------------------------------------------------------------------------------
// ZKP_BV_TIMELINE=1 in the environment: zkp_batch_verify_proofs records a timed event after every copy, front-end kernel and
// ingestion launch and prints, after the call, when each of them ENDED (ms after the call's first event) -- a development
// aid for the slab pipeline (no profiler needed); the call runs as usual otherwise.
struct bv_timeline {
  bool on = false;
  std::vector<std::pair<std::string, cudaEvent_t>> marks;
  bv_timeline() { const char* e = getenv("ZKP_BV_TIMELINE"); on = e && *e && *e != '0'; }
  void mark(const char* what, size_t idx, cudaStream_t s) {
    if (!on) return;
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, s);
    marks.emplace_back(std::string(what) + "[" + std::to_string(idx) + "]", ev);
  }
  void report() {
    if (!on || marks.empty()) return;
    cudaDeviceSynchronize();
    for (auto& m : marks) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, marks[0].second, m.second);
      fprintf(stderr, "bv_timeline %-18s %9.3f ms\n", m.first.c_str(), ms);
    }
    for (auto& m : marks) cudaEventDestroy(m.second);
    marks.clear();
  }
};
// The MSM terms of a batch are a (rows x N) matrix behind the nc static terms (term (row, j) = nc + row * N + j).  One
// ingestion step over the columns [j0, j0 + cnt): the points of the P rows from p_row0 on are decompressed, and the thread
// that takes point (p_row0 + q, j) also takes the digits of the terms (q, j), (q + P, j), (q + 2 P, j) -- as many of the
// `rows` rows as exist -- so the digit work of ALL rows rides under the decompression of P of them (rows <= 3 P; rows beyond
// that get digit-only launches).  Rows with the same number of term ranges go out as ONE grid (blockIdx.y = row).
template <int MODE>
static int32_t bv_rows_launch(zkp_ctx* ctx, const msm_plan& pl, const void* dsc, const void* dpts, size_t n, size_t nc,
                              size_t N, size_t rows, size_t p_row0, size_t P, size_t j0, size_t cnt) {
  auto clampP = [&](size_t x) { return x > P ? P : x; };
  const size_t g3 = rows > 2 * P ? clampP(rows - 2 * P) : 0, g2 = rows > P ? clampP(rows - P) : 0;
  const size_t lo[4] = {0, g3, g2, P};   // q in [lo[i], lo[i+1]) carries 3 - i term ranges
  for (int i = 0; i < 3; i++) {
    const size_t q0 = lo[i], q1 = lo[i + 1];
    if (q1 <= q0) continue;
    ingest_args a = {};
    a.p_lo = nc + (p_row0 + q0) * N + j0; a.p_cnt = cnt; a.y_p = N;
    for (int j = 0; j < 3 - i; j++) {
      a.s_lo[j] = nc + (q0 + (size_t)j * P) * N + j0; a.s_cnt[j] = cnt; a.y_s[j] = N;
    }
    int32_t r = launch_ingest2<MODE>(ctx, pl, dsc, dpts, n, a, (unsigned)(q1 - q0));
    if (r != ZKP_OK) return r;
  }
  for (size_t row = 3 * P; row < rows; row += 3) {   // digit-only: no decompression left to hide them under
    ingest_args a = {};
    for (int j = 0; j < 3 && row + j < rows; j++) {
      a.s_lo[j] = nc + (row + j) * N + j0; a.s_cnt[j] = cnt;
    }
    int32_t r = launch_ingest2<MODE>(ctx, pl, dsc, dpts, n, a);
    if (r != ZKP_OK) return r;
  }
  return ZKP_OK;
}
static int32_t batch_verify_proofs_impl(zkp_ctx* ctx, const zkp_statement_desc* sd, const uint32_t* prefix_state,
                                        size_t N, const uint8_t* instance_enc, const uint8_t* common_enc,
                                        const uint8_t* commitments, const uint8_t* responses,
                                        const uint8_t* rho_seed, int32_t* accept, int64_t* first_bad,
                                        uint8_t* coeff_out, uint8_t* points_out) {
  if (!ctx || !sd || !prefix_state || !accept) return ZKP_ERR_SIZE;
  *accept = 0;
  uint8_t own_seed[32];
  if (!rho_seed) {   // no seed: the weights' seed comes from the OS CSPRNG, fresh for this call (batch_verifier.rs:179 thread_rng)
    if (!os_random_bytes(own_seed, 32)) {
      ctx->err = "getrandom failed";
      return ZKP_ERR_CUDA;
    }
    rho_seed = own_seed;
  }
  if (first_bad) *first_bad = -1;
  const int m = sd->m, ni = sd->ni, nc = sd->nc, k = sd->k;
  if (!statement_ok(sd)) {
    ctx->err = "inconsistent statement description";
    return ZKP_ERR_SIZE;
  }
  if (ni > ZKP_BV_MAX_VARS || nc > ZKP_BV_MAX_VARS) {
    ctx->err = "statement too large for the device front end";
    return ZKP_ERR_SIZE;
  }
  if (N && ((ni && !instance_enc) || (k && !commitments) || (m && !responses))) return ZKP_ERR_SIZE;
  if (nc && !common_enc) return ZKP_ERR_SIZE;
  for (int s0 = 0; s0 < nc; s0++) {   // identity encoding of a static point fails at allocation (toolbox/mod.rs:191)
    uint8_t o = 0;
    for (int b = 0; b < 32; b++) o |= common_enc[32 * s0 + b];
    if (!o && N) return ZKP_ERR_POINT;
  }
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // ---- flatten the statement: label offsets, transcript script, constraint arrays -> one small device blob ----
  bv_plan bp;
  bv_make_plan(sd, prefix_state, rho_seed, common_enc, &bp);
  const std::vector<uint8_t>& blob = bp.blob;
  const size_t blob_sz = blob.size(), o_prefix = bp.o_prefix, o_seed = bp.o_seed, o_tm = bp.o_tm, o_ss = bp.o_ss, o_sg = bp.o_sg;
  const int script_blocks = bp.script_blocks;
  const size_t rows = (size_t)ni + k, n = (size_t)nc + rows * N;
  if (n >= 0x7fffffffull) return ZKP_ERR_SIZE;
  const size_t chunk = ctx->bv_chunk_terms / (rows ? rows : 1) > 1024 ? (ctx->bv_chunk_terms * 4 / (rows ? rows : 1)) & ~(size_t)127
                                                                    : 1024;   // proofs per chunk (multiple of 128)
  const size_t nchunks = N ? (N + chunk - 1) / chunk : 0;
  ENSURE(ctx, ctx->bv_misc, blob_sz);
  ENSURE(ctx, ctx->in_scalars, n * 32 + 32);
  ENSURE(ctx, ctx->in_points, n * 32 + 32);
  ENSURE(ctx, ctx->bv_com, (size_t)k * N * 32 + 32);
  ENSURE(ctx, ctx->bv_resp, (size_t)m * N * 32 + 32);
  ENSURE(ctx, ctx->result, 64);
  const unsigned nblocks_total = (unsigned)(nchunks * ((chunk + 127) / 128) + 1);
  ENSURE(ctx, ctx->bv_part, (size_t)nblocks_total * (nc ? nc : 1) * 32);
  if (!ctx->copy_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  while (ctx->chunk_ev.size() < nchunks + 1) {
    cudaEvent_t e;
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->chunk_ev.push_back(e);
  }
  uint8_t* dm = (uint8_t*)ctx->bv_misc.p;
  uint8_t* dpts = (uint8_t*)ctx->in_points.p;
  uint8_t* dsc = (uint8_t*)ctx->in_scalars.p;
  CUDA_TRY(ctx, cudaMemcpyAsync(dm, blob.data(), blob_sz, cudaMemcpyHostToDevice, st));
  if (nc) CUDA_TRY(ctx, cudaMemcpyAsync(dpts, common_enc, (size_t)nc * 32, cudaMemcpyHostToDevice, st));
  if (n == 0) {
    k_empty_result<<<1, 1, 0, st>>>((msm_result*)ctx->result.p);
    LAUNCH_CHECK(ctx);
    int32_t ident0 = 0;
    int32_t r0 = fetch_result(ctx, nullptr, &ident0, first_bad);
    if (r0 != ZKP_OK) return r0;
    *accept = ident0;
    return ZKP_OK;
  }
  msm_plan pl;
  int32_t r = msm_prepare(ctx, n, &pl);   // sizes the MSM workspace, resets flags and the digit histogram
  if (r != ZKP_OK) return r;
  bv_desc d;
  bv_fill_desc(&d, sd, bp, dm);
  // the copy stream must not run ahead of earlier work on `st` that still uses the staging buffers
  CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[nchunks], st));
  CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[nchunks], 0));
  unsigned block_base = 0;
  const bool fused = use_fused_sort(ctx, pl);
  bv_timeline tl;
  tl.mark("start", 0, st);
  // Rows of a slab decompressed during phase 1 (under the histogram of all its rows, as the proofs arrive); the others follow
  // in phase 2 under the scatter.  The split is chosen PER SLAB.  Half of the rows per phase is the split under which neither
  // the histogram nor the scatter costs a pass, and it is right while the GPU is the bottleneck (every order costs the same).
  // When a slab lands and the GPU has already finished the previous one, the pipeline is waiting for the host link (several
  // GPUs sharing one host's memory bandwidth): the split moves one step towards phase 1 -- two thirds of the rows (the scatter
  // of all rows still hides under the last third: three term ranges per thread), then all of them (the slab's digits are then
  // scattered on their own in phase 2) -- and one step back whenever a slab finds the GPU still busy.  bv_phase1_rows > 0
  // fixes the split for every slab; -1 alternates between half and all (tests).
  const size_t R_half = rows / 2, R_23 = rows - (rows + 2) / 3, R_all = rows;
  const bool adaptive = fused && ctx->bv_phase1_rows == 0 && nchunks > 1;
  int level = 0;   // 0: half of the rows in phase 1; 1: two thirds (the scatter still hides under the last third); 2: all
  std::vector<size_t> r1_of(nchunks, R_half);
  if (ctx->bv_phase1_rows > 0)
    for (size_t c = 0; c < nchunks; c++) r1_of[c] = (size_t)ctx->bv_phase1_rows > rows ? rows : (size_t)ctx->bv_phase1_rows;
  if (ctx->bv_phase1_rows < 0)
    for (size_t c = 0; c < nchunks; c++) r1_of[c] = (c & 1) ? R_all : R_half;
  while (adaptive && ctx->ing_ev.size() < nchunks) {
    cudaEvent_t e;
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->ing_ev.push_back(e);
  }
  // the front end (k_bv_prepare2: Keccak and scalar arithmetic, ALU pipe) of slab i + 1 runs on its own stream, at a higher
  // priority and with its residency capped (dynamic shared memory it does not use), NEXT TO the decompression of slab i
  // (integer multiplier) instead of in front of it
  cudaStream_t ps = st;
  size_t prep_smem = 0;
  if (ctx->bv_prep_stream && fused && nchunks > 1) {
    if (ctx->sm_count < 0) CUDA_TRY(ctx, cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, ctx->device));
    if (!ctx->prep_stream) {
      int plo = 0, phi = 0;
      CUDA_TRY(ctx, cudaDeviceGetStreamPriorityRange(&plo, &phi));
      CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->prep_stream, cudaStreamNonBlocking, phi));
    }
    ps = ctx->prep_stream;
    prep_smem = (size_t)ctx->bv_prep_smem_kb << 10;
    if (prep_smem > ctx->bv_prep_smem_set) {
      CUDA_TRY(ctx, cudaFuncSetAttribute((const void*)k_bv_prepare2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
      ctx->bv_prep_smem_set = prep_smem;
    }
    while (ctx->prep_ev.size() < nchunks) {
      cudaEvent_t e;
      CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctx->prep_ev.push_back(e);
    }
    CUDA_TRY(ctx, cudaStreamWaitEvent(ps, ctx->chunk_ev[nchunks], 0));   // the blob, the reset flags and histogram
  }
  // H2D of a slab on the copy stream: instance slab (ni rows of cnt encodings), commitments and responses.  The copies run
  // two slabs ahead of the kernels (from pageable memory cudaMemcpyAsync returns only when the slab is staged)
  size_t copies_enqueued = 0;
  auto enqueue_copies_upto = [&](size_t upto) -> int32_t {
    for (; copies_enqueued < nchunks && copies_enqueued <= upto; copies_enqueued++) {
      const size_t cc = copies_enqueued, j0 = cc * chunk, cnt = j0 + chunk < N ? chunk : N - j0;
      if (ni)
        CUDA_TRY(ctx, cudaMemcpy2DAsync(dpts + ((size_t)nc + j0) * 32, N * 32, instance_enc + j0 * 32, N * 32, cnt * 32,
                                        (size_t)ni, cudaMemcpyHostToDevice, ctx->copy_stream));
      if (k)
        CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t*)ctx->bv_com.p + j0 * k * 32, commitments + j0 * k * 32, cnt * k * 32,
                                      cudaMemcpyHostToDevice, ctx->copy_stream));
      if (m)
        CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t*)ctx->bv_resp.p + j0 * m * 32, responses + j0 * m * 32, cnt * m * 32,
                                      cudaMemcpyHostToDevice, ctx->copy_stream));
      CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[cc], ctx->copy_stream));
      tl.mark("copy", cc, ctx->copy_stream);
    }
    return ZKP_OK;
  };
  for (size_t cidx = 0; cidx < nchunks; cidx++) {
    const size_t j0 = cidx * chunk, cnt = j0 + chunk < N ? chunk : N - j0;
    r = enqueue_copies_upto(cidx + 2);
    if (r != ZKP_OK) return r;
    if (adaptive) {
      // the host follows the arrival of the slabs: was the GPU already done with the previous slab when this one landed?
      CUDA_TRY(ctx, cudaEventSynchronize(ctx->chunk_ev[cidx]));
      const bool idle = cidx > 0 && cudaEventQuery(ctx->ing_ev[cidx - 1]) == cudaSuccess;
      level = idle ? (level < 2 ? level + 1 : 2) : (level > 0 ? level - 1 : 0);
      r1_of[cidx] = level == 2 ? R_all : level == 1 ? R_23 : R_half;
    }
    const size_t R1 = r1_of[cidx];
    CUDA_TRY(ctx, cudaStreamWaitEvent(ps, ctx->chunk_ev[cidx], 0));
    if (pl.sort != st) CUDA_TRY(ctx, cudaStreamWaitEvent(pl.sort, ctx->chunk_ev[cidx], 0));
    unsigned nb = (unsigned)((cnt + 127) / 128);
#ifdef ZKP_ABLATIONS
    if (!ctx->bv_compiled)
      k_bv_prepare<<<nb, 128, 0, ps>>>(d, (const uint32_t*)(dm + o_prefix), N, dpts + (size_t)nc * 32, dpts,
                                       (const uint8_t*)ctx->bv_com.p, (const uint8_t*)ctx->bv_resp.p, dm + o_seed, dsc, dpts,
                                       (uint8_t*)ctx->bv_part.p, nullptr, (int*)ctx->flags.p, j0, cnt, block_base);
    else
#endif
    {
      // next to the decompression: a RESIDENT grid (every block placed at once, looping over the slab's 128-proof groups).
      // The block scheduler serves the higher-priority grid first and waits while it has blocks it cannot place, so a
      // grid larger than its residency cap would keep the ingestion kernel's blocks out for its whole duration.
      if (ps != st && ctx->bv_prep_blocks > 0 && nb > (unsigned)(ctx->bv_prep_blocks * ctx->sm_count))
        nb = (unsigned)(ctx->bv_prep_blocks * ctx->sm_count);
      k_bv_prepare2<<<nb, 128, prep_smem, ps>>>(d, (const uint32_t*)(dm + o_prefix), N, dpts + (size_t)nc * 32,
                                                (const uint8_t*)ctx->bv_com.p, (const uint8_t*)ctx->bv_resp.p, dm + o_seed,
                                                script_blocks, (const unsigned long long*)(dm + o_tm),
                                                (const uint32_t*)(dm + o_ss), (const bv_seg*)(dm + o_sg), dsc, dpts,
                                                (uint8_t*)ctx->bv_part.p, nullptr, (int*)ctx->flags.p, j0, cnt, block_base);
    }
    LAUNCH_CHECK(ctx);
    block_base += nb;
    tl.mark("prepare", cidx, ps);
    if (ps != st) {
      CUDA_TRY(ctx, cudaEventRecord(ctx->prep_ev[cidx], ps));
      CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->prep_ev[cidx], 0));
    }
    // the slabs of this chunk are complete (points and coefficients)
    if (fused) {
      // phase 1 of the two-phase ingestion: histogram every slab, decompress the slabs of the first R1 rows
      r = bv_rows_launch<0>(ctx, pl, dsc, dpts, n, (size_t)nc, N, rows, 0, R1, j0, cnt);
      if (r != ZKP_OK) return r;
      if (adaptive) CUDA_TRY(ctx, cudaEventRecord(ctx->ing_ev[cidx], st));
      tl.mark(R1 == R_half ? "ingest1" : "ingest1*", cidx, st);
    } else {
      for (size_t row = 0; row < rows; row++) {
        r = msm_ingest(ctx, pl, dsc, dpts, (size_t)nc + row * N + j0, cnt, false);
        if (r != ZKP_OK) return r;
      }
    }
  }
  if (nc) {
    k_bv_static_sum<<<1, 256, 0, st>>>((const uint8_t*)ctx->bv_part.p, (int)block_base, nc, dsc);
    LAUNCH_CHECK(ctx);
    if (fused) r = launch_ingest2_range<0>(ctx, pl, dsc, dpts, n, 0, (size_t)nc, 0, (size_t)nc);
    else r = msm_ingest(ctx, pl, dsc, dpts, 0, (size_t)nc, false);
    if (r != ZKP_OK) return r;
  }
  if (fused) {
    // phase 2: the remaining rows are decompressed while the digits of all rows are scattered
    k_scan<<<pl.W, 1024, 0, st>>>((const uint32_t*)ctx->hist.p, pl.B, (uint32_t*)ctx->offs.p, (uint32_t*)ctx->cursor.p);
    LAUNCH_CHECK(ctx);
    tl.mark("scan", 0, st);
    for (size_t c0 = 0; c0 < nchunks;) {   // one step per run of slabs with the same split
      size_t c1 = c0 + 1;
      while (c1 < nchunks && r1_of[c1] == r1_of[c0]) c1++;
      const size_t j0 = c0 * chunk, j1 = c1 * chunk < N ? c1 * chunk : N;
      r = bv_rows_launch<1>(ctx, pl, dsc, dpts, n, (size_t)nc, N, rows, r1_of[c0], rows - r1_of[c0], j0, j1 - j0);
      if (r != ZKP_OK) return r;
      c0 = c1;
    }
    tl.mark("ingest2", 0, st);
    if (nc) {
      r = launch_ingest2_range<1>(ctx, pl, dsc, dpts, n, 0, 0, 0, (size_t)nc);
      if (r != ZKP_OK) return r;
    }
  }
  r = msm_finish(ctx, pl, dsc, n, (msm_result*)ctx->result.p, false, fused);
  if (r != ZKP_OK) return r;
  tl.mark("finish", 0, st);
  if (coeff_out) CUDA_TRY(ctx, cudaMemcpyAsync(coeff_out, dsc, n * 32, cudaMemcpyDeviceToHost, st));
  if (points_out) CUDA_TRY(ctx, cudaMemcpyAsync(points_out, dpts, n * 32, cudaMemcpyDeviceToHost, st));
  int32_t ident = 0;
  r = fetch_result(ctx, nullptr, &ident, first_bad);   // identity encodings / bad points -> 1, bad responses -> 3
  tl.report();
  if (r != ZKP_OK) return r;
  *accept = ident;
  return ZKP_OK;
}
extern "C" int32_t zkp_batch_verify_proofs(zkp_ctx* ctx, const zkp_statement_desc* sd, const uint32_t* prefix_state,
                                           size_t N, const uint8_t* instance_enc, const uint8_t* common_enc,
                                           const uint8_t* commitments, const uint8_t* responses,
                                           const uint8_t* rho_seed, int32_t* accept, int64_t* first_bad,
                                           uint8_t* coeff_out, uint8_t* points_out) {
  return guarded(ctx, [&]() -> int32_t {
    return batch_verify_proofs_impl(ctx, sd, prefix_state, N, instance_enc, common_enc, commitments, responses, rho_seed,
                                    accept, first_bad, coeff_out, points_out);
  });
}

// ---------------------------------------------------------------------------------------------------------
// batch proving: N proofs of one statement, per-proof transcript / nonce / response work on the device
// ---------------------------------------------------------------------------------------------------------
// One slice of a batch on one buffer set: everything up to the flags' way back is enqueued on the set's stream by
// prove_slice_enqueue; prove_slice_complete waits for it, reads the flags and only then lets the outputs leave the device.
// Two sets alternate, so the H2D copies of slice i + 1 and the D2H copies of slice i - 1 overlap the kernels of slice i.
struct pv_slice {
  size_t N = 0;
  const uint8_t* secrets = nullptr;
  const uint64_t* points = nullptr;
  const uint8_t* entropy = nullptr;
  uint8_t *encodings_out = nullptr, *commitments_out = nullptr, *responses_out = nullptr, *blindings_out = nullptr;
  bool share = false;
};
#define ZKP_PV_NOT_UNIFORM 1000   // internal: a "common" point differed between proofs while its table was shared

static int32_t prove_slice_enqueue(zkp_ctx* ctx, pv_set& S, const zkp_statement_desc* sd, const uint32_t* prefix_state,
                                   pv_slice& sl, bool share_allowed) {
  const size_t N = sl.N;
  const uint8_t* secrets = sl.secrets;
  const uint64_t* points = sl.points;
  const uint8_t* entropy = sl.entropy;
  const int m = sd->m, ni = sd->ni, nc = sd->nc, k = sd->k, p = ni + nc;
  const int n_terms = k ? sd->cons_off[k] : 0;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = S.stream;
  // ---- statement blob: prefix | label pool | label offsets/lengths | lhs | cons_off | term arrays | schedule slots ----
  std::vector<uint32_t> loff, llen;
  std::vector<uint8_t> pool;
  const char* lp = sd->labels;
  for (int i = 0; i < p; i++) {
    size_t len = strlen(lp);
    loff.push_back((uint32_t)pool.size());
    llen.push_back((uint32_t)len);
    pool.insert(pool.end(), lp, lp + len);
    lp += len + 1;
  }
  auto pad16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t o_prefix = 0, o_pool = pad16(53 * 4), o_lo = pad16(o_pool + pool.size()), o_ll = pad16(o_lo + (size_t)p * 4),
               o_lhs = pad16(o_ll + (size_t)p * 4), o_co = pad16(o_lhs + (size_t)k * 4), o_ts = pad16(o_co + (size_t)(k + 1) * 4),
               o_tp = pad16(o_ts + (size_t)n_terms * 4), o_sl = pad16(o_tp + (size_t)n_terms * 4),
               o_sh = pad16(o_sl + (size_t)k * 4), o_cs = pad16(o_sh + (size_t)n_terms * 4),
               o_cp = pad16(o_cs + (size_t)n_terms * 4), o_cq = pad16(o_cp + (size_t)p * 4),
               o_ut = pad16(o_cq + (size_t)p * 4), o_un = pad16(o_ut + ((size_t)n_terms + k) * 4),
               o_cu = pad16(o_un + ((size_t)n_terms + k) * 4), blob_sz = pad16(o_cu + (size_t)(k + 1) * 4) + 16;
  // batch-static bases: the statement's common points are the same for every proof when the caller says so
  // (points_are_uniform): their constant-time tables are built once and shared
  // (the caller passes a copy per proof, as the reference's per-proof assignments do: k_pv_gather compares every copy
  // with proof 0's on the device and the call is redone without sharing if one differs)
  const bool share = share_allowed && ctx->share_static_tables && nc > 0 && N > 1;
  // comb path: every point that serves as a base gets a comb slot -- shared (one comb per batch) for the batch-static
  // points when sharing is on, per proof otherwise (pv_plan.hpp)
  const bool comb = ctx->prove_comb && n_terms > 0;
  pv_plan plan;
  pv_make_plan(ni, p, k, sd->cons_off, sd->term_point, share, comb, &plan);
  if (comb) pv_make_units(k, sd->cons_off, ctx->prove_piece, &plan);   // at most n_terms + k units
  const std::vector<int32_t>&slot = plan.cons_slot, &term_shared = plan.term_shared, &term_slot = plan.term_slot,
                            &comb_slot_point = plan.comb_slot_point, &comb_shared_point = plan.comb_shared_point;
  const size_t U = comb_slot_point.size(), Us = comb_shared_point.size();
  std::vector<uint8_t> blob(blob_sz, 0);
  memcpy(&blob[o_prefix], prefix_state, 53 * 4);
  if (comb) {
    memcpy(&blob[o_cs], term_slot.data(), (size_t)n_terms * 4);
    if (U) memcpy(&blob[o_cp], comb_slot_point.data(), U * 4);
    if (Us) memcpy(&blob[o_cq], comb_shared_point.data(), Us * 4);
    memcpy(&blob[o_ut], plan.unit_term0.data(), plan.unit_term0.size() * 4);
    memcpy(&blob[o_un], plan.unit_nterms.data(), plan.unit_nterms.size() * 4);
    memcpy(&blob[o_cu], plan.cons_unit0.data(), plan.cons_unit0.size() * 4);
  }
  if (!pool.empty()) memcpy(&blob[o_pool], pool.data(), pool.size());
  if (p) { memcpy(&blob[o_lo], loff.data(), (size_t)p * 4); memcpy(&blob[o_ll], llen.data(), (size_t)p * 4); }
  if (k) { memcpy(&blob[o_lhs], sd->lhs, (size_t)k * 4); memcpy(&blob[o_sl], slot.data(), (size_t)k * 4); }
  if (k) memcpy(&blob[o_co], sd->cons_off, (size_t)(k + 1) * 4);
  if (n_terms) {
    memcpy(&blob[o_ts], sd->term_scalar, (size_t)n_terms * 4);
    memcpy(&blob[o_tp], sd->term_point, (size_t)n_terms * 4);
    memcpy(&blob[o_sh], term_shared.data(), (size_t)n_terms * 4);
  }
  const size_t total = N * (size_t)n_terms, M = N * (size_t)k;
  ENSURE(ctx, S.pv_misc, blob_sz);
  ENSURE(ctx, S.pv_limbs, N * (size_t)p * 160 + 160);
  ENSURE(ctx, S.pv_enc, N * (size_t)p * 32 + 32);
  ENSURE(ctx, S.pv_sec, N * (size_t)m * 32 + 32);
  ENSURE(ctx, S.pv_ent, N * 32);
  ENSURE(ctx, S.pv_state, N * 56 * 4);
  ENSURE(ctx, S.pv_blind, N * (size_t)m * 32 + 32);
  ENSURE(ctx, S.pv_resp, N * (size_t)m * 32 + 32);
  ENSURE(ctx, S.in_scalars, total * 32 + 32);
  ENSURE(ctx, S.niels, total * 128 + 128);
  // Straus tables: one KB per term; combs: one KB per per-proof base; both interleaved in groups of 32 proofs
  ENSURE(ctx, S.tables, ((N + 31) / 32 * 32) * (comb ? U : (size_t)n_terms) * 1024 + 1024);
  ENSURE(ctx, S.sk0, ((N + 31) / 32 * 32) * (size_t)n_terms * 32 + 32);   // group-interleaved when CTA-staged
  ENSURE(ctx, S.aux0, (M + 1) * 8);
  ENSURE(ctx, S.aux1, M * 32 + 32);
  ENSURE(ctx, S.multi, M * 4 + 16);
  ENSURE(ctx, S.flags, 16);
  // comb path, CTA-staged (k_comb_msm_cta): needs its tables and unit sums to fit the shared memory of one SM
  const size_t n_units = plan.unit_term0.size();
  const size_t cta_smem = comb ? comb_cta_smem_bytes(U, Us, n_units) : 0;
  if (comb && ctx->prove_comb >= 2 && ctx->smem_optin < 0)
    CUDA_TRY(ctx, cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  const bool cta = comb && ctx->prove_comb >= 2 && cta_smem <= (size_t)(ctx->smem_optin > 0 ? ctx->smem_optin : 0);
  uint8_t* dm = (uint8_t*)S.pv_misc.p;
  CUDA_TRY(ctx, cudaMemcpyAsync(dm, blob.data(), blob_sz, cudaMemcpyHostToDevice, st));
  if (p) CUDA_TRY(ctx, cudaMemcpyAsync(S.pv_limbs.p, points, N * (size_t)p * 160, cudaMemcpyHostToDevice, st));
  if (m) CUDA_TRY(ctx, cudaMemcpyAsync(S.pv_sec.p, secrets, N * (size_t)m * 32, cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(S.pv_ent.p, entropy, N * 32, cudaMemcpyHostToDevice, st));
  k_init_flags<<<1, 1, 0, st>>>((int*)S.flags.p);
  LAUNCH_CHECK(ctx);
  CUDA_TRY(ctx, cudaMemsetAsync((int*)S.flags.p + 2, 0, 8, st));   // flags[2]: a common point differs between proofs
  pv_desc d;
  d.m = m; d.p = p; d.k = k; d.n_terms = n_terms; d.ni = ni;
  d.term_shared = (const int32_t*)(dm + o_sh);
  d.label_off = (const uint32_t*)(dm + o_lo);
  d.label_len = (const uint32_t*)(dm + o_ll);
  d.labels = dm + o_pool;
  d.lhs = (const int32_t*)(dm + o_lhs);
  d.cons_off = (const int32_t*)(dm + o_co);
  d.term_scalar = (const int32_t*)(dm + o_ts);
  d.term_point = (const int32_t*)(dm + o_tp);
  d.cons_slot = (const int32_t*)(dm + o_sl);
  // the shared combs depend on the points only: built on the auxiliary stream (one thread per base, ~0.7 ms of serial
  // doublings and inversions) under the compressions and the transcript kernel of this slice
  const bool shared_combs_aux = comb && Us > 0;
  if (shared_combs_aux) {
    ENSURE(ctx, S.pv_static, Us * 768 + 256);   // eight affine Niels entries of 96 bytes per shared comb
    CUDA_TRY(ctx, cudaEventRecord(S.ev_fork, st));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->aux_stream, S.ev_fork, 0));
    // from proof 0's copy (k_pv_gather compares every other copy with it)
    k_build_combs<false><<<(unsigned)((Us + 63) / 64), 64, 0, ctx->aux_stream>>>(
        (const unsigned long long*)S.pv_limbs.p, Us, (uint32_t)Us, (uint32_t)p, (const int32_t*)(dm + o_cq),
        (uint4*)S.pv_static.p);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaEventRecord(S.ev_join, ctx->aux_stream));
  }
  // (1) every allocate_point compression (toolbox/mod.rs:180); with shared tables the batch-static points are the same
  // in every proof (checked by k_pv_gather), so proof 0's are compressed and the others get copies
  if (p) {
    if (share) {
      k_compress_limbs_shared<<<(unsigned)((N * ni + (size_t)nc + 255) / 256), 256, 0, st>>>((const unsigned long long*)S.pv_limbs.p, N,
                                                                               (uint32_t)p, (uint32_t)ni, (uint4*)S.pv_enc.p);
      LAUNCH_CHECK(ctx);
      k_replicate_common_enc<<<(unsigned)((N * (size_t)nc + 255) / 256), 256, 0, st>>>((uint4*)S.pv_enc.p, N, (uint32_t)p,
                                                                                       (uint32_t)ni);
      LAUNCH_CHECK(ctx);
    } else {
      k_compress_limbs<<<(unsigned)((N * p + 255) / 256), 256, 0, st>>>((const unsigned long long*)S.pv_limbs.p, N * p,
                                                                        (uint4*)S.pv_enc.p);
      LAUNCH_CHECK(ctx);
    }
  }
  // (2) transcripts up to the commitments + synthetic-nonce blindings (prover.rs:78-89)
  k_pv_blind<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(d, (const uint32_t*)(dm + o_prefix), N, (const uint8_t*)S.pv_enc.p,
                                                          (const uint8_t*)S.pv_sec.p, (const uint8_t*)S.pv_ent.p,
                                                          (uint32_t*)S.pv_state.p, (uint8_t*)S.pv_blind.p,
                                                          (int*)S.flags.p);
  LAUNCH_CHECK(ctx);
  // (3) the N*k constant-time MSMs + compress (prover.rs:93-103)
  if (M) {
    size_t gthreads = total > M ? total : M;
    if (!gthreads) gthreads = 1;
    k_pv_gather<<<(unsigned)((gthreads + 255) / 256), 256, 0, st>>>(d, N, (const unsigned long long*)S.pv_limbs.p,
                                                                    (const uint8_t*)S.pv_blind.p, (uint4*)S.in_scalars.p,
                                                                    (uint4*)S.niels.p, (unsigned long long*)S.aux0.p,
                                                                    (uint32_t*)S.multi.p, share ? (int*)S.flags.p + 2 : nullptr);
    LAUNCH_CHECK(ctx);
    const int32_t* shared_of = share ? d.term_shared : nullptr;
    uint4* shared_tables = nullptr;
    if (comb) {
      // one comb per base, sign-bit recoding of every blinding, then 64 columns per constraint MSM
      if (U) {
        k_build_combs<true><<<(unsigned)((N * U + 63) / 64), 64, 0, st>>>((const unsigned long long*)S.pv_limbs.p, N * U,
                                                                        (uint32_t)U, (uint32_t)p,
                                                                        (const int32_t*)(dm + o_cp), (uint4*)S.tables.p);
        LAUNCH_CHECK(ctx);
      }
      if (shared_combs_aux) CUDA_TRY(ctx, cudaStreamWaitEvent(st, S.ev_join, 0));   // the shared combs are ready
      if (cta) {
        k_comb_recode_il<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint4*)S.in_scalars.p, N, (uint32_t)n_terms,
                                                                         (uint32_t*)S.sk0.p);
        LAUNCH_CHECK(ctx);
        if (cta_smem > ctx->cta_smem_set) {
          CUDA_TRY(ctx, cudaFuncSetAttribute((const void*)k_comb_msm_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cta_smem));
          ctx->cta_smem_set = cta_smem;
        }
        const unsigned nw = n_units < 16 ? (unsigned)n_units : 16u;
        k_comb_msm_cta<<<(unsigned)((N + 31) / 32), 32 * nw, cta_smem, st>>>(
            (const uint32_t*)S.sk0.p, (const uint4*)S.tables.p, (const uint4*)S.pv_static.p,
            (const int32_t*)(dm + o_cs), (const int32_t*)(dm + o_ut), (const int32_t*)(dm + o_un),
            (const int32_t*)(dm + o_cu), N, (uint32_t)n_terms, (uint32_t)U, (uint32_t)Us, (uint32_t)n_units, (uint32_t)k,
            (uint4*)S.aux1.p);
        LAUNCH_CHECK(ctx);
      } else {
      k_comb_recode<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint4*)S.in_scalars.p, total, (uint4*)S.sk0.p);
      LAUNCH_CHECK(ctx);
      k_small_msm_comb<<<(unsigned)((M + 63) / 64), 64, 0, st>>>((const uint32_t*)S.sk0.p, (const uint4*)S.tables.p,
                                                              (const uint4*)S.pv_static.p, (const int32_t*)(dm + o_cs),
                                                              (const unsigned long long*)S.aux0.p,
                                                              (const uint32_t*)S.multi.p, M, (uint32_t)n_terms,
                                                              (uint32_t)U, (uint4*)S.aux1.p);
      LAUNCH_CHECK(ctx);
      }
    } else {
    if (share) {
      // tables of the nc batch-static points, built once from proof 0's copy (all copies were compared above)
      ENSURE(ctx, S.pv_static, (size_t)nc * (128 + 1024 + 32 + 32) + 256);
      uint8_t* ps = (uint8_t*)S.pv_static.p;
      uint4* s_ext = (uint4*)ps;
      shared_tables = (uint4*)(ps + (size_t)nc * 128);
      uint4* s_zero = (uint4*)(ps + (size_t)nc * (128 + 1024));            // dummy scalars (zero: canonical)
      uint4* s_bias = (uint4*)(ps + (size_t)nc * (128 + 1024 + 32));
      CUDA_TRY(ctx, cudaMemsetAsync(s_zero, 0, (size_t)nc * 32, st));
      k_limbs_to_ext<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>((const unsigned long long*)S.pv_limbs.p + (size_t)ni * 20,
                                                                   (size_t)nc, s_ext);
      LAUNCH_CHECK(ctx);
      k_build_tables<false><<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(s_ext, s_zero, (size_t)nc, 1u, shared_tables, s_bias,
                                                                          (int*)S.flags.p + 2);
      LAUNCH_CHECK(ctx);
    }
    if (total) {
      k_build_tables<true><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(
          (const uint4*)S.niels.p, (const uint4*)S.in_scalars.p, total, (uint32_t)n_terms, (uint4*)S.tables.p,
          (uint4*)S.sk0.p, (int*)S.flags.p + 2, shared_of);   // blindings are canonical
      LAUNCH_CHECK(ctx);
    }
    k_small_msm_ct<true><<<(unsigned)((M + 63) / 64), 64, 0, st>>>((const uint32_t*)S.sk0.p, (const uint4*)S.tables.p,
                                                                   (const unsigned long long*)S.aux0.p,
                                                                   (const uint32_t*)S.multi.p, M, (uint32_t)n_terms,
                                                                   (uint4*)S.aux1.p, shared_of, shared_tables);
    LAUNCH_CHECK(ctx);
    }
  }
  // (4) commitments into the transcript, challenge, responses (prover.rs:98-109)
  k_pv_finish<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(d, N, (const uint32_t*)S.pv_state.p, (const uint8_t*)S.aux1.p,
                                                           (const uint8_t*)S.pv_sec.p, (const uint8_t*)S.pv_blind.p,
                                                           (uint8_t*)S.pv_resp.p);
  LAUNCH_CHECK(ctx);
  sl.share = share;
  CUDA_TRY(ctx, cudaMemcpyAsync(S.h_flags, S.flags.p, 16, cudaMemcpyDeviceToHost, st));
  return ZKP_OK;
}

// Outputs leave the device only on the success path: the flags are read first.  (With shared tables a batch whose
// "common" points differ computes commitments from the wrong tables, and its responses s*c + b would share the
// blinding b with the redone call's s*c' + b: two challenges for one nonce reveal s.  Those bytes never reach the caller.)
static int32_t prove_slice_complete(zkp_ctx* ctx, pv_set& S, const zkp_statement_desc* sd, const pv_slice& sl) {
  const int m = sd->m, k = sd->k, p = sd->ni + sd->nc;
  const size_t N = sl.N, M = N * (size_t)k;
  cudaStream_t st = S.stream;
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  if (S.h_flags[1] != 0x7fffffff) return ZKP_ERR_SCALAR;   // a non-canonical secret
  if (sl.share && S.h_flags[2] != 0) return ZKP_PV_NOT_UNIFORM;
  if (p) CUDA_TRY(ctx, cudaMemcpyAsync(sl.encodings_out, S.pv_enc.p, N * (size_t)p * 32, cudaMemcpyDeviceToHost, st));
  if (k) CUDA_TRY(ctx, cudaMemcpyAsync(sl.commitments_out, S.aux1.p, M * 32, cudaMemcpyDeviceToHost, st));
  if (m) CUDA_TRY(ctx, cudaMemcpyAsync(sl.responses_out, S.pv_resp.p, N * (size_t)m * 32, cudaMemcpyDeviceToHost, st));
  if (m && sl.blindings_out)
    CUDA_TRY(ctx, cudaMemcpyAsync(sl.blindings_out, S.pv_blind.p, N * (size_t)m * 32, cudaMemcpyDeviceToHost, st));
  return ZKP_OK;
}

static int32_t prove_sets_ready(zkp_ctx* ctx) {
  for (int i = 0; i < 2; i++) {
    pv_set& S = ctx->pvs[i];
    if (i == 1 && !S.own_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&S.own_stream, cudaStreamNonBlocking));
    S.stream = i == 0 ? ctx->stream : S.own_stream;
    if (!S.h_flags) CUDA_TRY(ctx, cudaMallocHost((void**)&S.h_flags, 64));
    if (!S.ev_fork) {
      CUDA_TRY(ctx, cudaEventCreateWithFlags(&S.ev_fork, cudaEventDisableTiming));
      CUDA_TRY(ctx, cudaEventCreateWithFlags(&S.ev_join, cudaEventDisableTiming));
    }
  }
  if (!ctx->aux_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
  return ZKP_OK;
}

static int32_t prove_batch_run(zkp_ctx* ctx, const zkp_statement_desc* sd, const uint32_t* prefix_state, size_t N,
                               const uint8_t* secrets, const uint64_t* points, const uint8_t* entropy,
                               uint8_t* encodings_out, uint8_t* commitments_out, uint8_t* responses_out,
                               uint8_t* blindings_out, bool share_allowed) {
  const int m = sd->m, k = sd->k, p = sd->ni + sd->nc;
  int32_t rc = prove_sets_ready(ctx);
  if (rc != ZKP_OK) return rc;
  // slices: at most prove_chunk proofs (workspace bound: tables), and at most prove_pipe_chunk when the batch is large
  // enough to be worth pipelining (copies of one slice under the kernels of another)
  size_t chunk = ctx->prove_chunk;
  if (ctx->prove_pipe_chunk && N >= 2 * ctx->prove_pipe_chunk && ctx->prove_pipe_chunk < chunk) {
    // about prove_pipe_chunk proofs per slice, rounded so that a slice is a whole number of waves of the CTA-staged MSM
    // kernel (one CTA = 32 proofs, one CTA per SM): 2^14-proof slices on 148 SMs would run 3.46 waves, 13 % of them idle
    if (ctx->sm_count < 0) CUDA_TRY(ctx, cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, ctx->device));
    const size_t wave = (size_t)32 * (size_t)(ctx->sm_count > 0 ? ctx->sm_count : 1);
    if (ctx->prove_pipe_chunk >= wave) {
      const size_t slices = (N + ctx->prove_pipe_chunk / 2) / ctx->prove_pipe_chunk;          // >= 2
      const size_t waves_total = (N + wave - 1) / wave;
      const size_t waves_per_slice = (waves_total + slices - 1) / slices;
      chunk = waves_per_slice * wave;
    } else {
      chunk = ctx->prove_pipe_chunk;   // slices smaller than a wave: taken as given (tests)
    }
    if (chunk > ctx->prove_chunk) chunk = ctx->prove_chunk;
  }
  pv_slice cur[2];
  bool pending[2] = {false, false};
  auto drain = [&]() {
    for (int i = 0; i < 2; i++)
      if (ctx->pvs[i].stream) cudaStreamSynchronize(ctx->pvs[i].stream);
    if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
  };
  size_t idx = 0;
  for (size_t lo = 0; lo < N; lo += chunk, idx++) {
    const int w = (int)(idx & 1);
    pv_set& S = ctx->pvs[w];
    if (pending[w]) {   // the set is still busy with slice idx - 2: finish it (outputs after the flags)
      rc = prove_slice_complete(ctx, S, sd, cur[w]);
      pending[w] = false;
      if (rc != ZKP_OK) break;
      CUDA_TRY(ctx, cudaStreamSynchronize(S.stream));   // its outputs have left before the buffers are overwritten
    }
    pv_slice& sl = cur[w];
    sl.N = N - lo < chunk ? N - lo : chunk;
    sl.secrets = secrets ? secrets + lo * (size_t)m * 32 : nullptr;
    sl.points = points ? points + lo * (size_t)p * 20 : nullptr;
    sl.entropy = entropy + lo * 32;
    sl.encodings_out = encodings_out ? encodings_out + lo * (size_t)p * 32 : nullptr;
    sl.commitments_out = commitments_out ? commitments_out + lo * (size_t)k * 32 : nullptr;
    sl.responses_out = responses_out ? responses_out + lo * (size_t)m * 32 : nullptr;
    sl.blindings_out = blindings_out ? blindings_out + lo * (size_t)m * 32 : nullptr;
    rc = prove_slice_enqueue(ctx, S, sd, prefix_state, sl, share_allowed);
    if (rc != ZKP_OK) break;
    pending[w] = true;
    if (pending[w ^ 1]) {   // while this slice runs, the previous one's flags are checked and its outputs leave
      rc = prove_slice_complete(ctx, ctx->pvs[w ^ 1], sd, cur[w ^ 1]);
      pending[w ^ 1] = false;
      if (rc != ZKP_OK) break;
    }
  }
  for (int w = 0; w < 2 && rc == ZKP_OK; w++)
    if (pending[w]) {
      rc = prove_slice_complete(ctx, ctx->pvs[w], sd, cur[w]);
      pending[w] = false;
    }
  drain();
  return rc;
}

static int32_t prove_batch_impl(zkp_ctx* ctx, const zkp_statement_desc* sd, const uint32_t* prefix_state, size_t N,
                                const uint8_t* secrets, const uint64_t* points, const uint8_t* entropy,
                                uint8_t* encodings_out, uint8_t* commitments_out, uint8_t* responses_out,
                                uint8_t* blindings_out) {
  if (!ctx || !sd || !prefix_state) return ZKP_ERR_SIZE;
  if (!statement_ok(sd) || sd->ni + sd->nc > 2 * ZKP_BV_MAX_VARS) {
    ctx->err = "inconsistent or oversized statement description";
    return ZKP_ERR_SIZE;
  }
  const int m = sd->m, k = sd->k, p = sd->ni + sd->nc;
  if (!N) return ZKP_OK;
  if ((m && (!secrets || !responses_out)) || (p && (!points || !encodings_out)) || (k && !commitments_out))
    return ZKP_ERR_SIZE;
  std::vector<uint8_t> own_entropy;
  if (!entropy) {   // the caller leaves the 32 bytes per proof of prover.rs:82 (thread_rng) to the OS CSPRNG
    own_entropy.resize(N * 32);
    if (!os_random_bytes(own_entropy.data(), own_entropy.size())) {
      ctx->err = "getrandom failed";
      return ZKP_ERR_CUDA;
    }
    entropy = own_entropy.data();
  }
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  int32_t rc = prove_batch_run(ctx, sd, prefix_state, N, secrets, points, entropy, encodings_out, commitments_out,
                               responses_out, blindings_out, /*share_allowed=*/true);
  if (rc == ZKP_PV_NOT_UNIFORM)   // the "common" points were not common after all: per-proof tables for every term
    rc = prove_batch_run(ctx, sd, prefix_state, N, secrets, points, entropy, encodings_out, commitments_out, responses_out,
                         blindings_out, /*share_allowed=*/false);
  return rc;
}
extern "C" int32_t zkp_prove_batch(zkp_ctx* ctx, const zkp_statement_desc* sd, const uint32_t* prefix_state, size_t N,
                                   const uint8_t* secrets, const uint64_t* points, const uint8_t* entropy,
                                   uint8_t* encodings_out, uint8_t* commitments_out, uint8_t* responses_out,
                                   uint8_t* blindings_out) {
  return guarded(ctx, [&]() -> int32_t {
    return prove_batch_impl(ctx, sd, prefix_state, N, secrets, points, entropy, encodings_out, commitments_out,
                            responses_out, blindings_out);
  });
}

// ---------------------------------------------------------------------------------------------------------
// batched decompress / compress
// ---------------------------------------------------------------------------------------------------------
extern "C" int32_t zkp_decompress_batch(zkp_ctx* ctx, const uint8_t* enc, size_t n, uint64_t* limbs_out,
                                        uint8_t* valid_out) {
  if (!ctx || (n && (!enc || !limbs_out || !valid_out))) return ZKP_ERR_SIZE;
  if (!n) return ZKP_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ENSURE(ctx, ctx->in_points, n * 32);
  ENSURE(ctx, ctx->aux0, n * 160);
  ENSURE(ctx, ctx->aux1, n);
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->in_points.p, enc, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  k_decompress_limbs<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
      (const uint4*)ctx->in_points.p, n, (unsigned long long*)ctx->aux0.p, (uint8_t*)ctx->aux1.p);
  LAUNCH_CHECK(ctx);
  CUDA_TRY(ctx, cudaMemcpyAsync(limbs_out, ctx->aux0.p, n * 160, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(valid_out, ctx->aux1.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKP_OK;
}

extern "C" int32_t zkp_compress_batch(zkp_ctx* ctx, const uint64_t* limbs_in, size_t n, uint8_t* enc_out) {
  if (!ctx || (n && (!limbs_in || !enc_out))) return ZKP_ERR_SIZE;
  if (!n) return ZKP_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ENSURE(ctx, ctx->aux0, n * 160);
  ENSURE(ctx, ctx->in_points, n * 32);
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->aux0.p, limbs_in, n * 160, cudaMemcpyHostToDevice, ctx->stream));
  k_compress_limbs<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const unsigned long long*)ctx->aux0.p, n,
                                                                         (uint4*)ctx->in_points.p);
  LAUNCH_CHECK(ctx);
  CUDA_TRY(ctx, cudaMemcpyAsync(enc_out, ctx->in_points.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// batched small MSMs (CSR)
// ---------------------------------------------------------------------------------------------------------
// MSM indices sorted by descending term count (counting sort): warps then hold MSMs of equal size
static void size_order(const uint64_t* offsets, size_t M, std::vector<uint32_t>* order) {
  const size_t CAP = 4096;
  std::vector<size_t> cnt(CAP + 2, 0);
  auto key = [&](size_t j) { size_t s = (size_t)(offsets[j + 1] - offsets[j]); return s > CAP ? CAP : s; };
  for (size_t j = 0; j < M; j++) cnt[CAP - key(j)]++;
  size_t acc = 0;
  for (size_t b = 0; b <= CAP; b++) { size_t c = cnt[b]; cnt[b] = acc; acc += c; }
  order->resize(M);
  for (size_t j = 0; j < M; j++) (*order)[cnt[CAP - key(j)]++] = (uint32_t)j;
}

static int32_t check_offsets(zkp_ctx* ctx, const uint64_t* offsets, size_t M, size_t* total) {
  if (!offsets) return ZKP_ERR_SIZE;
  if (offsets[0] != 0 || M >= 0xffffffffull) {
    ctx->err = "offsets[0] must be 0 and M < 2^32";
    return ZKP_ERR_SIZE;
  }
  for (size_t j = 0; j < M; j++)
    if (offsets[j + 1] < offsets[j]) {
      ctx->err = "offsets must be non-decreasing";
      return ZKP_ERR_SIZE;
    }
  *total = (size_t)offsets[M];
  return ZKP_OK;
}

extern "C" int32_t zkp_msm_vartime_batched(zkp_ctx* ctx, const uint8_t* scalars, const uint8_t* points,
                                           const uint64_t* offsets, size_t M, uint8_t* out, uint8_t* valid) {
  if (!ctx || (M && (!out || !valid))) return ZKP_ERR_SIZE;
  if (!M) return ZKP_OK;
  size_t total = 0;
  int32_t r = check_offsets(ctx, offsets, M, &total);
  if (r != ZKP_OK) return r;
  if (total && (!scalars || !points)) return ZKP_ERR_SIZE;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ENSURE(ctx, ctx->in_scalars, total * 32 + 32);
  ENSURE(ctx, ctx->in_points, total * 32 + 32);
  ENSURE(ctx, ctx->aux0, (M + 1) * 8);
  ENSURE(ctx, ctx->aux1, M * 32);
  ENSURE(ctx, ctx->aux2, M * 4);
  ENSURE(ctx, ctx->niels, (total + 1) * (16 * ZKP_NIELS_U4));
  ENSURE(ctx, ctx->sk0, total * 32 + 32);
  ENSURE(ctx, ctx->sk1, total * 32 + 32);
  cudaStream_t st = ctx->stream;
  if (total) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->in_scalars.p, scalars, total * 32, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->in_points.p, points, total * 32, cudaMemcpyHostToDevice, st));
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->aux0.p, offsets, (M + 1) * 8, cudaMemcpyHostToDevice, st));
  if (total) {
    const unsigned nb = (unsigned)((total + 255) / 256);
    k_decompress_valid<<<nb, 256, 0, st>>>((const uint4*)ctx->in_points.p, total, (uint4*)ctx->niels.p);
    LAUNCH_CHECK(ctx);
    k_prep_scalars_vt<<<nb, 256, 0, st>>>((const uint4*)ctx->in_scalars.p, total, (uint4*)ctx->sk0.p,
                                          (uint4*)ctx->sk1.p);
    LAUNCH_CHECK(ctx);
  }
  std::vector<uint32_t> order;
  try {
    size_order(offsets, M, &order);
  } catch (...) {
    ctx->err = "host allocation failed";
    return ZKP_ERR_NOMEM;
  }
  ENSURE(ctx, ctx->multi, M * 4 + 16);
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->multi.p, order.data(), M * 4, cudaMemcpyHostToDevice, st));
  if (M <= (size_t)ctx->coop_max_msms)   // too few MSMs to fill the GPU with one thread each: four lanes per MSM (latency)
    k_small_msm_vt<true><<<(unsigned)((4 * M + 63) / 64), 64, 0, st>>>(
        (const uint32_t*)ctx->sk0.p, (const uint32_t*)ctx->sk1.p, (const uint4*)ctx->niels.p,
        (const unsigned long long*)ctx->aux0.p, (const uint32_t*)ctx->multi.p, M, (uint4*)ctx->aux1.p,
        (int*)ctx->aux2.p);
  else
    k_small_msm_vt<false><<<(unsigned)((M + 63) / 64), 64, 0, st>>>(
        (const uint32_t*)ctx->sk0.p, (const uint32_t*)ctx->sk1.p, (const uint4*)ctx->niels.p,
        (const unsigned long long*)ctx->aux0.p, (const uint32_t*)ctx->multi.p, M, (uint4*)ctx->aux1.p,
        (int*)ctx->aux2.p);
  LAUNCH_CHECK(ctx);
  // status words -> valid bytes on the host
  int* hstat = (int*)malloc(M * 4);
  if (!hstat) return ZKP_ERR_NOMEM;
  cudaError_t e1 = cudaMemcpyAsync(out, ctx->aux1.p, M * 32, cudaMemcpyDeviceToHost, st);
  cudaError_t e2 = cudaMemcpyAsync(hstat, ctx->aux2.p, M * 4, cudaMemcpyDeviceToHost, st);
  cudaError_t e3 = cudaStreamSynchronize(st);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
    free(hstat);
    ctx->err = "copy back failed";
    cudaGetLastError();
    return ZKP_ERR_CUDA;
  }
  int32_t ret = ZKP_OK;
  for (size_t j = 0; j < M; j++) {
    valid[j] = hstat[j] == 0 ? 1 : 0;
    if (hstat[j] == 3) ret = ZKP_ERR_SCALAR;
  }
  free(hstat);
  return ret;
}

extern "C" int32_t zkp_msm_ct_batched(zkp_ctx* ctx, const uint8_t* scalars, const void* points, int32_t point_format,
                                      const uint64_t* offsets, size_t M, uint8_t* out) {
  if (!ctx || (M && !out)) return ZKP_ERR_SIZE;
  if (point_format != ZKP_POINTS_COMPRESSED && point_format != ZKP_POINTS_LIMBS51) return ZKP_ERR_SIZE;
  if (!M) return ZKP_OK;
  size_t total = 0;
  int32_t r = check_offsets(ctx, offsets, M, &total);
  if (r != ZKP_OK) return r;
  if (total && (!scalars || !points)) return ZKP_ERR_SIZE;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t psz = point_format == ZKP_POINTS_LIMBS51 ? 160 : 32;
  ENSURE(ctx, ctx->in_scalars, total * 32 + 32);
  ENSURE(ctx, ctx->in_points, total * psz + 160);
  ENSURE(ctx, ctx->aux0, (M + 1) * 8);
  ENSURE(ctx, ctx->aux1, M * 32);
  ENSURE(ctx, ctx->aux2, M * 4);
  ENSURE(ctx, ctx->niels, total * 128 + 128);
  ENSURE(ctx, ctx->flags, 16);
  cudaStream_t st = ctx->stream;
  if (total) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->in_scalars.p, scalars, total * 32, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->in_points.p, points, total * psz, cudaMemcpyHostToDevice, st));
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->aux0.p, offsets, (M + 1) * 8, cudaMemcpyHostToDevice, st));
  k_init_flags<<<1, 1, 0, st>>>((int*)ctx->flags.p);
  LAUNCH_CHECK(ctx);
  if (total) {
    const unsigned nb = (unsigned)((total + 255) / 256);
    if (point_format == ZKP_POINTS_LIMBS51)
      k_limbs_to_ext<<<nb, 256, 0, st>>>((const unsigned long long*)ctx->in_points.p, total, (uint4*)ctx->niels.p);
    else
      k_decompress_ext<<<nb, 256, 0, st>>>((const uint4*)ctx->in_points.p, total, (uint4*)ctx->niels.p,
                                           (int*)ctx->flags.p);
    LAUNCH_CHECK(ctx);
  }
  if (total) {
    ENSURE(ctx, ctx->tables, total * 1024);
    ENSURE(ctx, ctx->sk0, total * 32 + 32);
    k_build_tables<false><<<(unsigned)((total + 127) / 128), 128, 0, st>>>((const uint4*)ctx->niels.p,
                                                                           (const uint4*)ctx->in_scalars.p, total, 1u,
                                                                           (uint4*)ctx->tables.p, (uint4*)ctx->sk0.p,
                                                                           (int*)ctx->flags.p);
    LAUNCH_CHECK(ctx);
  }
  std::vector<uint32_t> order;
  try {
    size_order(offsets, M, &order);
  } catch (...) {
    ctx->err = "host allocation failed";
    return ZKP_ERR_NOMEM;
  }
  ENSURE(ctx, ctx->multi, M * 4 + 16);
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->multi.p, order.data(), M * 4, cudaMemcpyHostToDevice, st));
  if (M <= (size_t)ctx->coop_max_msms)   // few MSMs (a single proof's constraints): four lanes per MSM, latency schedule
    k_small_msm_ct<false, true><<<(unsigned)((4 * M + 63) / 64), 64, 0, st>>>(
        (const uint32_t*)ctx->sk0.p, (const uint4*)ctx->tables.p, (const unsigned long long*)ctx->aux0.p,
        (const uint32_t*)ctx->multi.p, M, 1u, (uint4*)ctx->aux1.p);
  else
    k_small_msm_ct<false, false><<<(unsigned)((M + 63) / 64), 64, 0, st>>>(
        (const uint32_t*)ctx->sk0.p, (const uint4*)ctx->tables.p, (const unsigned long long*)ctx->aux0.p,
        (const uint32_t*)ctx->multi.p, M, 1u, (uint4*)ctx->aux1.p);
  LAUNCH_CHECK(ctx);
  int hflags[4];
  CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->aux1.p, M * 32, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(hflags, ctx->flags.p, 16, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  if (hflags[0] != 0x7fffffff) return ZKP_ERR_POINT;
  if (hflags[1] != 0x7fffffff) return ZKP_ERR_SCALAR;
  return ZKP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// field-multiplier micro-benchmark
// ---------------------------------------------------------------------------------------------------------
struct event_pair {   // scoped events of the diagnostics below
  cudaEvent_t a = nullptr, b = nullptr;
  ~event_pair() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
};

// self-test of the device hashing: Merlin's conformance vector computed on the GPU
extern "C" int32_t zkp_selftest_hash(zkp_ctx* ctx, uint8_t* out32) {
  if (!ctx || !out32) return ZKP_ERR_SIZE;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ENSURE(ctx, ctx->result, 64);
  k_selftest_merlin<<<1, 1, 0, ctx->stream>>>((uint8_t*)ctx->result.p);
  LAUNCH_CHECK(ctx);
  CUDA_TRY(ctx, cudaMemcpyAsync(out32, ctx->result.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKP_OK;
}

#ifdef ZKP_ABLATIONS
// diagnostic: integer squaring warps and FP64 warps side by side; returns milliseconds
extern "C" int32_t zkp_bench_dual(zkp_ctx* ctx, int32_t mode, int32_t iters, double* ms_out) {
  if (!ctx || !ms_out || iters <= 0 || mode < 0 || mode > 2) return ZKP_ERR_SIZE;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  int sms = 0;
  CUDA_TRY(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
  const int threads = 256, blocks = sms * 8;
  ENSURE(ctx, ctx->aux0, (size_t)threads * blocks * 64);
  event_pair evp;
  CUDA_TRY(ctx, cudaEventCreate(&evp.a));
  CUDA_TRY(ctx, cudaEventCreate(&evp.b));
  cudaEvent_t &e0 = evp.a, &e1 = evp.b;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    k_bench_dual<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, iters, mode);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  *ms_out = best;
  return ZKP_OK;
}

#else
extern "C" int32_t zkp_bench_dual(zkp_ctx*, int32_t, int32_t, double*) { return ZKP_ERR_SIZE; }   // ablation build only
#endif

extern "C" int32_t zkp_bench_field(zkp_ctx* ctx, int32_t kind, int32_t iters, double* ops_per_sec) {
  // kinds 17 / 18: squaring / multiplication with the round-1 carry schedule (fe_sq_v1 / fe_mul_v1), variable-time tails
  if (!ctx || !ops_per_sec || iters <= 0 || kind < 0 || kind > 18 || (kind > 11 && kind < 17 && !ZKP_ABL(1))) return ZKP_ERR_SIZE;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  int sms = 0;
  CUDA_TRY(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
  const int threads = 256, blocks = sms * 8;
  ENSURE(ctx, ctx->aux0, (size_t)threads * blocks * 64);
  event_pair evp;
  CUDA_TRY(ctx, cudaEventCreate(&evp.a));
  CUDA_TRY(ctx, cudaEventCreate(&evp.b));
  cudaEvent_t &e0 = evp.a, &e1 = evp.b;
  float best = 1e30f;
  // kinds 12..16 (dynamic pool): `iters` squarings per thread of the usual grid, cut into units of 32 lanes x 256
  const int pool_units = (int)(((long long)threads * blocks / 32) * iters / 256);
  ENSURE(ctx, ctx->aux1, 256);
  for (int rep = 0; rep < 4; rep++) {
    if (kind >= 12 && kind <= 16) CUDA_TRY(ctx, cudaMemsetAsync(ctx->aux1.p, 0, 4, ctx->stream));
    CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    switch (kind) {
      case 0: k_bench_mul32<0, false><<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, iters); break;
      case 1: k_bench_sq32<0, false><<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, iters); break;
      case 6: k_bench_mul32<1, false><<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, iters); break;
      case 7: k_bench_sq32<1, false><<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, iters); break;
      case 8: k_bench_mul32<0, true><<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, iters); break;
      case 9: k_bench_sq32<0, true><<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, iters); break;
      case 10: k_bench_madd<false><<<blocks * 2, threads / 2, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, iters); break;
      case 11: k_bench_madd<true><<<blocks * 2, threads / 2, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, iters); break;
#ifdef ZKP_ABLATIONS
      case 12: k_bench_sq_mixed<0><<<sms * 2, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, pool_units, (unsigned int*)ctx->aux1.p); break;
      case 13: k_bench_sq_mixed<2><<<sms * 2, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, pool_units, (unsigned int*)ctx->aux1.p); break;
      case 14: k_bench_sq_mixed<3><<<sms * 2, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, pool_units, (unsigned int*)ctx->aux1.p); break;
      case 15: k_bench_sq_mixed<4><<<sms * 2, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, pool_units, (unsigned int*)ctx->aux1.p); break;
      case 16: k_bench_sq_mixed<8><<<sms * 2, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, pool_units, (unsigned int*)ctx->aux1.p); break;
#endif
      case 17: k_bench_sq32<0, true, 1><<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, iters); break;
      case 18: k_bench_mul32<0, true, 1><<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, iters); break;
      case 2: k_bench_mul51<<<blocks, threads, 0, ctx->stream>>>((unsigned long long*)ctx->aux0.p, iters); break;
      case 3: k_bench_mul25<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, iters); break;
      case 4: k_bench_wide_plain<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, iters); break;
      case 5: k_bench_wide_carry<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->aux0.p, iters); break;
      default: break;
    }
    LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  if (kind >= 12 && kind <= 16) *ops_per_sec = (double)pool_units * 32.0 * 256.0 / (best * 1e-3);
  else *ops_per_sec = (double)threads * blocks * (double)iters / (best * 1e-3);
  return ZKP_OK;
}
