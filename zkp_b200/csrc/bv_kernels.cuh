// BatchVerifier::verify_batchable front end on the device (SURVEY.md section 8f rows f1 + f2, "next"):
// per-proof Merlin transcripts, challenges, random weights and the coefficient fold of
// /root/reference/src/toolbox/batch_verifier.rs:152-206, producing exactly the scalar / point vectors the reference
// feeds to optional_multiscalar_mul at :219-228 -- which then run through the same MSM kernels (kernels.cuh).
// One thread per proof.  The statement arrives as flat arrays (the define_proof! expansion, macros.rs:336-370).
#pragma once
#include "hash.cuh"
#include "kernels.cuh"
#include "scl.cuh"

namespace zkp {

#define ZKP_BV_MAX_VARS 40   // instance or static point variables per statement (CMZ: 13 / 12)
#define ZKP_BV_MAX_CONS 64   // constraints per statement (CMZ: 11)

struct bv_op {         // one recorded transcript operation after the batch-wide prefix
  uint32_t kind;       // 0 instance point, 1 static point, 2 blinding commitment
  uint32_t label_off, label_len;
  uint32_t idx;        // variable / constraint index
};

struct bv_desc {       // device copy of the statement
  int m, ni, nc, k, n_ops, n_terms;
  const bv_op* ops;                 // [n_ops]
  const uint8_t* labels;            // label byte pool
  const int32_t* lhs_kind;          // [k] 0 instance / 1 static
  const int32_t* lhs_idx;           // [k]
  const int32_t* cons_off;          // [k+1]
  const int32_t* term_scalar;       // [n_terms]
  const int32_t* term_pkind;        // [n_terms]
  const int32_t* term_pidx;         // [n_terms]
};

// Labels are spelled as local byte arrays (no string literals in device code).
#define ZKP_LABEL(name, ...) const uint8_t name[] = {__VA_ARGS__}

__device__ __forceinline__ void load32(uint8_t* dst, const uint8_t* src) {
  const uint4* p = (const uint4*)src;
  uint4 a = __ldg(p), b = __ldg(p + 1);
  uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; i++) {
    dst[4 * i] = (uint8_t)w[i]; dst[4 * i + 1] = (uint8_t)(w[i] >> 8);
    dst[4 * i + 2] = (uint8_t)(w[i] >> 16); dst[4 * i + 3] = (uint8_t)(w[i] >> 24);
  }
}
__device__ __forceinline__ void load_scl(scl& s, const uint8_t* src) {
  const uint4* p = (const uint4*)src;
  uint4 a = __ldg(p), b = __ldg(p + 1);
  s.v[0] = a.x; s.v[1] = a.y; s.v[2] = a.z; s.v[3] = a.w; s.v[4] = b.x; s.v[5] = b.y; s.v[6] = b.z; s.v[7] = b.w;
}
__device__ __forceinline__ void store_scl(uint8_t* dst, const scl& s) {
  uint4* p = (uint4*)dst;
  p[0] = make_uint4(s.v[0], s.v[1], s.v[2], s.v[3]);
  p[1] = make_uint4(s.v[4], s.v[5], s.v[6], s.v[7]);
}

// Weights and coefficient fold of proof j, shared by both front-end kernels:
//   rho_i = bytes [16 i, 16 i + 16) of SHAKE256(seed || le64(j))  (stand-in for thread_rng, batch_verifier.rs:179)
//   fold: batch_verifier.rs:176-206.  inst / stat accumulate the coefficients of the instance / static points.
__device__ __forceinline__ void bv_weights_and_fold(const bv_desc& d, size_t j, size_t N, const scl& minus_c,
                                                    const uint8_t* __restrict__ rho_seed,
                                                    const uint8_t* __restrict__ responses, uint8_t* __restrict__ msm_scalars,
                                                    scl* inst, scl* stat, int* __restrict__ flags) {
  uint8_t msg[40];
#pragma unroll
  for (int b = 0; b < 32; b++) msg[b] = rho_seed[b];
#pragma unroll
  for (int b = 0; b < 8; b++) msg[32 + b] = (uint8_t)((unsigned long long)j >> (8 * b));
  uint8_t all[16 * ZKP_BV_MAX_CONS];
  shake256_short(all, (uint32_t)(16 * d.k), msg, 40);
  for (int i = 0; i < d.k; i++) {
    scl rho;
#pragma unroll
    for (int w = 0; w < 4; w++)
      rho.v[w] = (uint32_t)all[16 * i + 4 * w] | ((uint32_t)all[16 * i + 4 * w + 1] << 8) |
                 ((uint32_t)all[16 * i + 4 * w + 2] << 16) | ((uint32_t)all[16 * i + 4 * w + 3] << 24);
    rho.v[4] = rho.v[5] = rho.v[6] = rho.v[7] = 0;
    scl t;
    scl_neg(t, rho);   // instance_coeffs[(num_i + i, j)] -= rho
    store_scl(msm_scalars + ((size_t)d.nc + (size_t)(d.ni + i) * N + j) * 32, t);
    scl_mul_128(t, rho, minus_c);
    if (d.lhs_kind[i]) scl_add(stat[d.lhs_idx[i]], stat[d.lhs_idx[i]], t);
    else scl_add(inst[d.lhs_idx[i]], inst[d.lhs_idx[i]], t);
    for (int q = d.cons_off[i]; q < d.cons_off[i + 1]; q++) {
      scl resp;
      load_scl(resp, responses + ((size_t)j * d.m + d.term_scalar[q]) * 32);
      if (!scl_is_canonical(resp.v)) atomicMin(&flags[1], (int)j);
      scl_mul_128(t, rho, resp);
      const int pi = d.term_pidx[q];
      if (d.term_pkind[q]) scl_add(stat[pi], stat[pi], t);
      else scl_add(inst[pi], inst[pi], t);
    }
  }
  for (int i = 0; i < d.ni; i++) store_scl(msm_scalars + ((size_t)d.nc + (size_t)i * N + j) * 32, inst[i]);
}
// block partial sums of the static coefficients (all 128 threads of the block take part)
__device__ __forceinline__ void bv_static_block_sums(const bv_desc& d, const scl* stat, scl* red,
                                                     uint8_t* __restrict__ static_part, unsigned block_index) {
  for (int sidx = 0; sidx < d.nc; sidx++) {
    red[threadIdx.x] = stat[sidx];
    __syncthreads();
    for (int off = 64; off >= 1; off >>= 1) {
      if ((int)threadIdx.x < off) scl_add(red[threadIdx.x], red[threadIdx.x], red[threadIdx.x + off]);
      __syncthreads();
    }
    if (threadIdx.x == 0) store_scl(static_part + ((size_t)block_index * d.nc + sidx) * 32, red[0]);
    __syncthreads();
  }
}

#ifdef ZKP_ABLATIONS   // byte-wise STROBE on the device: the cross-check of k_bv_prepare2 (api.cu: bv_compiled = 0)
// Per proof j: replay the transcript, derive the challenge, draw the weights, fold the coefficients.
//   prefix        strobe state shared by all proofs (user transcript + dom-sep + scalar labels), 53 words
//   instance_enc  [ni][N][32]   commitments [N][k][32]   responses [N][m][32]   common_enc [nc][32]
//   msm_scalars   [nc + (ni+k)*N][32]  (static part written by k_bv_static_sum)
//   msm_points    [nc + (ni+k)*N][32]  (static + instance rows are copied by the host; commitment rows written here)
//   static_part   [gridDim.x][nc][32]  per-block partial sums of the static coefficients
//   flags[2]      set to the first proof index with an identity encoding or a non-canonical response
__global__ void __launch_bounds__(128) k_bv_prepare(bv_desc d, const uint32_t* __restrict__ prefix, size_t N,
                                                    const uint8_t* __restrict__ instance_enc,
                                                    const uint8_t* common_enc,
                                                    const uint8_t* __restrict__ commitments,
                                                    const uint8_t* __restrict__ responses,
                                                    const uint8_t* __restrict__ rho_seed, uint8_t* __restrict__ msm_scalars,
                                                    uint8_t* msm_points, uint8_t* __restrict__ static_part,
                                                    uint8_t* __restrict__ minus_c_out, int* __restrict__ flags,
                                                    size_t j0, size_t cnt, unsigned block_base) {
  // proofs [j0, j0 + cnt) of the batch (chunked ingestion: H2D of the next chunk overlaps this kernel)
  __shared__ scl red[128];
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t j = j0 + tid;
  const bool live = tid < cnt;
  scl inst[ZKP_BV_MAX_VARS], stat[ZKP_BV_MAX_VARS];
  for (int i = 0; i < d.ni; i++) scl_zero(inst[i]);
  for (int i = 0; i < d.nc; i++) scl_zero(stat[i]);
  if (live) {
    // ---- transcript (batch_verifier.rs:115-134 instance points, :100-112 static points, :152-160 commitments) ----
    strobe_t s;
#pragma unroll
    for (int i = 0; i < 25; i++) s.st[i] = (uint64_t)prefix[2 * i] | ((uint64_t)prefix[2 * i + 1] << 32);
    s.pos = prefix[50]; s.pos_begin = prefix[51]; s.cur_flags = prefix[52];
    ZKP_LABEL(L_BLINDCOM, 'b', 'l', 'i', 'n', 'd', 'c', 'o', 'm');
    ZKP_LABEL(L_PTVAR, 'p', 't', 'v', 'a', 'r');
    ZKP_LABEL(L_VAL, 'v', 'a', 'l');
    ZKP_LABEL(L_CHAL, 'c', 'h', 'a', 'l');
    uint8_t enc[32];
    bool bad = false;
    for (int o = 0; o < d.n_ops; o++) {
      const bv_op op = d.ops[o];
      const uint8_t* src = op.kind == 0 ? instance_enc + ((size_t)op.idx * N + j) * 32
                         : op.kind == 1 ? common_enc + (size_t)op.idx * 32
                                        : commitments + ((size_t)j * d.k + op.idx) * 32;
      load32(enc, src);
      uint32_t any = 0;
#pragma unroll
      for (int b = 0; b < 32; b++) any |= enc[b];
      if (any == 0) bad = true;   // identity encoding (toolbox/mod.rs:191, :215)
      if (op.kind == 2) {
        transcript_append(s, L_BLINDCOM, 8, d.labels + op.label_off, op.label_len);
        uint4* dst = (uint4*)(msm_points + ((size_t)d.nc + (size_t)(d.ni + op.idx) * N + j) * 32);
        const uint4* s4 = (const uint4*)src;
        dst[0] = __ldg(s4);
        dst[1] = __ldg(s4 + 1);
      } else {
        transcript_append(s, L_PTVAR, 5, d.labels + op.label_off, op.label_len);
      }
      transcript_append(s, L_VAL, 3, enc, 32);
    }
    if (bad) atomicMin(&flags[0], (int)j);
    uint8_t wide[64];
    transcript_challenge(s, L_CHAL, 4, wide, 64);
    scl c, minus_c;
    scl_from_wide(c, wide);
    scl_neg(minus_c, c);
    if (minus_c_out) store_scl(minus_c_out + j * 32, minus_c);
    bv_weights_and_fold(d, j, N, minus_c, rho_seed, responses, msm_scalars, inst, stat, flags);
  }
  bv_static_block_sums(d, stat, red, static_part, block_base + blockIdx.x);
}

#endif  // ZKP_ABLATIONS

// -----------------------------------------------------------------------------------------------------------------
// The same front end with the transcript COMPILED on the host.  Every proof of a batch follows the same script (same
// labels, same lengths), so STROBE's framing -- operation headers, length fields, labels, the batch-wide static point
// encodings, the padding bytes of every run_f and the block boundaries -- is identical for all proofs and only the 32-byte
// per-proof values (instance encodings, commitments) differ.  The host runs a symbolic STROBE over the script
// (api.cu bv_compile_script) and hands over
//   tmpl[nblocks][21]   the constant bytes of every 168-byte rate block (166 rate bytes + the two padding positions)
//   segs[]              where the bytes [src_off, src_off + len) of per-proof value (kind, idx) land in which block
// and a thread absorbs block b as  state[0..20] ^= tmpl[b] ^ (its values shifted into place), Keccak-f in registers:
// no byte loops, no local-memory state (the byte-wise k_bv_prepare spends ~4x its instruction-issue time there).
// The challenge is the first 64 bytes of the state after the last block (the script ends with the forced run_f of `prf`).
// -----------------------------------------------------------------------------------------------------------------
struct bv_seg {
  uint32_t kind;      // 0 instance point, 2 blinding commitment
  uint32_t idx;       // variable / constraint index
  uint32_t src_off;   // first byte of the value used by this segment
  uint32_t len;       // bytes
  int32_t shift;      // dst_off - src_off: byte position of value byte 0 in the block (may be negative)
  uint32_t pad[3];
};

__global__ void __launch_bounds__(128) k_bv_prepare2(bv_desc d, const uint32_t* __restrict__ prefix, size_t N,
                                                     const uint8_t* __restrict__ instance_enc,
                                                     const uint8_t* __restrict__ commitments,
                                                     const uint8_t* __restrict__ responses,
                                                     const uint8_t* __restrict__ rho_seed, int nblocks,
                                                     const unsigned long long* __restrict__ tmpl,
                                                     const uint32_t* __restrict__ seg_start,   // [nblocks + 1]
                                                     const bv_seg* __restrict__ segs, uint8_t* __restrict__ msm_scalars,
                                                     uint8_t* __restrict__ msm_points, uint8_t* __restrict__ static_part,
                                                     uint8_t* __restrict__ minus_c_out, int* __restrict__ flags, size_t j0,
                                                     size_t cnt, unsigned block_base) {
  __shared__ unsigned long long buf[21][128];   // the rate block under construction, one column per thread
  __shared__ scl red[128];
  const int tx = threadIdx.x;
  scl inst[ZKP_BV_MAX_VARS], stat[ZKP_BV_MAX_VARS];
  for (int i = 0; i < d.nc; i++) scl_zero(stat[i]);
  // a block takes the 128-proof groups blockIdx.x, blockIdx.x + gridDim.x, ...: the grid may be smaller than the slab (a
  // resident set next to the previous slab's decompression, api.cu); the static coefficients accumulate over the groups
  for (size_t vb = blockIdx.x; vb * 128 < cnt; vb += gridDim.x) {
  const size_t tid = vb * 128 + threadIdx.x;
  const size_t j = j0 + tid;
  const bool live = tid < cnt;
  for (int i = 0; i < d.ni; i++) scl_zero(inst[i]);
  if (live) {
    uint64_t st[25];
#pragma unroll
    for (int i = 0; i < 25; i++) st[i] = (uint64_t)prefix[2 * i] | ((uint64_t)prefix[2 * i + 1] << 32);
    bool bad = false;
    for (int b = 0; b < nblocks; b++) {
#pragma unroll
      for (int l = 0; l < 21; l++) buf[l][tx] = __ldg(tmpl + (size_t)b * 21 + l);
      for (uint32_t si = seg_start[b]; si < seg_start[b + 1]; si++) {
        const bv_seg sg = segs[si];
        const uint8_t* src = sg.kind == 0 ? instance_enc + ((size_t)sg.idx * N + j) * 32
                                          : commitments + ((size_t)j * d.k + sg.idx) * 32;
        const uint4 a = __ldg((const uint4*)src), c = __ldg((const uint4*)src + 1);
        uint64_t w[4] = {(uint64_t)a.x | ((uint64_t)a.y << 32), (uint64_t)a.z | ((uint64_t)a.w << 32),
                         (uint64_t)c.x | ((uint64_t)c.y << 32), (uint64_t)c.z | ((uint64_t)c.w << 32)};
        if (sg.src_off == 0) {   // first (or only) piece of this value: the checks and copies done once per value
          if ((w[0] | w[1] | w[2] | w[3]) == 0) bad = true;   // identity encoding (toolbox/mod.rs:191, :215)
          if (sg.kind == 2) {
            uint4* dst = (uint4*)(msm_points + ((size_t)d.nc + (size_t)(d.ni + sg.idx) * N + j) * 32);
            dst[0] = a;
            dst[1] = c;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          // bytes of word q inside [src_off, src_off + len)
          const int lo = (int)sg.src_off - 8 * q, hi = (int)(sg.src_off + sg.len) - 8 * q;   // range within the word
          if (hi <= 0 || lo >= 8) continue;
          uint64_t v = w[q];
          if (lo > 0) v &= ~0ULL << (8 * lo);
          if (hi < 8) v &= ~0ULL >> (8 * (8 - hi));
          const int pos = sg.shift + 8 * q;            // byte position of this word's byte 0 (may be negative)
          const int lane = pos >> 3, r = pos & 7;      // arithmetic shift: floor division
          if (r == 0) {
            if (lane >= 0 && lane < 21) buf[lane][tx] ^= v;
          } else {
            if (lane >= 0 && lane < 21) buf[lane][tx] ^= v << (8 * r);
            if (lane + 1 >= 0 && lane + 1 < 21) buf[lane + 1][tx] ^= v >> (64 - 8 * r);
          }
        }
      }
#pragma unroll
      for (int l = 0; l < 21; l++) st[l] ^= buf[l][tx];
      keccak_f1600_dev(st);
    }
    if (bad) atomicMin(&flags[0], (int)j);
    uint8_t wide[64];
#pragma unroll
    for (int i = 0; i < 64; i++) wide[i] = (uint8_t)(st[i >> 3] >> (8 * (i & 7)));
    scl c, minus_c;
    scl_from_wide(c, wide);
    scl_neg(minus_c, c);
    if (minus_c_out) store_scl(minus_c_out + j * 32, minus_c);
    bv_weights_and_fold(d, j, N, minus_c, rho_seed, responses, msm_scalars, inst, stat, flags);
  }
  }
  bv_static_block_sums(d, stat, red, static_part, block_base + blockIdx.x);
}

// self-test: Merlin's published conformance vector computed by one device thread (zkp_selftest_hash)
__global__ void k_selftest_merlin(uint8_t* out32) {
  strobe_t s;
  for (int i = 0; i < 25; i++) s.st[i] = 0;
  const uint8_t init[18] = {1, 168, 1, 0, 1, 96, 'S', 'T', 'R', 'O', 'B', 'E', 'v', '1', '.', '0', '.', '2'};
  for (int i = 0; i < 18; i++) st_bytes(s)[i] ^= init[i];
  keccak_f1600_dev(s.st);
  s.pos = 0; s.pos_begin = 0; s.cur_flags = 0;
  ZKP_LABEL(L_MERLIN, 'M', 'e', 'r', 'l', 'i', 'n', ' ', 'v', '1', '.', '0');
  ZKP_LABEL(L_DOMSEP, 'd', 'o', 'm', '-', 's', 'e', 'p');
  ZKP_LABEL(L_PROTO, 't', 'e', 's', 't', ' ', 'p', 'r', 'o', 't', 'o', 'c', 'o', 'l');
  ZKP_LABEL(L_STEP1, 's', 't', 'e', 'p', '1');
  ZKP_LABEL(L_SOME, 's', 'o', 'm', 'e', ' ', 'd', 'a', 't', 'a');
  ZKP_LABEL(L_CHALLENGE, 'c', 'h', 'a', 'l', 'l', 'e', 'n', 'g', 'e');
  ZKP_LABEL(L_BIGDATA, 'b', 'i', 'g', 'd', 'a', 't', 'a');
  ZKP_LABEL(L_CHDATA, 'c', 'h', 'a', 'l', 'l', 'e', 'n', 'g', 'e', 'd', 'a', 't', 'a');
  strobe_meta_ad(s, L_MERLIN, 11, false);
  transcript_append(s, L_DOMSEP, 7, L_PROTO, 13);
  transcript_append(s, L_STEP1, 5, L_SOME, 9);
  uint8_t ch[32], big[64];
  for (int i = 0; i < 64; i++) big[i] = 0x63;
  for (int r = 0; r < 32; r++) {
    transcript_challenge(s, L_CHALLENGE, 9, ch, 32);
    // append_message("bigdata", 1024 x 0x63) in 64-byte pieces
    uint8_t l4[4] = {0, 4, 0, 0};
    strobe_meta_ad(s, L_BIGDATA, 7, false);
    strobe_meta_ad(s, l4, 4, true);
    strobe_ad(s, big, 64, false);
    for (int q = 1; q < 16; q++) strobe_ad(s, big, 64, true);
    transcript_append(s, L_CHDATA, 13, ch, 32);
  }
  for (int i = 0; i < 32; i++) out32[i] = ch[i];
}

// static_coeffs[s] = sum over blocks of static_part[b][s]; one block, thread per (strided) partial
__global__ void __launch_bounds__(256) k_bv_static_sum(const uint8_t* __restrict__ static_part, int nblocks, int nc,
                                                       uint8_t* __restrict__ msm_scalars) {
  __shared__ scl red[256];
  for (int sidx = 0; sidx < nc; sidx++) {
    scl acc, t;
    scl_zero(acc);
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
      load_scl(t, static_part + ((size_t)b * nc + sidx) * 32);
      scl_add(acc, acc, t);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off >= 1; off >>= 1) {
      if ((int)threadIdx.x < off) scl_add(red[threadIdx.x], red[threadIdx.x], red[threadIdx.x + off]);
      __syncthreads();
    }
    if (threadIdx.x == 0) store_scl(msm_scalars + (size_t)sidx * 32, red[0]);
    __syncthreads();
  }
}

}  // namespace zkp
