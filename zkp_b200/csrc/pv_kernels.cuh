// Prover front end on the device: N independent proofs of one statement, everything of
// /root/reference/src/toolbox/prover.rs:76-112 (prove_impl) per proof on the GPU --
//   k_pv_blind    transcript replay up to the commitments (allocate_point appends, toolbox/mod.rs:175-184), the synthetic
//                 nonce generator of prover.rs:78-89 (merlin TranscriptRng: rekey with every witness, finalize with 32
//                 bytes of caller entropy, 64 PRF bytes per blinding reduced mod l)
//   k_pv_gather   the (scalar, point) vectors of the k constraint MSMs of every proof (prover.rs:93-97), CSR offsets
//   (k_build_tables + k_small_msm_ct of small_msm.cuh: the constant-time MSMs + compress, prover.rs:94-100)
//   k_pv_finish   commitments into the transcript, challenge, responses s*c + b (prover.rs:98-109)
// SURVEY.md section 8f-1 for the prover: with the MSMs on the GPU the host's Merlin hashing (~45 Keccak-f per CMZ proof)
// is the limit of prove_many.  One thread per proof; secrets never select a branch or an address here either (STROBE
// and the scalar arithmetic of scl.cuh are straight-line in the data).
#pragma once
#include "bv_kernels.cuh"
#include "small_msm.cuh"

namespace zkp {

struct pv_desc {       // device copy of the statement (points indexed over instance ++ common, as allocated)
  int m, p, k, n_terms, ni;
  const int32_t* term_shared;       // [n_terms] index of the common point a term uses, -1 for an instance point
  const uint32_t* label_off;        // [p]
  const uint32_t* label_len;        // [p]
  const uint8_t* labels;            // label byte pool
  const int32_t* lhs;               // [k] point index of the left-hand side
  const int32_t* cons_off;          // [k+1]
  const int32_t* term_scalar;       // [n_terms]
  const int32_t* term_point;        // [n_terms]
  const int32_t* cons_slot;         // [k] position of the constraint in the size-ordered MSM schedule
};

__device__ __forceinline__ void pv_load_state(strobe_t& s, const uint32_t* __restrict__ w) {
#pragma unroll
  for (int i = 0; i < 25; i++) s.st[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
  s.pos = w[50]; s.pos_begin = w[51]; s.cur_flags = w[52];
}
__device__ __forceinline__ void pv_store_state(uint32_t* __restrict__ w, const strobe_t& s) {
#pragma unroll
  for (int i = 0; i < 25; i++) { w[2 * i] = (uint32_t)s.st[i]; w[2 * i + 1] = (uint32_t)(s.st[i] >> 32); }
  w[50] = s.pos; w[51] = s.pos_begin; w[52] = s.cur_flags;
}

//   prefix     strobe state shared by all proofs (user transcript + dom-sep + scalar labels), 53 words
//   enc        [N][p][32] encodings of the proof's points (k_compress_limbs)      secrets [N][m][32] (canonical)
//   entropy    [N][32]     states [N][56] words (transcript after the point appends)     blind [N][m][32]
__global__ void __launch_bounds__(128) k_pv_blind(pv_desc d, const uint32_t* __restrict__ prefix, size_t N,
                                                  const uint8_t* __restrict__ enc, const uint8_t* __restrict__ secrets,
                                                  const uint8_t* __restrict__ entropy, uint32_t* __restrict__ states,
                                                  uint8_t* __restrict__ blind, int* __restrict__ flags) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  strobe_t s;
  pv_load_state(s, prefix);
  ZKP_LABEL(L_PTVAR, 'p', 't', 'v', 'a', 'r');
  ZKP_LABEL(L_VAL, 'v', 'a', 'l');
  uint8_t buf[64];
  for (int i = 0; i < d.p; i++) {
    load32(buf, enc + ((size_t)j * d.p + i) * 32);
    transcript_append(s, L_PTVAR, 5, d.labels + d.label_off[i], d.label_len[i]);
    transcript_append(s, L_VAL, 3, buf, 32);
  }
  pv_store_state(states + j * 56, s);
  // TranscriptRng: clone of the transcript, rekeyed with every witness, finalized with the caller's entropy
  for (int i = 0; i < d.m; i++) {
    scl w;
    load_scl(w, secrets + ((size_t)j * d.m + i) * 32);
    if (!scl_is_canonical(w.v)) atomicMin(&flags[1], (int)j);
    load32(buf, secrets + ((size_t)j * d.m + i) * 32);
    rng_rekey_with_witness(s, buf, 0, buf, 32);
  }
  load32(buf, entropy + j * 32);
  rng_finalize(s, buf);
  for (int i = 0; i < d.m; i++) {
    rng_fill_bytes(s, buf, 64);
    scl b;
    scl_from_wide(b, buf);
    store_scl(blind + ((size_t)j * d.m + i) * 32, b);
  }
}

// allocate_point compressions of a batch whose common points (indices ni .. p-1) are the same in every proof: the N * ni
// instance points, one per thread with no idle lanes in between (thread t = point t % ni of proof t / ni), then the p - ni
// common points of proof 0 ...  (Round 1 mapped thread t to point t % p of proof t / p and let the common points' threads
// return: 12 of every 25 lanes of a warp idled through the 254-squaring encode -- 0.91 ms per 18 944 CMZ proofs against
// 0.48 ms now.)
__global__ void __launch_bounds__(256) k_compress_limbs_shared(const unsigned long long* __restrict__ limbs, size_t N,
                                                               uint32_t p, uint32_t ni, uint4* __restrict__ enc) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n_inst = N * ni;
  if (t >= n_inst + (p - ni)) return;
  const size_t j = t < n_inst ? t / ni : 0;
  const uint32_t i = t < n_inst ? (uint32_t)(t % ni) : ni + (uint32_t)(t - n_inst);
  const size_t at = j * p + i;
  ge_ext pt;
  load_ext_limbs51(pt, limbs + 20 * at);
  uint32_t w[8];
  ristretto_encode(w, pt);
  enc[2 * at] = make_uint4(w[0], w[1], w[2], w[3]);
  enc[2 * at + 1] = make_uint4(w[4], w[5], w[6], w[7]);
}
// ... and those get proof 0's encodings
__global__ void __launch_bounds__(256) k_replicate_common_enc(uint4* __restrict__ enc, size_t N, uint32_t p, uint32_t ni) {
  const uint32_t nc = p - ni;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * nc) return;
  const size_t j = t / nc;
  const uint32_t i = ni + (uint32_t)(t % nc);
  if (j == 0) return;
  enc[2 * (j * p + i)] = enc[2 * (size_t)i];
  enc[2 * (j * p + i) + 1] = enc[2 * (size_t)i + 1];
}

// one thread per term (j, q): flat scalar and extended point of the constraint MSMs; one thread per (j, c): CSR + order
__global__ void __launch_bounds__(256) k_pv_gather(pv_desc d, size_t N, const unsigned long long* __restrict__ limbs,
                                                   const uint8_t* __restrict__ blind, uint4* __restrict__ scalars_flat,
                                                   uint4* __restrict__ ext_flat, unsigned long long* __restrict__ offsets,
                                                   uint32_t* __restrict__ order, int* __restrict__ not_uniform) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t T = (size_t)d.n_terms;
  if (t < N * T) {
    const size_t j = t / T;
    const int q = (int)(t % T);
    const uint4* src = (const uint4*)(blind + ((size_t)j * d.m + d.term_scalar[q]) * 32);
    scalars_flat[2 * t] = src[0];
    scalars_flat[2 * t + 1] = src[1];
    const unsigned long long* lp = limbs + ((size_t)j * d.p + d.term_point[q]) * 20;
    if (not_uniform && d.term_shared[q] >= 0) {
      // a batch-static base: its table is shared, so this proof's copy must equal proof 0's (public data)
      const unsigned long long* l0 = limbs + (size_t)d.term_point[q] * 20;
      unsigned long long diff = 0;
#pragma unroll
      for (int i = 0; i < 20; i++) diff |= lp[i] ^ l0[i];
      if (diff) atomicExch(not_uniform, 1);
    }
    ge_ext pt;
    load_ext_limbs51(pt, lp);
    store_ext(ext_flat + 8 * t, pt);
  }
  if (t < N * (size_t)d.k) {
    const size_t j = t / d.k;
    const int c = (int)(t % d.k);
    offsets[t] = j * T + (size_t)d.cons_off[c];
    order[(size_t)d.cons_slot[c] * N + j] = (uint32_t)t;
  }
  if (t == 0) offsets[N * (size_t)d.k] = N * T;
}

//   com [N][k][32] (k_small_msm_ct)   ->   resp [N][m][32]
__global__ void __launch_bounds__(128) k_pv_finish(pv_desc d, size_t N, const uint32_t* __restrict__ states,
                                                   const uint8_t* __restrict__ com, const uint8_t* __restrict__ secrets,
                                                   const uint8_t* __restrict__ blind, uint8_t* __restrict__ resp) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  strobe_t s;
  pv_load_state(s, states + j * 56);
  ZKP_LABEL(L_BLINDCOM, 'b', 'l', 'i', 'n', 'd', 'c', 'o', 'm');
  ZKP_LABEL(L_VAL, 'v', 'a', 'l');
  ZKP_LABEL(L_CHAL, 'c', 'h', 'a', 'l');
  uint8_t buf[64];
  for (int c = 0; c < d.k; c++) {
    const int l = d.lhs[c];
    load32(buf, com + ((size_t)j * d.k + c) * 32);
    transcript_append(s, L_BLINDCOM, 8, d.labels + d.label_off[l], d.label_len[l]);
    transcript_append(s, L_VAL, 3, buf, 32);
  }
  transcript_challenge(s, L_CHAL, 4, buf, 64);
  scl ch;
  scl_from_wide(ch, buf);
  for (int i = 0; i < d.m; i++) {
    scl w, b, r;
    load_scl(w, secrets + ((size_t)j * d.m + i) * 32);
    load_scl(b, blind + ((size_t)j * d.m + i) * 32);
    scl_mul(r, w, ch);
    scl_add(r, r, b);
    store_scl(resp + ((size_t)j * d.m + i) * 32, r);
  }
}
}  // namespace zkp
