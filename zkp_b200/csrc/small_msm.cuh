// Batched small MSMs: one thread per MSM, M independent MSMs described CSR-style.
//
//   k_small_msm_vt   variable time; replaces Straus::optional_multiscalar_mul (curve25519-dalek 2.x
//                    backend/serial/scalar_mul/straus.rs [ext], sizes < 190) as called per constraint at
//                    /root/reference/src/toolbox/verifier.rs:97-106 and per proof at verifier.rs:162-166.
//                    Signed binary (NAF) digits taken from k and 3k, no tables: registers only.
//   k_small_msm_ct   constant time; replaces Straus::multiscalar_mul (radix-16 signed digits, table of
//                    1P..8P, full-table scan) as called per constraint at /root/reference/src/toolbox/prover.rs:94-97,
//                    fused with the compress of append_blinding_commitment (toolbox/mod.rs:204).
//                    No branch and no address depends on a scalar: digits come from k + 0x88..8 (so that
//                    nibble - 8 is the signed digit, no carry chain), every table entry is read for every
//                    digit and the selection is arithmetic masking; checked in SASS.
// Outputs are ristretto encodings, so they are byte-identical to the reference's whatever the digit set.
#pragma once
#include "comb.cuh"
#include "kernels.cuh"

namespace zkp {

// ---- preparation passes (one thread per term) -------------------------------------------------------------

// decompress to affine Niels; an invalid encoding is marked by an all-zero first element (y+x is never 0 on
// the curve)
__global__ void __launch_bounds__(256, 3) k_decompress_valid(const uint4* __restrict__ enc, size_t n,
                                                          uint4* __restrict__ niels) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t w[8];
  load_words8(w, enc + 2 * i);
  fe x, y, t;
  uint32_t ok = ristretto_decode(x, y, t, w);
  ge_aniels q;
  ge_aniels_from_affine(q, x, y, t);
  if (!ok) fe_zero(q.yplusx);
  uint4* o = niels + ZKP_NIELS_U4 * i;
  store_fe(o, q.yplusx);
  store_fe(o + 2, q.yminusx);
  store_fe(o + 4, q.xy2d);
}

// vartime scalars: kk = min(s, l-s) with the fold sign in bit 255, k3 = 3*kk; non-canonical -> kk word7 = ~0
__global__ void __launch_bounds__(256) k_prep_scalars_vt(const uint4* __restrict__ scalars, size_t n,
                                                         uint4* __restrict__ kk, uint4* __restrict__ k3) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[8], k[8], t[8], u[8];
  load_words8(s, scalars + 2 * i);
  uint32_t neg;
  uint32_t canonical = sc_fold_sign(k, neg, s);
  add8(t, k, k);
  add8(u, t, k);  // 3k < 2^254
  k[7] |= neg << 31;
  if (!canonical) k[7] = 0xffffffffu;
  kk[2 * i] = make_uint4(k[0], k[1], k[2], k[3]);
  kk[2 * i + 1] = make_uint4(k[4], k[5], k[6], k[7]);
  k3[2 * i] = make_uint4(u[0], u[1], u[2], u[3]);
  k3[2 * i + 1] = make_uint4(u[4], u[5], u[6], u[7]);
}

// COOP: four adjacent lanes share one MSM (ge_double_coop4 / ge_madd_coop4 of kernels.cuh): ~2.5x shorter latency per
// MSM for 4x the threads -- used when the batch is too small to fill the GPU with one thread per MSM (a single proof's
// verification is 254 doublings in a row).
template <bool COOP>
__global__ void __launch_bounds__(64) k_small_msm_vt(const uint32_t* __restrict__ kk, const uint32_t* __restrict__ k3,
                                                     const uint4* __restrict__ niels,
                                                     const unsigned long long* __restrict__ offsets,
                                                     const uint32_t* __restrict__ order, size_t M,
                                                     uint4* __restrict__ out, int* __restrict__ status) {
  const size_t thread = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t tid = COOP ? thread >> 2 : thread;
  const int sub = COOP ? (int)(threadIdx.x & 3) : 0;
  const int gbase = (int)(threadIdx.x & 31 & ~3);
  const unsigned gmask = 0xfu << gbase;
  if (tid >= M) return;          // the four lanes of a group leave together
  const size_t j = order[tid];   // MSMs sorted by size so that the lanes of a warp do equal work
  const size_t lo = offsets[j], hi = offsets[j + 1];
  int st = 0;
  for (size_t t = lo; t < hi; t++) {
    uint4 first = __ldg(niels + ZKP_NIELS_U4 * t), second = __ldg(niels + ZKP_NIELS_U4 * t + 1);
    if ((first.x | first.y | first.z | first.w | second.x | second.y | second.z | second.w) == 0) st = 1;
    if (__ldg(kk + 8 * t + 7) == 0xffffffffu && st == 0) st = 3;
  }
  if (st != 0) {
    if (sub == 0) {
      status[j] = st;
      out[2 * j] = make_uint4(0, 0, 0, 0);
      out[2 * j + 1] = make_uint4(0, 0, 0, 0);
    }
    return;
  }
  ge_ext acc;
  ge_identity(acc);
  bool started = false;
  // digit_i = bit_{i+1}(3k) - bit_{i+1}(k), i = 253 .. 0   (3k < 2^254)
  for (int i = 253; i >= 0; i--) {
    if (started) {
      if (COOP) ge_double_coop4(acc, sub, gmask, gbase);
      else ge_double(acc, acc);
    }
    const int b = i + 1, wi = b >> 5, sh = b & 31;
    for (size_t t = lo; t < hi; t++) {
      uint32_t w3 = __ldg(k3 + 8 * t + wi), w1 = __ldg(kk + 8 * t + wi);
      if (wi == 7) w1 &= 0x7fffffffu;
      int d = (int)((w3 >> sh) & 1u) - (int)((w1 >> sh) & 1u);
      if (d != 0) {
        uint32_t neg = (__ldg(kk + 8 * t + 7) >> 31) ^ (d < 0 ? 1u : 0u);
        ge_aniels q;
        load_aniels(q, niels, (uint32_t)t);
        ge_aniels_cneg(q, neg);
        if (COOP) ge_madd_coop4(acc, q, sub, gmask, gbase);
        else ge_madd(acc, acc, q);
        started = true;
      }
    }
  }
  if (sub != 0) return;
  uint32_t enc[8];
  ristretto_encode(enc, acc);
  out[2 * j] = make_uint4(enc[0], enc[1], enc[2], enc[3]);
  out[2 * j + 1] = make_uint4(enc[4], enc[5], enc[6], enc[7]);
  status[j] = 0;
}

// ---- one small MSM (the reference's n < 190 Straus dispatch: verifier.rs:162-166 -> dalek edwards.rs [ext]) ----------
// zkp_msm_vartime / zkp_batch_verify with few terms must not pay for the sort pipeline (21 launches, bucket reduction
// levels): G groups of four lanes take the terms g, g + G, ... each (doublings shared by the four lanes as in
// k_small_msm_vt<true>), write their partial sums, and k_single_finish adds them, encodes and fills the msm_result.
// Inputs as prepared by k_decompress_valid / k_prep_scalars_vt (invalid point: zero first element; bad scalar: word 7 = ~0).
__global__ void __launch_bounds__(64) k_single_msm_vt(const uint32_t* __restrict__ kk, const uint32_t* __restrict__ k3,
                                                      const uint4* __restrict__ niels, size_t n, uint32_t G,
                                                      uint4* __restrict__ partials, int* __restrict__ flags) {
  const size_t thread = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t g = thread >> 2;
  const int sub = (int)(threadIdx.x & 3);
  const int gbase = (int)(threadIdx.x & 31 & ~3);
  const unsigned gmask = 0xfu << gbase;
  if (g >= G) return;          // the four lanes of a group leave together
  bool ok = true;
  for (size_t t = g; t < n; t += G) {
    uint4 first = __ldg(niels + ZKP_NIELS_U4 * t), second = __ldg(niels + ZKP_NIELS_U4 * t + 1);
    if ((first.x | first.y | first.z | first.w | second.x | second.y | second.z | second.w) == 0) {
      if (sub == 0) atomicMin(&flags[0], (int)t);
      ok = false;
    }
    if (__ldg(kk + 8 * t + 7) == 0xffffffffu) {
      if (sub == 0) atomicMin(&flags[1], (int)t);
      ok = false;
    }
  }
  ge_ext acc;
  ge_identity(acc);
  bool started = false;
  if (ok) {
    for (int i = 253; i >= 0; i--) {
      if (started) ge_double_coop4(acc, sub, gmask, gbase);
      const int b = i + 1, wi = b >> 5, sh = b & 31;
      for (size_t t = g; t < n; t += G) {
        uint32_t w3 = __ldg(k3 + 8 * t + wi), w1 = __ldg(kk + 8 * t + wi);
        if (wi == 7) w1 &= 0x7fffffffu;
        int d = (int)((w3 >> sh) & 1u) - (int)((w1 >> sh) & 1u);
        if (d != 0) {
          uint32_t neg = (__ldg(kk + 8 * t + 7) >> 31) ^ (d < 0 ? 1u : 0u);
          ge_aniels q;
          load_aniels(q, niels, (uint32_t)t);
          ge_aniels_cneg(q, neg);
          ge_madd_coop4(acc, q, sub, gmask, gbase);
          started = true;
        }
      }
    }
  }
  if (sub == 0) store_ext(partials + 8 * g, acc);
}

// one block of 128 threads: sum of the G partial sums, ristretto encode, status (the layout k_finish writes)
__global__ void __launch_bounds__(128) k_single_finish(const uint4* __restrict__ partials, uint32_t G,
                                                       const int* __restrict__ flags, msm_result* __restrict__ res,
                                                       uint4* __restrict__ partial_out) {
  __shared__ uint4 sm[4 * 8];
  ge_ext acc, p;
  ge_identity(acc);
  for (uint32_t i = threadIdx.x; i < G; i += blockDim.x) {
    load_ext(p, partials + (size_t)i * 8);
    ge_add(acc, acc, p);
  }
#pragma unroll 1
  for (int off = 16; off >= 1; off >>= 1) {
    shfl_down_ext(p, acc, off, 32);
    ge_add(acc, acc, p);
  }
  if ((threadIdx.x & 31) == 0) store_ext(sm + (threadIdx.x >> 5) * 8, acc);
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int k = 1; k < 4; k++) {
    load_ext(p, sm + k * 8);
    ge_add(acc, acc, p);
  }
  if (partial_out) store_ext(partial_out, acc);
  uint32_t enc[8];
  ristretto_encode(enc, acc);
  uint32_t z = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { res->enc[i] = enc[i]; z |= enc[i]; }
  const int bad_pt = flags[0], bad_sc = flags[1];
  int status = 0;
  long long first_bad = -1;
  if (bad_pt != 0x7fffffff) { status = 1; first_bad = bad_pt; }
  else if (bad_sc != 0x7fffffff) { status = 3; first_bad = bad_sc; }
  res->status = status;
  res->is_identity = (z == 0) ? 1 : 0;
  res->first_bad = first_bad;
}

// ---- constant-time path ----------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_decompress_ext(const uint4* __restrict__ enc, size_t n, uint4* __restrict__ ext,
                                                        int* __restrict__ flags) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t w[8];
  load_words8(w, enc + 2 * i);
  ge_ext p;
  uint32_t ok = ristretto_decode(p.X, p.Y, p.T, w);
  fe_one(p.Z);
  if (!ok) {
    ge_identity(p);
    atomicMin(&flags[0], (int)i);
  }
  store_ext(ext + 8 * i, p);
}

__global__ void __launch_bounds__(256) k_limbs_to_ext(const unsigned long long* __restrict__ limbs, size_t n,
                                                      uint4* __restrict__ ext) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge_ext p;
  load_ext_limbs51(p, limbs + 20 * i);
  store_ext(ext + 8 * i, p);
}

__device__ __forceinline__ void store_pniels(uint4* o, const ge_pniels& q) {
  store_fe(o, q.YplusX); store_fe(o + 2, q.YminusX); store_fe(o + 4, q.Z); store_fe(o + 6, q.T2d);
}

// per term: table of 1P..8P as projective Niels (8 x 128 B) and the biased scalar k + 0x88..8.
// Table layouts (uint4 units; entry e = 1..8, word q = 0..7):
//   IL = false  tables[t * 64 + 8 (e-1) + q]                                  one contiguous KB per term (any CSR shape)
//   IL = true   tables[((g * T + r) * 64 + 8 (e-1) + q) * 32 + l]             term t = j * T + r of proof j = 32 g + l:
//               the 32 proofs of a group are interleaved 16 bytes at a time, so the warp of k_small_msm_ct that works on
//               the same constraint of 32 consecutive proofs reads every table word with ONE coalesced 512-byte access
//               (batch proving: all proofs share the statement, T terms each)
template <bool IL>
__device__ __forceinline__ size_t ct_table_index(size_t t, uint32_t T) {
  if (!IL) return t * 64;
  const size_t j = t / T, r = t % T;
  return ((j >> 5) * T + r) * 64 * 32 + (j & 31);
}
// shared_of (batch proving, optional): shared_of[r] >= 0 says that term r of every proof uses the batch-static point with
// that index, whose table is built ONCE (k_build_tables<false> over the static points) and read by all proofs; such
// terms only get their biased scalar here.  (in CMZ 21 of the 31 prover terms use batch-static bases.)
template <bool IL>
__global__ void __launch_bounds__(128) k_build_tables(const uint4* __restrict__ ext, const uint4* __restrict__ scalars,
                                                      size_t n, uint32_t T, uint4* __restrict__ tables,
                                                      uint4* __restrict__ biased, int* __restrict__ flags,
                                                      const int32_t* __restrict__ shared_of = nullptr) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  {
    uint32_t s[8], l[8], d[8], bias[8], r[8];
    load_words8(s, scalars + 2 * i);
    sc_load_l(l);
    uint32_t canonical = sub8(d, s, l);
    if (!canonical) atomicMin(&flags[1], (int)i);
#pragma unroll
    for (int k = 0; k < 8; k++) bias[k] = 0x88888888u;
    add8(r, s, bias);
    biased[2 * i] = make_uint4(r[0], r[1], r[2], r[3]);
    biased[2 * i + 1] = make_uint4(r[4], r[5], r[6], r[7]);
  }
  if (shared_of && shared_of[i % T] >= 0) return;   // public: depends on the statement only
  const size_t stride = IL ? 32 : 1;
  ge_ext p, m;
  load_ext(p, ext + 8 * i);
  ge_pniels pn, cur;
  ge_to_pniels(pn, p);
  uint4* tab = tables + ct_table_index<IL>(i, T);
  auto put = [&](int e, const ge_pniels& q) {   // entry e (0-based)
    uint4* o = tab + (size_t)(8 * e) * stride;
    const fe* f[4] = {&q.YplusX, &q.YminusX, &q.Z, &q.T2d};
#pragma unroll
    for (int c = 0; c < 4; c++) {
      o[(size_t)(2 * c) * stride] = make_uint4(f[c]->v[0], f[c]->v[1], f[c]->v[2], f[c]->v[3]);
      o[(size_t)(2 * c + 1) * stride] = make_uint4(f[c]->v[4], f[c]->v[5], f[c]->v[6], f[c]->v[7]);
    }
  };
  put(0, pn);
  m = p;
#pragma unroll 1
  for (int k = 1; k < 8; k++) {
    ge_add_pniels(m, m, pn);
    ge_to_pniels(cur, m);
    put(k, cur);
  }
}

// COOP: four adjacent lanes share one MSM (kernels.cuh ge_double_coop4 / ge_add_pniels_coop4; every lane scans the whole
// table, so the lookup stays constant time and the lanes stay in lock step): the latency schedule for small batches.
template <bool IL, bool COOP = false>
__global__ void __launch_bounds__(64) k_small_msm_ct(const uint32_t* __restrict__ biased, const uint4* __restrict__ tables,
                                                     const unsigned long long* __restrict__ offsets,
                                                     const uint32_t* __restrict__ order, size_t M, uint32_t T,
                                                     uint4* __restrict__ out,
                                                     const int32_t* __restrict__ shared_of = nullptr,
                                                     const uint4* __restrict__ shared_tables = nullptr) {
  const size_t thread = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t tid = COOP ? thread >> 2 : thread;
  const int sub = COOP ? (int)(threadIdx.x & 3) : 0;
  const int gbase = (int)(threadIdx.x & 31 & ~3);
  const unsigned gmask = 0xfu << gbase;
  if (tid >= M) return;
  const size_t j = order[tid];   // public: MSMs sorted by (public) size so the lanes of a warp do equal work
  const size_t lo = offsets[j], hi = offsets[j + 1];   // public
  // table of the first term; consecutive terms of one MSM belong to one proof, so their tables are `tstep` apart
  const size_t tab0 = ct_table_index<IL>(lo, T), tstep = IL ? (size_t)64 * 32 : 64;
  const uint32_t r0 = shared_of ? (uint32_t)(lo % T) : 0u;   // index of the first term within its proof
  ge_ext acc;
  ge_identity(acc);
#pragma unroll 1
  for (int w = 63; w >= 0; w--) {
    if (COOP) {
#pragma unroll 1
      for (int dd = 0; dd < 4; dd++) ge_double_coop4(acc, sub, gmask, gbase);
    } else {
      ge_double(acc, acc);
      ge_double(acc, acc);
      ge_double(acc, acc);
      ge_double(acc, acc);
    }
#pragma unroll 1
    for (size_t t = lo; t < hi; t++) {
      uint32_t word = __ldg(biased + 8 * t + (w >> 3));
      int d = (int)((word >> ((w & 7) * 4)) & 15u) - 8;   // signed digit in [-8, 7]
      uint32_t sign = (uint32_t)d >> 31;
      uint32_t mag = (uint32_t)((d ^ -(int)sign) + (int)sign);
      // constant-time lookup: start from the identity, scan all eight entries
      ge_pniels sel;
      fe_one(sel.YplusX); fe_one(sel.YminusX); fe_one(sel.Z); fe_zero(sel.T2d);
      size_t stride = IL ? 32 : 1;
      const uint4* tab = tables + tab0 + (t - lo) * tstep;   // public: depends on the term index only
      if (shared_of) {   // batch-static base: one table for all proofs (the whole warp reads the same addresses)
        const int32_t sh = shared_of[r0 + (uint32_t)(t - lo)];
        if (sh >= 0) {
          tab = shared_tables + (size_t)sh * 64;
          stride = 1;
        }
      }
#pragma unroll 1
      for (uint32_t e = 1; e <= 8; e++) {
        uint32_t mask = 0u - (uint32_t)(mag == e);
        const uint4* ent = tab + (size_t)(8 * (e - 1)) * stride;
        uint32_t* dst = (uint32_t*)&sel;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          uint4 v = __ldg(ent + (size_t)q * stride);
          dst[4 * q + 0] ^= mask & (dst[4 * q + 0] ^ v.x);
          dst[4 * q + 1] ^= mask & (dst[4 * q + 1] ^ v.y);
          dst[4 * q + 2] ^= mask & (dst[4 * q + 2] ^ v.z);
          dst[4 * q + 3] ^= mask & (dst[4 * q + 3] ^ v.w);
        }
      }
      ge_pniels_cneg(sel, sign);
      if (COOP) ge_add_pniels_coop4(acc, sel, sub, gmask, gbase);
      else ge_add_pniels(acc, acc, sel);
    }
  }
  if (sub != 0) return;
  uint32_t enc[8];
  ristretto_encode(enc, acc);
  out[2 * j] = make_uint4(enc[0], enc[1], enc[2], enc[3]);
  out[2 * j + 1] = make_uint4(enc[4], enc[5], enc[6], enc[7]);
}

// ---- comb path of batch proving (comb.cuh; SURVEY 8f row f4, second half) ------------------------------------------
// Every base a statement's constraints use gets ONE signed four-tooth comb -- per proof for instance points, per batch
// for the batch-static points -- and every constraint MSM then runs 64 doublings instead of 256.
// Layouts (uint4 units): per-proof combs, projective Niels entries e = 0..7 of 8 uint4, interleaved like the Straus tables,
// combs[ct_table_index<true>(j * U + u, U) + (8 e + q) * 32] for slot u of proof j; shared combs, affine Niels entries of
// 6 uint4 (normalised once per batch: mixed additions, shorter scans), combs[u * 48 + 6 e + q].

// one thread per (proof j, slot u), i = j * U + u: the comb of the proof's copy of point slot_point[u]
// (IL = false: the shared combs, built from proof 0's copy: n = U threads)
template <bool IL>
__global__ void __launch_bounds__(64) k_build_combs(const unsigned long long* __restrict__ limbs, size_t n, uint32_t U,
                                                    uint32_t points_per_proof, const int32_t* __restrict__ slot_point,
                                                    uint4* __restrict__ combs) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t j = i / U;
  const uint32_t u = (uint32_t)(i % U);
  ge_ext P;
  load_ext_limbs51(P, limbs + (j * points_per_proof + (size_t)slot_point[u]) * 20);
  ge_pniels E[8];
  comb_build(E, P);
  if (IL) {
    uint4* tab = combs + ct_table_index<true>(i, U);
#pragma unroll 1
    for (int e = 0; e < 8; e++) {
      const uint32_t* w = (const uint32_t*)&E[e];
#pragma unroll
      for (int q = 0; q < 8; q++)
        tab[(size_t)(8 * e + q) * 32] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    }
  } else {
    uint4* tab = combs + i * 48;
#pragma unroll 1
    for (int e = 0; e < 8; e++) {
      ge_aniels a;
      comb_entry_to_affine(a, E[e]);
      const uint32_t* w = (const uint32_t*)&a;
#pragma unroll
      for (int q = 0; q < 6; q++) tab[6 * e + q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    }
  }
}

// one thread per term: the sign bits of the signed-digit form of its (canonical) scalar
__global__ void __launch_bounds__(256) k_comb_recode(const uint4* __restrict__ scalars, size_t n, uint4* __restrict__ recoded) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[8], m[8];
  load_words8(s, scalars + 2 * i);
  comb_recode(m, s);
  recoded[2 * i] = make_uint4(m[0], m[1], m[2], m[3]);
  recoded[2 * i + 1] = make_uint4(m[4], m[5], m[6], m[7]);
}

// one thread per MSM (a constraint of a proof; CSR over the batch's term list, T terms per proof).  term_slot[r] for
// term r of a proof: u >= 0 = per-proof comb slot u, -(s + 1) = shared comb s.  Public: order, offsets, term_slot.
__global__ void __launch_bounds__(64) k_small_msm_comb(const uint32_t* __restrict__ recoded, const uint4* __restrict__ combs,
                                                       const uint4* __restrict__ shared_combs,
                                                       const int32_t* __restrict__ term_slot,
                                                       const unsigned long long* __restrict__ offsets,
                                                       const uint32_t* __restrict__ order, size_t M, uint32_t T, uint32_t U,
                                                       uint4* __restrict__ out) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= M) return;
  const size_t j = order[tid];
  const size_t lo = offsets[j], hi = offsets[j + 1];
  const size_t proof = lo / T;
  const uint32_t r0 = (uint32_t)(lo % T);
  ge_ext acc;
  ge_identity(acc);
#pragma unroll 1
  for (int col = 63; col >= 0; col--) {
    ge_double(acc, acc);
#pragma unroll 1
    for (size_t t = lo; t < hi; t++) {
      const int32_t slot = term_slot[r0 + (uint32_t)(t - lo)];
      const uint32_t* mw = recoded + 8 * t + (col >> 5);
      uint32_t idx, neg;
      comb_column_words(idx, neg, __ldg(mw), __ldg(mw + 2), __ldg(mw + 4), __ldg(mw + 6), col & 31);
      if (slot >= 0) {   // public: depends on the statement only (and is the same for every lane of the warp)
        const uint4* tab = combs + (((proof >> 5) * U + (uint32_t)slot) * 64 * 32 + (proof & 31));   // ct_table_index<true>(proof * U + slot, U)
        ge_pniels sel;
        comb_select(sel, [&](uint32_t e, int q, uint32_t& x, uint32_t& y, uint32_t& z, uint32_t& ww) {
          const uint4 v = __ldg(tab + (size_t)(8 * e + q) * 32);
          x = v.x; y = v.y; z = v.z; ww = v.w;
        }, idx, neg);
        ge_add_pniels(acc, acc, sel);
      } else {
        const uint4* tab = shared_combs + (size_t)(-slot - 1) * 48;
        ge_aniels sel;
        comb_select_affine(sel, [&](uint32_t e, int q, uint32_t& x, uint32_t& y, uint32_t& z, uint32_t& ww) {
          const uint4 v = __ldg(tab + 6 * e + q);
          x = v.x; y = v.y; z = v.z; ww = v.w;
        }, idx, neg);
        ge_madd(acc, acc, sel);
      }
    }
  }
  uint32_t enc[8];
  ristretto_encode(enc, acc);
  out[2 * j] = make_uint4(enc[0], enc[1], enc[2], enc[3]);
  out[2 * j + 1] = make_uint4(enc[4], enc[5], enc[6], enc[7]);
}


// ---- comb path staged in shared memory: one CTA per group of 32 proofs --------------------------------------------------
// The constant-time scans are the memory side of the prover: eight entries per term and column, ~117 GB per 2^16 CMZ
// proofs when they come from global memory (k_small_msm_ct / k_small_msm_comb).  Here a CTA owns 32 consecutive proofs:
// it stages their per-proof combs (U x 32 KB, contiguous in the group-interleaved layout of k_build_combs<true>) and the
// batch's shared combs (Us x 768 B) in shared memory with 1-D bulk copies (cp.async.bulk, completion on an mbarrier) and
// scans from there: lane l of every warp works for proof 32 g + l, so a per-proof entry word is one conflict-free
// 512-byte LDS.128 and a shared entry word is one broadcast.  Warps take UNITS -- a constraint cut into pieces of at
// most `piece` terms by pv_make_units -- round robin, so the warps of a CTA do equal work whatever the constraint sizes
// are (CMZ: ten 2-term constraints and one 11-term constraint = 16 units of <= 2 terms); the pieces of a constraint are
// added and encoded at the end.  Scalars arrive recoded and group-interleaved (k_comb_recode_il).  Public: everything
// but `recoded`; the digits only ever feed arithmetic masks.
//   shared memory (uint4 units): [0, Us*48) shared combs | [.., + U*2048) per-proof combs | [.., + n_units*256) unit sums
//   | one mbarrier
__global__ void __launch_bounds__(256) k_comb_recode_il(const uint4* __restrict__ scalars, size_t N, uint32_t T,
                                                        uint32_t* __restrict__ recoded_il) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // term r of proof j: i = j * T + r
  if (i >= N * (size_t)T) return;
  const size_t j = i / T, r = i % T;
  uint32_t s[8], m[8];
  load_words8(s, scalars + 2 * i);
  comb_recode(m, s);
  uint32_t* o = recoded_il + (((j >> 5) * T + r) * 8) * 32 + (j & 31);
#pragma unroll
  for (int w = 0; w < 8; w++) o[w * 32] = m[w];
}

#ifdef ZKP_HOST_EMUL
static uint4 emul_dynamic_smem[14528];   // 227 KB: blocks run one at a time under emul_launch_mt
#endif

static inline size_t comb_cta_smem_bytes(size_t U, size_t Us, size_t n_units) {
  return 16 * (Us * 48 + U * 2048 + n_units * 256) + 16;
}

__global__ void __launch_bounds__(512, 1)
    k_comb_msm_cta(const uint32_t* __restrict__ recoded_il, const uint4* __restrict__ combs,
                   const uint4* __restrict__ shared_combs, const int32_t* __restrict__ term_slot,
                   const int32_t* __restrict__ unit_term0, const int32_t* __restrict__ unit_nterms,
                   const int32_t* __restrict__ cons_unit0, size_t N, uint32_t T, uint32_t U, uint32_t Us,
                   uint32_t n_units, uint32_t k, uint4* __restrict__ out) {
#ifdef ZKP_HOST_EMUL
  uint4* smem = emul_dynamic_smem;
#else
  extern __shared__ uint4 comb_cta_smem[];
  uint4* smem = comb_cta_smem;
#endif
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t g = blockIdx.x, proof = g * 32 + lane;
  uint4* s_shared = smem;
  uint4* s_combs = smem + (size_t)Us * 48;
  uint4* s_res = s_combs + (size_t)U * 2048;
  const uint4* g_combs = combs + g * (size_t)U * 2048;
#if ZKP_DEVICE_ASM
  {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(s_res + (size_t)n_units * 256);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t bytes = (Us * 48 + U * 2048) * 16;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      if (Us)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(s_shared)),
                     "l"(shared_combs), "r"(Us * 768u), "r"(bar)
                     : "memory");
      for (uint32_t u = 0; u < U; u++)   // 32 KB per slot
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(s_combs + (size_t)u * 2048)),
                     "l"(g_combs + (size_t)u * 2048), "r"(32768u), "r"(bar)
                     : "memory");
    }
    uint32_t done;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(bar), "r"(0u)
                   : "memory");
    } while (!done);
  }
#else
  if (threadIdx.x == 0) {
    for (size_t i = 0; i < (size_t)Us * 48; i++) s_shared[i] = shared_combs[i];
    for (size_t i = 0; i < (size_t)U * 2048; i++) s_combs[i] = g_combs[i];
  }
  __syncthreads();
#endif
  // ---- the units of this warp -------------------------------------------------------------------------------------
#pragma unroll 1
  for (uint32_t u = warp; u < n_units; u += nw) {
    const uint32_t r0 = (uint32_t)unit_term0[u], nt = (uint32_t)unit_nterms[u];
    // Shared combs (the same table for all 32 lanes = proofs, each lane wanting its own entry): with at most two terms
    // per unit the table lives in REGISTERS, spread over the warp -- word q of entry e in slot q / 4 of lane
    // 8 (q % 4) + e -- and a lane fetches its entry with 24 shuffles whose source lane is 8 (q % 4) + idx: no memory is
    // indexed by a digit, and the 48 shared-memory loads + 192 masked selects of the scan go away.
    const bool in_regs = nt <= 2;
    uint32_t stab0[6] = {0, 0, 0, 0, 0, 0}, stab1[6] = {0, 0, 0, 0, 0, 0};
    if (in_regs) {
#pragma unroll 1
      for (uint32_t t = 0; t < nt; t++) {
        const int32_t slot = term_slot[r0 + t];
        if (slot < 0) {
          const uint32_t* tab32 = (const uint32_t*)(s_shared + (size_t)(-slot - 1) * 48) + 24 * (lane & 7) + (lane >> 3);
#pragma unroll
          for (int r = 0; r < 6; r++) {
            const uint32_t w = tab32[4 * r];
            if (t == 0) stab0[r] = w;
            else stab1[r] = w;
          }
        }
      }
    }
    ge_ext acc;
    ge_identity(acc);
#pragma unroll 1
    for (int col = 63; col >= 0; col--) {
      ge_double(acc, acc);
#pragma unroll 1
      for (uint32_t t = 0; t < nt; t++) {
        const uint32_t r = r0 + t;
        const int32_t slot = term_slot[r];
        const uint32_t* mw = recoded_il + ((g * T + r) * 8 + (uint32_t)(col >> 5)) * 32 + lane;
        uint32_t idx, neg;
        comb_column_words(idx, neg, __ldg(mw), __ldg(mw + 64), __ldg(mw + 128), __ldg(mw + 192), col & 31);
        if (slot >= 0) {   // public: the statement's, the same for every lane
          const uint4* tab = s_combs + (size_t)slot * 2048 + lane;
          ge_pniels sel;
          comb_select(sel, [&](uint32_t e, int q, uint32_t& x, uint32_t& y, uint32_t& z, uint32_t& ww) {
            const uint4 v = tab[(8 * e + q) * 32];
            x = v.x; y = v.y; z = v.z; ww = v.w;
          }, idx, neg);
          ge_add_pniels(acc, acc, sel);
        } else if (in_regs) {
          ge_aniels sel;
          uint32_t* d = (uint32_t*)&sel;
#pragma unroll
          for (int q = 0; q < 24; q++) {
            const uint32_t mine = t == 0 ? stab0[q >> 2] : stab1[q >> 2];
            d[q] = __shfl_sync(0xffffffffu, mine, (int)(((q & 3) << 3) + idx));
          }
          ge_aniels_cneg(sel, neg);
          ge_madd(acc, acc, sel);
        } else {
          const uint4* tab = s_shared + (size_t)(-slot - 1) * 48;
          ge_aniels sel;
          comb_select_affine(sel, [&](uint32_t e, int q, uint32_t& x, uint32_t& y, uint32_t& z, uint32_t& ww) {
            const uint4 v = tab[6 * e + q];
            x = v.x; y = v.y; z = v.z; ww = v.w;
          }, idx, neg);
          ge_madd(acc, acc, sel);
        }
      }
    }
    uint4* o = s_res + (size_t)u * 256 + lane;
    const uint32_t* w = (const uint32_t*)&acc;
#pragma unroll
    for (int q = 0; q < 8; q++) o[q * 32] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  }
  __syncthreads();
  // ---- constraints: add the pieces, encode ------------------------------------------------------------------------
#pragma unroll 1
  for (uint32_t c = warp; c < k; c += nw) {
    const uint32_t u0 = (uint32_t)cons_unit0[c], u1 = (uint32_t)cons_unit0[c + 1];
    ge_ext acc, p;
    auto load_unit = [&](ge_ext& dst, uint32_t u) {
      const uint4* o = s_res + (size_t)u * 256 + lane;
      uint32_t* w = (uint32_t*)&dst;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const uint4 v = o[q * 32];
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
    };
    load_unit(acc, u0);
#pragma unroll 1
    for (uint32_t u = u0 + 1; u < u1; u++) {
      load_unit(p, u);
      ge_add(acc, acc, p);
    }
    uint32_t enc[8];
    ristretto_encode(enc, acc);
    if (proof < N) {
      out[2 * (proof * k + c)] = make_uint4(enc[0], enc[1], enc[2], enc[3]);
      out[2 * (proof * k + c) + 1] = make_uint4(enc[4], enc[5], enc[6], enc[7]);
    }
  }
}

}  // namespace zkp
