// The batch prover's MSM kernels compiled for the HOST (cuda_shim.h + -DZKP_HOST_EMUL) and run thread by thread:
// k_pv_gather -> (k_build_tables + k_small_msm_ct | k_build_combs + k_comb_recode + k_small_msm_comb) with the plan of
// pv_plan.hpp, exactly the launch sequence of zkp_prove_batch (api.cu).  Checks the CSR / schedule / interleaved-table
// indexing and the kernels' logic against the oracle without a GPU.  TEST INFRASTRUCTURE ONLY.
#include "cuda_shim.h"
#include <vector>
#include "../../zkp_b200/csrc/pv_kernels.cuh"
#include "../../zkp_b200/csrc/pv_plan.hpp"
using namespace zkp;

// step (3) of zkp_prove_batch: the N * k constant-time MSMs + compress; returns the not-uniform flag
static int run_msms(const pv_desc& d, const pv_plan& plan, int nc, size_t N, const uint64_t* limbs, const uint8_t* blind,
                    int share, int comb, uint8_t* com) {
  const int p = d.p, k = d.k, T = d.n_terms, ni = d.ni;
  const size_t total = N * (size_t)T, M = N * (size_t)k, Npad = (N + 31) / 32 * 32;
  std::vector<uint4> scalars_flat(2 * total + 2), ext_flat(8 * total + 8), out(2 * M + 2);
  std::vector<unsigned long long> offsets(M + 1);
  std::vector<uint32_t> order(M + 1);
  int flags[4] = {0x7fffffff, 0x7fffffff, 0, 0};
  const size_t gthreads = std::max(total, M) ? std::max(total, M) : 1;
  emul_launch((unsigned)((gthreads + 255) / 256), 256, k_pv_gather, d, N, (const unsigned long long*)limbs, blind,
              scalars_flat.data(), ext_flat.data(), offsets.data(), order.data(), share ? &flags[2] : (int*)nullptr);
  std::vector<uint4> biased(2 * total + 2);
  if (comb) {
    const size_t U = plan.comb_slot_point.size(), Us = plan.comb_shared_point.size();
    std::vector<uint4> combs(Npad * U * 64 + 64), shared(Us * 64 + 64);
    if (U)
      emul_launch((unsigned)((N * U + 63) / 64), 64, k_build_combs<true>, (const unsigned long long*)limbs, N * U, (uint32_t)U,
                  (uint32_t)p, plan.comb_slot_point.data(), combs.data());
    if (Us)
      emul_launch((unsigned)((Us + 63) / 64), 64, k_build_combs<false>, (const unsigned long long*)limbs, Us, (uint32_t)Us,
                  (uint32_t)p, plan.comb_shared_point.data(), shared.data());
    emul_launch((unsigned)((total + 255) / 256), 256, k_comb_recode, (const uint4*)scalars_flat.data(), total, biased.data());
    emul_launch((unsigned)((M + 63) / 64), 64, k_small_msm_comb, (const uint32_t*)biased.data(), (const uint4*)combs.data(),
                (const uint4*)shared.data(), plan.term_slot.data(), (const unsigned long long*)offsets.data(),
                (const uint32_t*)order.data(), M, (uint32_t)T, (uint32_t)U, out.data());
  } else {
    std::vector<uint4> tables(Npad * (size_t)T * 64 + 64), s_ext(8 * (size_t)nc + 8), s_tab(64 * (size_t)nc + 64),
        s_zero(2 * (size_t)nc + 2, make_uint4(0, 0, 0, 0)), s_bias(2 * (size_t)nc + 2);
    const int32_t* shared_of = share ? plan.term_shared.data() : nullptr;
    if (share) {
      emul_launch((unsigned)((nc + 255) / 256), 256, k_limbs_to_ext, (const unsigned long long*)limbs + (size_t)ni * 20,
                  (size_t)nc, s_ext.data());
      emul_launch((unsigned)((nc + 127) / 128), 128, k_build_tables<false>, (const uint4*)s_ext.data(), (const uint4*)s_zero.data(),
                  (size_t)nc, 1u, s_tab.data(), s_bias.data(), &flags[2], (const int32_t*)nullptr);
    }
    emul_launch((unsigned)((total + 127) / 128), 128, k_build_tables<true>, (const uint4*)ext_flat.data(),
                (const uint4*)scalars_flat.data(), total, (uint32_t)T, tables.data(), biased.data(), &flags[2], shared_of);
    emul_launch((unsigned)((M + 63) / 64), 64, k_small_msm_ct<true, false>, (const uint32_t*)biased.data(),
                (const uint4*)tables.data(), (const unsigned long long*)offsets.data(), (const uint32_t*)order.data(), M,
                (uint32_t)T, out.data(), shared_of, share ? (const uint4*)s_tab.data() : (const uint4*)nullptr);
  }
  memcpy(com, out.data(), M * 32);
  return flags[2];
}

static void fill_desc(pv_desc* d, const pv_plan& plan, int m, int ni, int nc, int k, const int32_t* lhs,
                      const int32_t* cons_off, const int32_t* term_scalar, const int32_t* term_point) {
  memset(d, 0, sizeof *d);
  d->m = m; d->p = ni + nc; d->k = k; d->n_terms = k ? cons_off[k] : 0; d->ni = ni;
  d->term_shared = plan.term_shared.data();
  d->lhs = lhs; d->cons_off = cons_off; d->term_scalar = term_scalar; d->term_point = term_point;
  d->cons_slot = plan.cons_slot.data();
}

extern "C" {
// limbs [N][p][20] (FieldElement51 X,Y,Z,T), blind [N][m][32] canonical -> com [N][k][32]; returns the not-uniform flag
// (1 if a common point differed between proofs while share was requested: the caller reruns without sharing, as api.cu does)
int emul_prove_msms(int m, int ni, int nc, int k, const int32_t* lhs, const int32_t* cons_off, const int32_t* term_scalar,
                    const int32_t* term_point, size_t N, const uint64_t* limbs, const uint8_t* blind, int share, int comb,
                    uint8_t* com) {
  pv_plan plan;
  pv_make_plan(ni, ni + nc, k, cons_off, term_point, share != 0, comb != 0, &plan);
  pv_desc d;
  fill_desc(&d, plan, m, ni, nc, k, lhs, cons_off, term_scalar, term_point);
  return run_msms(d, plan, nc, N, limbs, blind, share, comb, com);
}

// All of zkp_prove_batch (api.cu) for N proofs: k_compress_limbs -> k_pv_blind -> MSMs -> k_pv_finish.
// labels: p NUL-terminated point labels (instance ++ common); prefix53: the transcript state every proof starts from.
// Returns 0, 3 for a non-canonical secret (ZKP_ERR_SCALAR), or -100 - flag when sharing was requested for points that differ.
int emul_prove_batch(int m, int ni, int nc, int k, const char* labels, const int32_t* lhs, const int32_t* cons_off,
                     const int32_t* term_scalar, const int32_t* term_point, const uint32_t* prefix53, size_t N,
                     const uint8_t* secrets, const uint64_t* limbs, const uint8_t* entropy, int share, int comb,
                     uint8_t* enc_out, uint8_t* com_out, uint8_t* resp_out) {
  const int p = ni + nc;
  pv_plan plan;
  pv_make_plan(ni, p, k, cons_off, term_point, share != 0, comb != 0, &plan);
  pv_desc d;
  fill_desc(&d, plan, m, ni, nc, k, lhs, cons_off, term_scalar, term_point);
  std::vector<uint32_t> loff, llen;
  std::vector<uint8_t> pool;
  const char* lp = labels;
  for (int i = 0; i < p; i++) {
    const size_t len = strlen(lp);
    loff.push_back((uint32_t)pool.size());
    llen.push_back((uint32_t)len);
    pool.insert(pool.end(), lp, lp + len);
    lp += len + 1;
  }
  pool.push_back(0);
  d.label_off = loff.data(); d.label_len = llen.data(); d.labels = pool.data();
  int flags[4] = {0x7fffffff, 0x7fffffff, 0, 0};
  std::vector<uint4> enc(2 * N * p + 2);
  if (p) emul_launch((unsigned)((N * p + 255) / 256), 256, k_compress_limbs, (const unsigned long long*)limbs, N * (size_t)p, enc.data());
  std::vector<uint32_t> states(N * 56 + 56);
  std::vector<uint8_t> blind(N * (size_t)m * 32 + 32), com(N * (size_t)k * 32 + 32), resp(N * (size_t)m * 32 + 32);
  emul_launch((unsigned)((N + 127) / 128), 128, k_pv_blind, d, prefix53, N, (const uint8_t*)enc.data(), secrets, entropy,
              states.data(), blind.data(), flags);
  if (k) {
    const int nu = run_msms(d, plan, nc, N, limbs, blind.data(), share, comb, com.data());
    if (nu) return -100 - nu;
  }
  emul_launch((unsigned)((N + 127) / 128), 128, k_pv_finish, d, N, (const uint32_t*)states.data(), (const uint8_t*)com.data(),
              secrets, (const uint8_t*)blind.data(), resp.data());
  memcpy(enc_out, enc.data(), N * (size_t)p * 32);
  memcpy(com_out, com.data(), N * (size_t)k * 32);
  memcpy(resp_out, resp.data(), N * (size_t)m * 32);
  return flags[1] != 0x7fffffff ? 3 : 0;
}

// zkp_msm_vartime_batched (api.cu): k_decompress_valid + k_prep_scalars_vt + k_small_msm_vt<false> over a CSR batch of
// M small MSMs given as encodings; out [M][32], status [M] (0 ok, 1 undecodable point, 3 non-canonical scalar)
void emul_msm_vartime_batched(const uint8_t* scalars, const uint8_t* points, const uint64_t* offsets, size_t M, uint8_t* out,
                              int32_t* status) {
  const size_t total = (size_t)offsets[M];
  std::vector<uint4> sc(2 * total + 2), pt(2 * total + 2), niels(6 * total + 6), kk(2 * total + 2), k3(2 * total + 2), o(2 * M + 2);
  if (total) { memcpy(sc.data(), scalars, total * 32); memcpy(pt.data(), points, total * 32); }
  std::vector<unsigned long long> off(offsets, offsets + M + 1);
  std::vector<uint32_t> order(M + 1);
  for (size_t j = 0; j < M; j++) order[j] = (uint32_t)(M - 1 - j);   // any permutation is a valid schedule
  std::vector<int> st(M + 1, -1);
  if (total) {
    emul_launch((unsigned)((total + 255) / 256), 256, k_decompress_valid, (const uint4*)pt.data(), total, niels.data());
    emul_launch((unsigned)((total + 255) / 256), 256, k_prep_scalars_vt, (const uint4*)sc.data(), total, kk.data(), k3.data());
  }
  emul_launch((unsigned)((M + 63) / 64), 64, k_small_msm_vt<false>, (const uint32_t*)kk.data(), (const uint32_t*)k3.data(),
              (const uint4*)niels.data(), (const unsigned long long*)off.data(), (const uint32_t*)order.data(), M, o.data(),
              st.data());
  memcpy(out, o.data(), M * 32);
  for (size_t j = 0; j < M; j++) status[j] = st[j];
}

// zkp_msm_ct_batched (api.cu) with compressed points: k_decompress_ext + k_build_tables<false> + k_small_msm_ct<false>;
// returns 0, 1 (undecodable point) or 3 (non-canonical scalar) like the entry point
int emul_msm_ct_batched(const uint8_t* scalars, const uint8_t* points, const uint64_t* offsets, size_t M, uint8_t* out) {
  const size_t total = (size_t)offsets[M];
  std::vector<uint4> sc(2 * total + 2), pt(2 * total + 2), ext(8 * total + 8), tables(64 * total + 64), biased(2 * total + 2),
      o(2 * M + 2);
  if (total) { memcpy(sc.data(), scalars, total * 32); memcpy(pt.data(), points, total * 32); }
  std::vector<unsigned long long> off(offsets, offsets + M + 1);
  std::vector<uint32_t> order(M + 1);
  for (size_t j = 0; j < M; j++) order[j] = (uint32_t)j;
  int flags[4] = {0x7fffffff, 0x7fffffff, 0, 0};
  if (total) {
    emul_launch((unsigned)((total + 255) / 256), 256, k_decompress_ext, (const uint4*)pt.data(), total, ext.data(), flags);
    emul_launch((unsigned)((total + 127) / 128), 128, k_build_tables<false>, (const uint4*)ext.data(), (const uint4*)sc.data(),
                total, 1u, tables.data(), biased.data(), flags, (const int32_t*)nullptr);
  }
  emul_launch((unsigned)((M + 63) / 64), 64, k_small_msm_ct<false, false>, (const uint32_t*)biased.data(),
              (const uint4*)tables.data(), (const unsigned long long*)off.data(), (const uint32_t*)order.data(), M, 1u, o.data(),
              (const int32_t*)nullptr, (const uint4*)nullptr);
  memcpy(out, o.data(), M * 32);
  if (flags[0] != 0x7fffffff) return 1;
  if (flags[1] != 0x7fffffff) return 3;
  return 0;
}
}
