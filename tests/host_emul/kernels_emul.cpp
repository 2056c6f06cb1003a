// The batch prover's MSM kernels compiled for the HOST (cuda_shim.h + -DZKP_HOST_EMUL) and run thread by thread:
// k_pv_gather -> (k_build_tables + k_small_msm_ct | k_build_combs + k_comb_recode + k_small_msm_comb) with the plan of
// pv_plan.hpp, exactly the launch sequence of zkp_prove_batch (api.cu).  Checks the CSR / schedule / interleaved-table
// indexing and the kernels' logic against the oracle without a GPU.  TEST INFRASTRUCTURE ONLY.
#include "cuda_shim.h"
#include <vector>
#include "../../zkp_b200/csrc/pv_kernels.cuh"
#include "../../zkp_b200/csrc/pv_plan.hpp"
#include "../../zkp_b200/csrc/bv_plan.hpp"
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
    std::vector<uint4> combs(Npad * U * 64 + 64), shared(Us * 48 + 16);   // the sizes of api.cu: 1 KB / 768 B per comb
    if (U)
      emul_launch((unsigned)((N * U + 63) / 64), 64, k_build_combs<true>, (const unsigned long long*)limbs, N * U, (uint32_t)U,
                  (uint32_t)p, plan.comb_slot_point.data(), combs.data());
    if (Us)
      emul_launch((unsigned)((Us + 63) / 64), 64, k_build_combs<false>, (const unsigned long long*)limbs, Us, (uint32_t)Us,
                  (uint32_t)p, plan.comb_shared_point.data(), shared.data());
    if (comb >= 2) {   // CTA-staged: group-interleaved recoding, one block per 32 proofs, units from pv_make_units
      std::vector<uint32_t> rec(Npad * (size_t)T * 8 + 8);
      emul_launch((unsigned)((total + 255) / 256), 256, k_comb_recode_il, (const uint4*)scalars_flat.data(), N, (uint32_t)T,
                  rec.data());
      const unsigned n_units = (unsigned)plan.unit_term0.size(), nw = n_units < 16 ? n_units : 16;
      if (comb_cta_smem_bytes(U, Us, n_units) > sizeof(emul_dynamic_smem)) return -1;
      emul_launch_mt((unsigned)((N + 31) / 32), 32 * nw, k_comb_msm_cta, (const uint32_t*)rec.data(),
                     (const uint4*)combs.data(), (const uint4*)shared.data(), plan.term_slot.data(), plan.unit_term0.data(),
                     plan.unit_nterms.data(), plan.cons_unit0.data(), N, (uint32_t)T, (uint32_t)U, (uint32_t)Us, n_units,
                     (uint32_t)k, out.data());
    } else {
    emul_launch((unsigned)((total + 255) / 256), 256, k_comb_recode, (const uint4*)scalars_flat.data(), total, biased.data());
    emul_launch((unsigned)((M + 63) / 64), 64, k_small_msm_comb, (const uint32_t*)biased.data(), (const uint4*)combs.data(),
                (const uint4*)shared.data(), plan.term_slot.data(), (const unsigned long long*)offsets.data(),
                (const uint32_t*)order.data(), M, (uint32_t)T, (uint32_t)U, out.data());
    }
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

static int g_emul_piece = 2;   // terms per unit of the CTA-staged comb kernel (api.cu: option prove_piece)
extern "C" {
void emul_set_piece(int piece) { g_emul_piece = piece; }
// limbs [N][p][20] (FieldElement51 X,Y,Z,T), blind [N][m][32] canonical -> com [N][k][32]; returns the not-uniform flag
// (1 if a common point differed between proofs while share was requested: the caller reruns without sharing, as api.cu does)
int emul_prove_msms(int m, int ni, int nc, int k, const int32_t* lhs, const int32_t* cons_off, const int32_t* term_scalar,
                    const int32_t* term_point, size_t N, const uint64_t* limbs, const uint8_t* blind, int share, int comb,
                    uint8_t* com) {
  pv_plan plan;
  pv_make_plan(ni, ni + nc, k, cons_off, term_point, share != 0, comb != 0, &plan);
  if (comb) pv_make_units(k, cons_off, g_emul_piece, &plan);
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
  if (comb) pv_make_units(k, cons_off, g_emul_piece, &plan);
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
void emul_msm_vartime_batched(const uint8_t* scalars, const uint8_t* points, const uint64_t* offsets, size_t M, int coop,
                              uint8_t* out, int32_t* status) {
  const size_t total = (size_t)offsets[M];
  std::vector<uint4> sc(2 * total + 2), pt(2 * total + 2), niels(ZKP_NIELS_U4 * total + ZKP_NIELS_U4), kk(2 * total + 2), k3(2 * total + 2), o(2 * M + 2);
  if (total) { memcpy(sc.data(), scalars, total * 32); memcpy(pt.data(), points, total * 32); }
  std::vector<unsigned long long> off(offsets, offsets + M + 1);
  std::vector<uint32_t> order(M + 1);
  for (size_t j = 0; j < M; j++) order[j] = (uint32_t)(M - 1 - j);   // any permutation is a valid schedule
  std::vector<int> st(M + 1, -1);
  if (total) {
    emul_launch((unsigned)((total + 255) / 256), 256, k_decompress_valid, (const uint4*)pt.data(), total, niels.data());
    emul_launch((unsigned)((total + 255) / 256), 256, k_prep_scalars_vt, (const uint4*)sc.data(), total, kk.data(), k3.data());
  }
  if (coop)   // four lanes per MSM (the latency schedule of small batches): the lanes exchange through shuffles
    emul_launch_mt((unsigned)((4 * M + 63) / 64), 64, k_small_msm_vt<true>, (const uint32_t*)kk.data(), (const uint32_t*)k3.data(),
                   (const uint4*)niels.data(), (const unsigned long long*)off.data(), (const uint32_t*)order.data(), M, o.data(),
                   st.data());
  else
    emul_launch((unsigned)((M + 63) / 64), 64, k_small_msm_vt<false>, (const uint32_t*)kk.data(), (const uint32_t*)k3.data(),
                (const uint4*)niels.data(), (const unsigned long long*)off.data(), (const uint32_t*)order.data(), M, o.data(),
                st.data());
  memcpy(out, o.data(), M * 32);
  for (size_t j = 0; j < M; j++) status[j] = st[j];
}

// zkp_msm_ct_batched (api.cu) with compressed points: k_decompress_ext + k_build_tables<false> + k_small_msm_ct<false>;
// returns 0, 1 (undecodable point) or 3 (non-canonical scalar) like the entry point
int emul_msm_ct_batched(const uint8_t* scalars, const uint8_t* points, const uint64_t* offsets, size_t M, int coop, uint8_t* out) {
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
  if (coop)
    emul_launch_mt((unsigned)((4 * M + 63) / 64), 64, k_small_msm_ct<false, true>, (const uint32_t*)biased.data(),
                   (const uint4*)tables.data(), (const unsigned long long*)off.data(), (const uint32_t*)order.data(), M, 1u,
                   o.data(), (const int32_t*)nullptr, (const uint4*)nullptr);
  else
    emul_launch((unsigned)((M + 63) / 64), 64, k_small_msm_ct<false, false>, (const uint32_t*)biased.data(),
                (const uint4*)tables.data(), (const unsigned long long*)off.data(), (const uint32_t*)order.data(), M, 1u, o.data(),
                (const int32_t*)nullptr, (const uint4*)nullptr);
  memcpy(out, o.data(), M * 32);
  if (flags[0] != 0x7fffffff) return 1;
  if (flags[1] != 0x7fffffff) return 3;
  return 0;
}

// The variable-time MSM, device-resident path (api.cu msm_vartime_launch + msm_finish with the two-phase ingestion, the
// size-ordered accumulation and the chunked bucket reduction), kernel by kernel: window width c, work-item length S
// (0 = the heuristic of api.cu).  res64 receives the msm_result (enc[32] | status | is_identity | first_bad).
// The launch sequence mirrors api.cu; grids of the grid-stride kernels are smaller (the result does not depend on them).
// host_window_scan != 0: the per-window scan (W blocks of 1024 threads, by far the slowest thing to emulate) is replaced by a
// host prefix sum with the same outputs; the single-block scans and every other kernel still run as kernels.
void emul_msm_vartime(const uint8_t* scalars, const uint8_t* points, size_t n, int c, uint32_t S_forced, int balance,
                      int host_window_scan, uint8_t* res64) {
  msm_result res;
  memset(&res, 0, sizeof res);
  if (n == 0) {
    emul_launch(1, 1, k_empty_result, &res);
    memcpy(res64, &res, sizeof res);
    return;
  }
  const int W = (253 + c - 1) / c;
  const uint32_t B = 1u << (c - 1), total_buckets = (uint32_t)W * B;
  std::vector<uint4> sc(2 * n), pt(2 * n), niels(ZKP_NIELS_U4 * n), buckets((size_t)total_buckets * 8);
  memcpy(sc.data(), scalars, n * 32);
  memcpy(pt.data(), points, n * 32);
  std::vector<uint32_t> hist((size_t)W * B, 0), offs((size_t)W * (B + 1)), cursor((size_t)W * B), sorted((size_t)W * n);
  int flags[4];
  emul_launch(1, 1, k_init_flags, flags);
  auto ingest = [&](auto kernel, size_t p_lo, size_t p_cnt, size_t s_lo, size_t s_cnt, uint32_t* counters) {
    const size_t half = (s_cnt + 1) / 2;
    size_t threads = std::max(p_cnt, std::max(half, s_cnt - half));
    if (!threads) return;
    ingest_args a = {};
    a.p_lo = p_lo; a.p_cnt = p_cnt;
    a.s_lo[0] = s_lo; a.s_cnt[0] = half;
    a.s_lo[1] = s_lo + half; a.s_cnt[1] = s_cnt - half;
    emul_launch((unsigned)((threads + 255) / 256), 256, kernel, (const uint4*)pt.data(), niels.data(), (const uint4*)sc.data(), a,
                n, c, W, B, counters, sorted.data(), flags);
  };
  const size_t half_p = (n + 1) / 2;
  ingest(k_ingest2<0, 2>, 0, half_p, 0, n, hist.data());
  if (host_window_scan) {
    for (int w = 0; w < W; w++) {
      uint32_t acc = 0;
      for (uint32_t b = 0; b < B; b++) {
        offs[(size_t)w * (B + 1) + b] = acc;
        cursor[(size_t)w * B + b] = acc;
        acc += hist[(size_t)w * B + b];
      }
      offs[(size_t)w * (B + 1) + B] = acc;
    }
  } else {
    emul_launch_mt((unsigned)W, 1024, k_scan, (const uint32_t*)hist.data(), B, offs.data(), cursor.data());
  }
  ingest(k_ingest2<1, 2>, half_p, n - half_p, 0, n, cursor.data());
  // ---- msm_finish ----
  uint32_t S = S_forced;
  if (!S) {
    const double mean = (double)n / (double)B, want = 2.0 * mean, cap = (double)n * W / 200000.0;
    double v = want < cap ? want : cap;
    if (v < 16.0) v = 16.0;
    if (v > 4096.0) v = 4096.0;
    S = (uint32_t)v;
  }
  const size_t max_items = (size_t)total_buckets + ((size_t)W * n) / S + 1;
  std::vector<uint32_t> aux0(total_buckets), aux1((size_t)total_buckets + 1), aux2(total_buckets), multi((size_t)total_buckets + 4, 0);
  std::vector<work_item> items(max_items);
  std::vector<uint4> partials(max_items * 8);
  const unsigned tb = (total_buckets + 255) / 256;
  emul_launch(tb, 256, k_plan, (const uint32_t*)offs.data(), B, total_buckets, S, aux0.data());
  if (total_buckets <= 65536) {
    emul_launch_mt(1, 1024, k_scan, (const uint32_t*)aux0.data(), total_buckets, aux1.data(), aux2.data());
  } else {
    const uint32_t tiles = (total_buckets + ZKP_SCAN_TILE - 1) / ZKP_SCAN_TILE;
    std::vector<uint32_t> tt((size_t)tiles * 3 + 8);
    emul_launch_mt(tiles, 1024, k_scan_tiles, (const uint32_t*)aux0.data(), total_buckets, aux1.data(), tt.data());
    emul_launch_mt(1, 1024, k_scan, (const uint32_t*)tt.data(), tiles, tt.data() + tiles, tt.data() + 2 * tiles + 1);
    emul_launch((total_buckets + 1023) / 1024, 1024, k_scan_add, aux1.data(), total_buckets, (const uint32_t*)(tt.data() + tiles));
  }
  emul_launch(tb, 256, k_items, (const uint32_t*)offs.data(), (const uint32_t*)aux1.data(), B, total_buckets, S, items.data(),
              multi.data() + 4, multi.data());
  const unsigned blocks = (unsigned)((max_items + 127) / 128);
  const uint32_t* n_items = aux1.data() + total_buckets;
  std::vector<uint32_t> lh(3 * (size_t)S + 8, 0), order(max_items);
  const uint32_t* ord = nullptr;
  if (balance && max_items >= 4096) {
    const unsigned ib = (unsigned)((max_items + 255) / 256);
    emul_launch(ib, 256, k_len_hist, (const work_item*)items.data(), n_items, S, lh.data());
    emul_launch_mt(1, 1024, k_scan, (const uint32_t*)lh.data(), S + 1, lh.data() + S + 1, lh.data() + 2 * S + 3);
    emul_launch(ib, 256, k_len_scatter, (const work_item*)items.data(), n_items, S, lh.data() + 2 * S + 3, order.data());
    ord = order.data();
  }
  emul_launch(blocks, 128, k_accumulate<4>, (const uint4*)niels.data(), (const uint32_t*)sorted.data(), (const work_item*)items.data(),
              ord, n_items, n, buckets.data(), partials.data());
  emul_launch_mt(2, 128, k_merge, (const uint32_t*)aux1.data(), (const uint32_t*)(multi.data() + 4), (const uint32_t*)multi.data(),
                 (const uint4*)partials.data(), buckets.data());
  const uint32_t chunks0 = B / ZKP_CHUNK_L;
  std::vector<uint4> lvlT[2], lvlU, usum((size_t)8 * W * 8);
  lvlT[0].resize((size_t)W * (chunks0 + 1) * 8);
  lvlT[1].resize((size_t)W * (chunks0 / ZKP_CHUNK_L + 1) * 8);
  lvlU.resize((size_t)W * (chunks0 + 1) * 8);
  const uint4* cur = buckets.data();
  uint32_t m = B;
  int nl = 0;
  while (m > ZKP_CHUNK_L) {
    const uint32_t chunks = m / ZKP_CHUNK_L;
    uint4* T = lvlT[nl & 1].data();
    const unsigned threads = chunks * (unsigned)W;
    emul_launch((threads + 127) / 128, 128, k_chunk_reduce, cur, m, W, T, lvlU.data());
    emul_launch_mt((unsigned)W, 256, k_tree_sum, (const uint4*)lvlU.data(), chunks, usum.data() + (size_t)nl * W * 8);
    cur = T;
    m = chunks;
    nl++;
  }
  emul_launch_mt(1, 64, k_finish, (const uint4*)usum.data(), nl, cur, m, W, c, n, (const int*)flags, &res, (uint4*)nullptr);
  memcpy(res64, &res, sizeof res);
}

// The front end of zkp_batch_verify_proofs (api.cu): transcripts, challenges, weights and coefficient fold of N proofs ->
// the MSM inputs coeff_out / points_out [nc + (ni + k) N][32].  compiled = 1: k_bv_prepare2 (host-compiled script),
// 0: k_bv_prepare (byte-wise STROBE); `chunk` proofs per launch (a multiple of 128, as in api.cu), then k_bv_static_sum.
// Returns 0, or 1 / 3 for the first identity encoding / non-canonical response (flags as k_finish reports them).
int emul_bv_front_end(int m, int ni, int nc, int k, const char* labels, const int32_t* lhs, const int32_t* cons_off,
                      const int32_t* term_scalar, const int32_t* term_point, const uint32_t* prefix53, size_t N,
                      const uint8_t* instance_enc, const uint8_t* common_enc, const uint8_t* commitments,
                      const uint8_t* responses, const uint8_t* rho_seed, int compiled, size_t chunk, uint8_t* coeff_out,
                      uint8_t* points_out, long long* first_bad) {
  zkp_statement_desc sd;
  sd.m = m; sd.ni = ni; sd.nc = nc; sd.k = k; sd.labels = labels; sd.lhs = lhs; sd.cons_off = cons_off;
  sd.term_scalar = term_scalar; sd.term_point = term_point;
  bv_plan bp;
  bv_make_plan(&sd, prefix53, rho_seed, common_enc, &bp);
  // device-side copies keep the 16-byte alignment the kernels' uint4 loads rely on
  std::vector<uint4> blob4((bp.blob.size() + 15) / 16 + 1);
  memcpy(blob4.data(), bp.blob.data(), bp.blob.size());
  const uint8_t* dm = (const uint8_t*)blob4.data();
  bv_desc d;
  bv_fill_desc(&d, &sd, bp, dm);
  const size_t rows = (size_t)ni + k, n = (size_t)nc + rows * N;
  std::vector<uint4> dsc4(2 * n + 2), dpts4(2 * n + 2), com4(2 * N * k + 2), resp4(2 * N * m + 2);
  uint8_t *dsc = (uint8_t*)dsc4.data(), *dpts = (uint8_t*)dpts4.data();
  if (nc) memcpy(dpts, common_enc, (size_t)nc * 32);
  if (ni) memcpy(dpts + (size_t)nc * 32, instance_enc, (size_t)ni * N * 32);
  if (k) memcpy(com4.data(), commitments, N * (size_t)k * 32);
  if (m) memcpy(resp4.data(), responses, N * (size_t)m * 32);
  const size_t nchunks = N ? (N + chunk - 1) / chunk : 0;
  const unsigned nblocks_total = (unsigned)(nchunks * ((chunk + 127) / 128) + 1);
  std::vector<uint4> part4((size_t)nblocks_total * (nc ? nc : 1) * 2 + 2);
  int flags[4];
  emul_launch(1, 1, k_init_flags, flags);
  unsigned block_base = 0;
  for (size_t cidx = 0; cidx < nchunks; cidx++) {
    const size_t j0 = cidx * chunk, cnt = j0 + chunk < N ? chunk : N - j0;
    unsigned nb = (unsigned)((cnt + 127) / 128);
    if (compiled && chunk > 128) nb = 1;   // a resident grid smaller than the slab: the block loops over the 128-proof groups
    if (compiled)
      emul_launch_mt(nb, 128, k_bv_prepare2, d, (const uint32_t*)(dm + bp.o_prefix), N, (const uint8_t*)(dpts + (size_t)nc * 32),
                     (const uint8_t*)com4.data(), (const uint8_t*)resp4.data(), dm + bp.o_seed, bp.script_blocks,
                     (const unsigned long long*)(dm + bp.o_tm), (const uint32_t*)(dm + bp.o_ss), (const bv_seg*)(dm + bp.o_sg),
                     dsc, dpts, (uint8_t*)part4.data(), (uint8_t*)nullptr, flags, j0, cnt, block_base);
    else
      emul_launch_mt(nb, 128, k_bv_prepare, d, (const uint32_t*)(dm + bp.o_prefix), N, (const uint8_t*)(dpts + (size_t)nc * 32),
                     (const uint8_t*)dpts, (const uint8_t*)com4.data(), (const uint8_t*)resp4.data(), dm + bp.o_seed, dsc, dpts,
                     (uint8_t*)part4.data(), (uint8_t*)nullptr, flags, j0, cnt, block_base);
    block_base += nb;
  }
  if (nc) emul_launch_mt(1, 256, k_bv_static_sum, (const uint8_t*)part4.data(), (int)block_base, nc, dsc);
  memcpy(coeff_out, dsc, n * 32);
  memcpy(points_out, dpts, n * 32);
  *first_bad = -1;
  if (flags[0] != 0x7fffffff) { *first_bad = flags[0]; return 1; }
  if (flags[1] != 0x7fffffff) { *first_bad = flags[1]; return 3; }
  return 0;
}
}
