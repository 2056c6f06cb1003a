// Host-side plan of the batch prover's MSMs (zkp_prove_batch): everything here depends on the statement only (public).
// Shared by api.cu and by the host emulation of the kernels (tests/host_emul/kernels_emul.cpp), so the CPU tests run the
// plan the device gets.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <vector>

namespace zkp {

struct pv_plan {
  std::vector<int32_t> cons_slot;          // [k] position of the constraint in the size-ordered MSM schedule
  std::vector<int32_t> term_shared;        // [n_terms] index of the common point a term uses when its table is shared, else -1
  // comb path (comb.cuh): one comb per base
  std::vector<int32_t> term_slot;          // [n_terms] u >= 0: per-proof comb slot u; -(s + 1): shared comb s
  std::vector<int32_t> comb_slot_point;    // [U]  point index (instance ++ common) of per-proof slot u
  std::vector<int32_t> comb_shared_point;  // [Us] point index of shared slot s
  // CTA-staged comb kernel (k_comb_msm_cta): every constraint is cut into units of at most `piece` consecutive terms, so
  // that the warps of a CTA (one unit each per round) do equal work whatever the constraint sizes are
  std::vector<int32_t> unit_term0;         // [n_units] first term (index within the proof's term list) of the unit
  std::vector<int32_t> unit_nterms;        // [n_units] 1 .. piece
  std::vector<int32_t> cons_unit0;         // [k + 1] the units of constraint c are cons_unit0[c] .. cons_unit0[c + 1] - 1
};

// units of at most `piece` terms per constraint (a constraint without terms keeps one empty unit: it commits to the identity)
static inline void pv_make_units(int k, const int32_t* cons_off, int piece, pv_plan* pl) {
  pl->unit_term0.clear();
  pl->unit_nterms.clear();
  pl->cons_unit0.assign(k + 1, 0);
  if (piece < 1) piece = 1;
  for (int c = 0; c < k; c++) {
    pl->cons_unit0[c] = (int32_t)pl->unit_term0.size();
    const int lo = cons_off[c], hi = cons_off[c + 1];
    if (lo == hi) {
      pl->unit_term0.push_back(lo);
      pl->unit_nterms.push_back(0);
    }
    for (int q = lo; q < hi; q += piece) {
      pl->unit_term0.push_back(q);
      pl->unit_nterms.push_back(hi - q < piece ? hi - q : piece);
    }
  }
  pl->cons_unit0[k] = (int32_t)pl->unit_term0.size();
}

// k constraints with term ranges cons_off[k + 1], term_point[n_terms] over instance (ni) ++ common points.
// share: the batch-static (common) points get one table / comb per batch; comb: fill the comb slots as well.
static inline void pv_make_plan(int ni, int p, int k, const int32_t* cons_off, const int32_t* term_point, bool share,
                                bool comb, pv_plan* pl) {
  const int n_terms = k ? cons_off[k] : 0;
  // MSM schedule: constraints ordered by (public) size, largest first, so that the lanes of a warp do equal work
  std::vector<int32_t> by_size(k);
  pl->cons_slot.assign(k, 0);
  for (int c = 0; c < k; c++) by_size[c] = c;
  std::stable_sort(by_size.begin(), by_size.end(),
                   [&](int a, int b) { return cons_off[a + 1] - cons_off[a] > cons_off[b + 1] - cons_off[b]; });
  for (int r = 0; r < k; r++) pl->cons_slot[by_size[r]] = r;
  pl->term_shared.assign(n_terms, -1);
  if (share)
    for (int q = 0; q < n_terms; q++)
      if (term_point[q] >= ni) pl->term_shared[q] = term_point[q] - ni;
  pl->term_slot.assign(n_terms, 0);
  pl->comb_slot_point.clear();
  pl->comb_shared_point.clear();
  if (!comb) return;
  std::vector<int32_t> slot_of(p, 0);   // 0 = unused, u + 1 = per-proof slot u, -(s + 1) = shared slot s
  for (int q = 0; q < n_terms; q++) {
    const int b = term_point[q];
    if (!slot_of[b]) {
      if (share && b >= ni) {
        pl->comb_shared_point.push_back(b);
        slot_of[b] = -(int32_t)pl->comb_shared_point.size();
      } else {
        pl->comb_slot_point.push_back(b);
        slot_of[b] = (int32_t)pl->comb_slot_point.size();
      }
    }
    pl->term_slot[q] = slot_of[b] > 0 ? slot_of[b] - 1 : slot_of[b];
  }
}

}  // namespace zkp
