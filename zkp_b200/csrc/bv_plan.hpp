// Host-side plan of the batch-verification front end (zkp_batch_verify_proofs): the symbolic STROBE that compiles a
// batch's transcript script, and the small blob the device kernels read the statement from.  Everything here depends on
// the statement, the batch-wide transcript prefix and the static points only (public).  Shared by api.cu and by the host
// emulation of the kernels (tests/host_emul/kernels_emul.cpp), so the CPU tests run the plan the device gets.
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/zkp_b200.h"
#include "bv_kernels.cuh"

namespace zkp {

// ---------------------------------------------------------------------------------------------------------
// Symbolic STROBE-128 over the per-proof transcript script of a batch (bv_kernels.cuh k_bv_prepare2): records the
// constant bytes of every rate block and where the per-proof 32-byte values land.  Mirrors merlin 2.0 strobe.rs [ext]
// exactly as csrc/hash.cuh and host/merlin.cpp do (begin_op framing, run_f padding, forced run_f of C-flagged ops).
// ---------------------------------------------------------------------------------------------------------
struct bv_script {
  static const uint32_t R = 166;
  uint32_t pos, pos_begin;
  std::vector<uint8_t> cur;                 // 168 bytes of the block under construction
  std::vector<uint64_t> tmpl;               // [nblocks][21]
  std::vector<uint32_t> seg_start;          // [nblocks + 1]
  std::vector<bv_seg> segs;
  bv_script(uint32_t p, uint32_t pb) : pos(p), pos_begin(pb), cur(168, 0), seg_start(1, 0) {}
  void run_f() {
    cur[pos] ^= (uint8_t)pos_begin;
    cur[pos + 1] ^= 0x04;
    cur[R + 1] ^= 0x80;
    for (int l = 0; l < 21; l++) {
      uint64_t v = 0;
      for (int b = 7; b >= 0; b--) v = (v << 8) | cur[8 * l + b];
      tmpl.push_back(v);
    }
    seg_start.push_back((uint32_t)segs.size());
    std::fill(cur.begin(), cur.end(), 0);
    pos = 0;
    pos_begin = 0;
  }
  void absorb(const uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
      cur[pos] ^= d[i];
      if (++pos == R) run_f();
    }
  }
  void absorb_value(uint32_t kind, uint32_t idx) {   // 32 per-proof bytes, split at block boundaries
    uint32_t done = 0;
    while (done < 32) {
      const uint32_t room = R - pos, take = 32 - done < room ? 32 - done : room;
      bv_seg sg;
      memset(&sg, 0, sizeof sg);
      sg.kind = kind; sg.idx = idx; sg.src_off = done; sg.len = take; sg.shift = (int32_t)pos - (int32_t)done;
      segs.push_back(sg);
      pos += take;
      done += take;
      if (pos == R) run_f();
    }
  }
  void begin_op(uint32_t flags, bool more) {
    if (more) return;
    uint8_t hdr[2] = {(uint8_t)pos_begin, (uint8_t)flags};
    pos_begin = pos + 1;
    absorb(hdr, 2);
    if ((flags & (ZKP_FLAG_C | ZKP_FLAG_K)) && pos != 0) run_f();
  }
  void meta_ad(const uint8_t* d, size_t n, bool more) { begin_op(ZKP_FLAG_M | ZKP_FLAG_A, more); absorb(d, n); }
  void append_header(const char* label, const uint8_t* l2, size_t l2len) {   // append_message(label, <l2len bytes>) minus the data
    uint8_t l4[4] = {(uint8_t)l2len, (uint8_t)(l2len >> 8), (uint8_t)(l2len >> 16), (uint8_t)(l2len >> 24)};
    (void)l2;
    meta_ad((const uint8_t*)label, strlen(label), false);
    meta_ad(l4, 4, true);
    begin_op(ZKP_FLAG_A, false);
  }
  void append_const(const char* label, const uint8_t* msg, size_t n) { append_header(label, msg, n); absorb(msg, n); }
  void append_value(const char* label, uint32_t kind, uint32_t idx) { append_header(label, nullptr, 32); absorb_value(kind, idx); }
  void challenge(const char* label, uint32_t n) {   // up to the forced run_f of prf; the squeeze reads the fresh state
    uint8_t l4[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    meta_ad((const uint8_t*)label, strlen(label), false);
    meta_ad(l4, 4, true);
    // the two header bytes leave pos != 0 (forced run_f) or end exactly on the block boundary (run_f inside absorb):
    // either way the block is closed here and the squeeze starts at byte 0 of the permuted state
    begin_op(ZKP_FLAG_I | ZKP_FLAG_A | ZKP_FLAG_C, false);
  }
};

// the per-proof script of BatchVerifier (batch_verifier.rs:100-134, :152-167): instance points, static points, commitments,
// challenge -- in the allocation order of the define_proof! expansion (macros.rs:348-365)
static inline void bv_compile_batch_verify(bv_script& script, const zkp_statement_desc* sd, const uint8_t* common_enc) {
  const int ni = sd->ni, nc = sd->nc, k = sd->k;
  const char* q = sd->labels;
  std::vector<std::string> names;
  for (int i = 0; i < ni + nc; i++) {
    names.push_back(std::string(q));
    q += names.back().size() + 1;
  }
  for (int i = 0; i < ni; i++) {
    script.append_const("ptvar", (const uint8_t*)names[i].data(), names[i].size());
    script.append_value("val", 0u, (uint32_t)i);
  }
  for (int i = 0; i < nc; i++) {
    script.append_const("ptvar", (const uint8_t*)names[ni + i].data(), names[ni + i].size());
    script.append_const("val", common_enc + 32 * (size_t)i, 32);
  }
  for (int c = 0; c < k; c++) {
    const std::string& nm = names[sd->lhs[c]];
    script.append_const("blindcom", (const uint8_t*)nm.data(), nm.size());
    script.append_value("val", 2u, (uint32_t)c);
  }
  script.challenge("chal", 64);
}


struct bv_plan {
  std::vector<uint8_t> blob;   // prefix(53 w) | rho_seed(32 B) | label pool | ops | int arrays | block templates | segments
  size_t o_prefix, o_seed, o_pool, o_ops, o_lk, o_li, o_co, o_ts, o_tk, o_ti, o_tm, o_ss, o_sg;
  int n_ops, n_terms, script_blocks;
};

// sd must be consistent (api.cu statement_ok); common_enc may be null when the statement has no common points
static inline void bv_make_plan(const zkp_statement_desc* sd, const uint32_t* prefix_state, const uint8_t* rho_seed,
                                const uint8_t* common_enc, bv_plan* pl) {
  const int ni = sd->ni, nc = sd->nc, k = sd->k;
  std::vector<uint32_t> loff, llen;
  const char* lp = sd->labels;
  std::vector<uint8_t> pool;
  for (int i = 0; i < ni + nc; i++) {
    size_t len = strlen(lp);
    loff.push_back((uint32_t)pool.size());
    llen.push_back((uint32_t)len);
    pool.insert(pool.end(), lp, lp + len);
    lp += len + 1;
  }
  const int n_terms = k ? sd->cons_off[k] : 0;
  std::vector<bv_op> ops;
  for (int i = 0; i < ni; i++) ops.push_back(bv_op{0u, loff[i], llen[i], (uint32_t)i});
  for (int i = 0; i < nc; i++) ops.push_back(bv_op{1u, loff[ni + i], llen[ni + i], (uint32_t)i});
  std::vector<int32_t> lhs_kind(k), lhs_idx(k), tkind(n_terms), tidx(n_terms);
  for (int c = 0; c < k; c++) {
    const int l = sd->lhs[c];
    lhs_kind[c] = l >= ni;
    lhs_idx[c] = l >= ni ? l - ni : l;
    ops.push_back(bv_op{2u, loff[l], llen[l], (uint32_t)c});
  }
  for (int q = 0; q < n_terms; q++) {
    const int pnt = sd->term_point[q];
    tkind[q] = pnt >= ni;
    tidx[q] = pnt >= ni ? pnt - ni : pnt;
  }
  // the compiled transcript script (k_bv_prepare2): allocation order instance, static, then the commitments
  bv_script script(prefix_state[50], prefix_state[51]);
  bv_compile_batch_verify(script, sd, common_enc);
  pl->script_blocks = (int)(script.tmpl.size() / 21);
  pl->n_ops = (int)ops.size();
  pl->n_terms = n_terms;
  auto pad16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  pl->o_prefix = 0;
  pl->o_seed = pad16(53 * 4);
  pl->o_pool = pl->o_seed + 32;
  pl->o_ops = pad16(pl->o_pool + pool.size());
  pl->o_lk = pad16(pl->o_ops + ops.size() * sizeof(bv_op));
  pl->o_li = pad16(pl->o_lk + (size_t)k * 4);
  pl->o_co = pad16(pl->o_li + (size_t)k * 4);
  pl->o_ts = pad16(pl->o_co + (size_t)(k + 1) * 4);
  pl->o_tk = pad16(pl->o_ts + (size_t)n_terms * 4);
  pl->o_ti = pad16(pl->o_tk + (size_t)n_terms * 4);
  pl->o_tm = pad16(pl->o_ti + (size_t)n_terms * 4);
  pl->o_ss = pad16(pl->o_tm + script.tmpl.size() * 8);
  pl->o_sg = pad16(pl->o_ss + script.seg_start.size() * 4);
  const size_t blob_sz = pad16(pl->o_sg + script.segs.size() * sizeof(bv_seg)) + 16;
  std::vector<uint8_t>& blob = pl->blob;
  blob.assign(blob_sz, 0);
  if (!script.tmpl.empty()) memcpy(&blob[pl->o_tm], script.tmpl.data(), script.tmpl.size() * 8);
  memcpy(&blob[pl->o_ss], script.seg_start.data(), script.seg_start.size() * 4);
  if (!script.segs.empty()) memcpy(&blob[pl->o_sg], script.segs.data(), script.segs.size() * sizeof(bv_seg));
  memcpy(&blob[pl->o_prefix], prefix_state, 53 * 4);
  memcpy(&blob[pl->o_seed], rho_seed, 32);
  if (!pool.empty()) memcpy(&blob[pl->o_pool], pool.data(), pool.size());
  if (!ops.empty()) memcpy(&blob[pl->o_ops], ops.data(), ops.size() * sizeof(bv_op));
  if (k) {
    memcpy(&blob[pl->o_lk], lhs_kind.data(), (size_t)k * 4);
    memcpy(&blob[pl->o_li], lhs_idx.data(), (size_t)k * 4);
    memcpy(&blob[pl->o_co], sd->cons_off, (size_t)(k + 1) * 4);
  }
  if (n_terms) {
    memcpy(&blob[pl->o_ts], sd->term_scalar, (size_t)n_terms * 4);
    memcpy(&blob[pl->o_tk], tkind.data(), (size_t)n_terms * 4);
    memcpy(&blob[pl->o_ti], tidx.data(), (size_t)n_terms * 4);
  }
}

// the statement descriptor the kernels take, pointing into a copy of the blob at `base` (device or host address)
static inline void bv_fill_desc(bv_desc* d, const zkp_statement_desc* sd, const bv_plan& pl, const uint8_t* base) {
  d->m = sd->m; d->ni = sd->ni; d->nc = sd->nc; d->k = sd->k; d->n_ops = pl.n_ops; d->n_terms = pl.n_terms;
  d->ops = (const bv_op*)(base + pl.o_ops);
  d->labels = base + pl.o_pool;
  d->lhs_kind = (const int32_t*)(base + pl.o_lk);
  d->lhs_idx = (const int32_t*)(base + pl.o_li);
  d->cons_off = (const int32_t*)(base + pl.o_co);
  d->term_scalar = (const int32_t*)(base + pl.o_ts);
  d->term_pkind = (const int32_t*)(base + pl.o_tk);
  d->term_pidx = (const int32_t*)(base + pl.o_ti);
}

}  // namespace zkp
