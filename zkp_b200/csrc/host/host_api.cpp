// extern "C" face of the host mirror (include/zkp_b200_host.h).
#include <sys/random.h>

#include <atomic>
#include <chrono>
#include <new>

#include "../../../include/zkp_b200_host.h"
#include "toolbox.hpp"

using namespace zkp_host;

struct zkph_statement { Statement st; };
struct zkph_transcript { Transcript t; };

// nothing may unwind through the C face: host allocation failures and the like come back as EngineFailure
template <class F>
static int32_t guarded(F&& body) {
  try {
    return body();
  } catch (...) {
    return EngineFailure;
  }
}

// Randomness the reference draws from rand::thread_rng() (prover.rs:82, verifier.rs:153, batch_verifier.rs:179): a caller
// that passes no seed gets fresh bytes from the OS CSPRNG, per call.  A seed that is given is used as is (reproducible
// tests and benches) and must then be fresh, secret CSPRNG output itself: batch soundness rests on weights the prover
// cannot predict, and a repeated prover seed repeats nonces.
static bool os_random(uint8_t* out, size_t n) {
  while (n) {
    ssize_t got = getrandom(out, n > 256 ? 256 : n, 0);
    if (got <= 0) return false;
    out += got;
    n -= (size_t)got;
  }
  return true;
}
struct seed_arg {
  uint8_t own[32];
  const uint8_t* p;
  size_t len;
  bool ok;
  seed_arg(const uint8_t* seed, size_t seed_len) : p(seed), len(seed_len), ok(true) {
    if (!seed) {
      ok = os_random(own, 32);
      p = own;
      len = 32;
    }
  }
};

extern "C" zkph_statement* zkph_statement_new(const char* name, const char* label, const char* labels, int32_t n_secrets,
                                              int32_t n_instance, int32_t n_common, int32_t n_constraints,
                                              const int32_t* lhs, const int32_t* cons_off, const int32_t* term_scalar,
                                              const int32_t* term_point) {
  // every index of the description is checked here, once: the classes behind it index with them freely
  if (!name || !label || n_secrets < 0 || n_instance < 0 || n_common < 0 || n_constraints < 0) return nullptr;
  const int n_points = n_instance + n_common;
  if ((n_secrets + n_points) && !labels) return nullptr;
  if (n_constraints) {
    if (!lhs || !cons_off || cons_off[0] != 0) return nullptr;
    for (int i = 0; i < n_constraints; i++)
      if (lhs[i] < 0 || lhs[i] >= n_points || cons_off[i + 1] < cons_off[i]) return nullptr;
    if (cons_off[n_constraints] && (!term_scalar || !term_point)) return nullptr;
    for (int t = 0; t < cons_off[n_constraints]; t++)
      if (term_scalar[t] < 0 || term_scalar[t] >= n_secrets || term_point[t] < 0 || term_point[t] >= n_points) return nullptr;
  }
  zkph_statement* h = nullptr;
  try {
    h = new zkph_statement();
    Statement& st = h->st;
    st.name = name;
    st.label = label;
    const char* p = labels;
    for (int i = 0; i < n_secrets + n_points; i++) {
      std::string s(p);
      p += s.size() + 1;
      if (i < n_secrets) st.secrets.push_back(s);
      else if (i < n_secrets + n_instance) st.instance.push_back(s);
      else st.common.push_back(s);
    }
    for (int i = 0; i < n_constraints; i++) {
      LinComb lc;
      for (int t = cons_off[i]; t < cons_off[i + 1]; t++) lc.push_back(std::make_pair((int)term_scalar[t], (int)term_point[t]));
      st.constraints.push_back(std::make_pair((int)lhs[i], lc));
    }
  } catch (...) {
    delete h;
    return nullptr;
  }
  return h;
}
extern "C" void zkph_statement_free(zkph_statement* st) { delete st; }

static Scalar load_scalar(const uint8_t* b) { return Scalar::from_bytes_mod_order(b); }

extern "C" zkph_transcript* zkph_transcript_new(const uint8_t* label, size_t len) {
  zkph_transcript* h = new (std::nothrow) zkph_transcript();
  if (h) h->t = Transcript(label, len);
  return h;
}
extern "C" zkph_transcript* zkph_transcript_clone(const zkph_transcript* t) {
  zkph_transcript* h = new (std::nothrow) zkph_transcript();
  if (h) h->t = t->t;
  return h;
}
extern "C" void zkph_transcript_free(zkph_transcript* t) { delete t; }
extern "C" void zkph_transcript_append_message(zkph_transcript* t, const uint8_t* label, size_t llen, const uint8_t* msg,
                                               size_t mlen) {
  t->t.append_message(label, llen, msg, mlen);
}
extern "C" void zkph_transcript_export_state(const zkph_transcript* t, uint32_t out53[53]) { t->t.export_state(out53); }
extern "C" void zkph_transcript_challenge_bytes(zkph_transcript* t, const uint8_t* label, size_t llen, uint8_t* out,
                                                size_t n) {
  t->t.challenge_bytes(label, llen, out, n);
}

extern "C" int32_t zkph_prove(zkp_ctx* ctx, const zkph_statement* h, const uint8_t* tl, size_t tl_len,
                              const uint8_t* secrets, const uint64_t* points, const uint8_t* seed, size_t seed_len,
                              int32_t batchable, uint8_t* encodings, uint8_t* challenge, uint8_t* commitments,
                              uint8_t* responses, uint8_t* blindings_out) {
  return guarded([&]() -> int32_t {
  zkph_transcript tr;
  tr.t = Transcript(tl, tl_len);
  return zkph_prove_t(ctx, h, &tr, secrets, points, seed, seed_len, batchable, encodings, challenge, commitments,
                      responses, blindings_out);
  });
}

extern "C" int32_t zkph_prove_t(zkp_ctx* ctx, const zkph_statement* h, zkph_transcript* tr, const uint8_t* secrets,
                                const uint64_t* points, const uint8_t* seed, size_t seed_len, int32_t batchable,
                                uint8_t* encodings, uint8_t* challenge, uint8_t* commitments, uint8_t* responses,
                                uint8_t* blindings_out) {
  return guarded([&]() -> int32_t {
  const Statement& st = h->st;
  Transcript& t = tr->t;
  const size_t m = st.secrets.size(), p = st.num_points(), k = st.constraints.size();
  std::vector<Scalar> sec(m);
  for (size_t i = 0; i < m; i++) sec[i] = load_scalar(secrets + 32 * i);
  std::vector<Limbs> pts(p);
  for (size_t i = 0; i < p; i++) memcpy(pts[i].data(), points + 20 * i, 160);
  seed_arg sa(seed, seed_len);
  if (!sa.ok) return EngineFailure;
  Rng rng(sa.p, sa.len);
  // re-implemented inline (instead of stmt_prove) to hand the blindings back for parity tests
  Prover pr(ctx, st.label, &t);
  for (size_t i = 0; i < m; i++) pr.allocate_scalar(st.secrets[i], sec[i]);
  for (size_t i = 0; i < p; i++) {
    Enc e;
    ProofError err;
    pr.allocate_point(st.point_name(i), pts[i], &e, &err);
    if (err != PROOF_OK) return err;
    memcpy(encodings + 32 * i, e.data(), 32);
  }
  for (auto& c : st.constraints) pr.constrain(c.first, c.second);
  ProofError e;
  if (batchable) {
    BatchableProof bp;
    e = pr.prove_batchable(rng, &bp);
    if (e != PROOF_OK) return e;
    for (size_t i = 0; i < k; i++) memcpy(commitments + 32 * i, bp.commitments[i].data(), 32);
    for (size_t i = 0; i < m; i++) bp.responses[i].to_bytes(responses + 32 * i);
  } else {
    CompactProof cp;
    e = pr.prove_compact(rng, &cp);
    if (e != PROOF_OK) return e;
    cp.challenge.to_bytes(challenge);
    for (size_t i = 0; i < m; i++) cp.responses[i].to_bytes(responses + 32 * i);
  }
  if (blindings_out)
    for (size_t i = 0; i < m; i++) pr.last_blindings[i].to_bytes(blindings_out + 32 * i);
  return PROOF_OK;
  });
}

static std::vector<Enc> load_encs(const uint8_t* b, size_t n) {
  std::vector<Enc> v(n);
  for (size_t i = 0; i < n; i++) memcpy(v[i].data(), b + 32 * i, 32);
  return v;
}

extern "C" int32_t zkph_verify_compact(zkp_ctx* ctx, const zkph_statement* h, const uint8_t* tl, size_t tl_len,
                                       const uint8_t* points_enc, const uint8_t* challenge, const uint8_t* responses,
                                       size_t n_responses) {
  return guarded([&]() -> int32_t {
  zkph_transcript tr;
  tr.t = Transcript(tl, tl_len);
  return zkph_verify_compact_t(ctx, h, &tr, points_enc, challenge, responses, n_responses);
  });
}
extern "C" int32_t zkph_verify_compact_t(zkp_ctx* ctx, const zkph_statement* h, zkph_transcript* tr,
                                         const uint8_t* points_enc, const uint8_t* challenge, const uint8_t* responses,
                                         size_t n_responses) {
  return guarded([&]() -> int32_t {
  const Statement& st = h->st;
  CompactProof cp;
  // Scalar deserialisation in the reference rejects non-canonical bytes; mirror that as a failure
  if (!Scalar::from_canonical_bytes(&cp.challenge, challenge)) return VerificationFailure;
  cp.responses.resize(n_responses);
  for (size_t i = 0; i < n_responses; i++)
    if (!Scalar::from_canonical_bytes(&cp.responses[i], responses + 32 * i)) return VerificationFailure;
  return stmt_verify_compact(ctx, st, &tr->t, load_encs(points_enc, st.num_points()), cp);
  });
}

extern "C" int32_t zkph_verify_batchable(zkp_ctx* ctx, const zkph_statement* h, const uint8_t* tl, size_t tl_len,
                                         const uint8_t* points_enc, const uint8_t* commitments, size_t n_commitments,
                                         const uint8_t* responses, size_t n_responses, const uint8_t* seed,
                                         size_t seed_len) {
  return guarded([&]() -> int32_t {
  zkph_transcript tr;
  tr.t = Transcript(tl, tl_len);
  return zkph_verify_batchable_t(ctx, h, &tr, points_enc, commitments, n_commitments, responses, n_responses, seed,
                                 seed_len);
  });
}
extern "C" int32_t zkph_verify_batchable_t(zkp_ctx* ctx, const zkph_statement* h, zkph_transcript* tr,
                                           const uint8_t* points_enc, const uint8_t* commitments, size_t n_commitments,
                                           const uint8_t* responses, size_t n_responses, const uint8_t* seed,
                                           size_t seed_len) {
  return guarded([&]() -> int32_t {
  const Statement& st = h->st;
  BatchableProof bp;
  bp.commitments = load_encs(commitments, n_commitments);
  bp.responses.resize(n_responses);
  for (size_t i = 0; i < n_responses; i++)
    if (!Scalar::from_canonical_bytes(&bp.responses[i], responses + 32 * i)) return VerificationFailure;
  seed_arg sa(seed, seed_len);
  if (!sa.ok) return EngineFailure;
  Rng rng(sa.p, sa.len);
  return stmt_verify_batchable(ctx, st, &tr->t, load_encs(points_enc, st.num_points()), bp, rng);
  });
}

// module::batch_verify over N transcripts (macros.rs:336-370); identical = every transcript holds the same state
static int32_t batch_verify_impl(zkp_ctx* ctx, const Statement& st, std::vector<Transcript>* transcripts, bool identical,
                                 size_t n_proofs, const uint8_t* instance_enc, const uint8_t* common_enc,
                                 const uint8_t* commitments, const uint8_t* responses, const uint8_t* seed, size_t seed_len,
                                 int32_t threads, uint8_t* coeff_out, uint8_t* points_out, double* host_seconds) {
  const size_t N = n_proofs;
  const size_t m = st.secrets.size(), k = st.constraints.size(), ni = st.instance.size(), nc = st.common.size();
  auto t0 = std::chrono::steady_clock::now();
  std::vector<BatchableProof> proofs(N);
  std::atomic<int> noncanon(0);
  parallel_for(N, threads, [&](size_t lo, size_t hi, int) {
    for (size_t j = lo; j < hi; j++) {
      proofs[j].commitments = load_encs(commitments + j * k * 32, k);
      proofs[j].responses.resize(m);
      for (size_t i = 0; i < m; i++)
        if (!Scalar::from_canonical_bytes(&proofs[j].responses[i], responses + (j * m + i) * 32)) noncanon.store(1);
    }
  });
  if (noncanon.load()) return VerificationFailure;
  seed_arg sa(seed, seed_len);
  if (!sa.ok) return EngineFailure;
  Rng rng(sa.p, sa.len);
  ProofError err;
  BatchVerifier bv(ctx, st.label, N, transcripts, &err, threads, identical);
  if (err != PROOF_OK) return err;
  for (auto& s : st.secrets) bv.allocate_scalar(s);
  std::vector<BatchPointVar> pv;
  for (size_t i = 0; i < ni; i++) {
    pv.push_back(bv.allocate_instance_point(st.instance[i], load_encs(instance_enc + i * N * 32, N), &err));
    if (err != PROOF_OK) return err;
  }
  for (size_t i = 0; i < nc; i++) {
    Enc e;
    memcpy(e.data(), common_enc + 32 * i, 32);
    pv.push_back(bv.allocate_static_point(st.common[i], e, &err));
    if (err != PROOF_OK) return err;
  }
  for (auto& c : st.constraints) {
    BatchLinComb lc;
    for (auto& term : c.second) lc.push_back(std::make_pair(term.first, pv[term.second]));
    bv.constrain(pv[c.first], lc);
  }
  std::vector<uint8_t> sc, ic, ip;
  err = bv.batch_coeffs(proofs, rng, threads, &sc, &ic, &ip);
  if (err != PROOF_OK) return err;
  if (host_seconds) *host_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::vector<uint8_t> sp(nc * 32);
  for (size_t s = 0; s < nc; s++) memcpy(&sp[32 * s], bv.static_points()[s].data(), 32);
  if (coeff_out) {
    memcpy(coeff_out, sc.data(), sc.size());
    memcpy(coeff_out + sc.size(), ic.data(), ic.size());
  }
  if (points_out) {
    memcpy(points_out, sp.data(), sp.size());
    memcpy(points_out + sp.size(), ip.data(), ip.size());
  }
  int32_t accept = 0;
  int64_t bad = -1;
  int32_t rc = zkp_batch_verify(ctx, sc.data(), sp.data(), nc, ic.data(), ip.data(), bv.rows(), N, &accept, &bad);
  if (rc == ZKP_ERR_POINT) return VerificationFailure;
  if (rc != ZKP_OK) return EngineFailure;
  return accept ? PROOF_OK : VerificationFailure;
}

extern "C" int32_t zkph_batch_verify(zkp_ctx* ctx, const zkph_statement* h, const uint8_t* tl, size_t tl_len, size_t N,
                                     const uint8_t* instance_enc, const uint8_t* common_enc, const uint8_t* commitments,
                                     const uint8_t* responses, const uint8_t* seed, size_t seed_len, int32_t threads,
                                     uint8_t* coeff_out, uint8_t* points_out, double* host_seconds) {
  return guarded([&]() -> int32_t {
  std::vector<Transcript> transcripts(N, Transcript(tl, tl_len));
  return batch_verify_impl(ctx, h->st, &transcripts, /*identical=*/true, N, instance_enc, common_enc, commitments, responses,
                           seed, seed_len, threads, coeff_out, points_out, host_seconds);
  });
}

// The reference's signature: one `&mut Transcript` per proof, each with its own prior state, each left advanced
// (batch_verifier.rs:67-88; macros.rs:336-346).  n_transcripts != n_proofs is BatchSizeMismatch (batch_verifier.rs:72-74).
extern "C" int32_t zkph_batch_verify_t(zkp_ctx* ctx, const zkph_statement* h, zkph_transcript* const* transcripts,
                                       size_t n_transcripts, size_t n_proofs, const uint8_t* instance_enc,
                                       const uint8_t* common_enc, const uint8_t* commitments, const uint8_t* responses,
                                       const uint8_t* seed, size_t seed_len, int32_t threads) {
  return guarded([&]() -> int32_t {
  if (n_transcripts != n_proofs) return BatchSizeMismatch;
  for (size_t j = 0; j < n_transcripts; j++)
    if (!transcripts || !transcripts[j]) return EngineFailure;
  std::vector<Transcript> ts(n_transcripts);
  for (size_t j = 0; j < n_transcripts; j++) ts[j] = transcripts[j]->t;
  int32_t rc = batch_verify_impl(ctx, h->st, &ts, /*identical=*/false, n_proofs, instance_enc, common_enc, commitments,
                                 responses, seed, seed_len, threads, nullptr, nullptr, nullptr);
  for (size_t j = 0; j < n_transcripts; j++) transcripts[j]->t = ts[j];   // advanced, like the reference's &mut borrows
  return rc;
  });
}

// statement -> flat descriptor of zkp_b200.h (labels of instance ++ common points, constraint CSR)
struct flat_stmt {
  std::string labels;
  std::vector<int32_t> lhs, off, ts, tp;
  zkp_statement_desc d;
  explicit flat_stmt(const Statement& st) : off(1, 0) {
    for (auto& s : st.instance) labels += s + std::string(1, '\0');
    for (auto& s : st.common) labels += s + std::string(1, '\0');
    for (auto& c : st.constraints) {
      lhs.push_back(c.first);
      for (auto& term : c.second) {
        ts.push_back(term.first);
        tp.push_back(term.second);
      }
      off.push_back((int32_t)ts.size());
    }
    d.m = (int32_t)st.secrets.size();
    d.ni = (int32_t)st.instance.size();
    d.nc = (int32_t)st.common.size();
    d.k = (int32_t)st.constraints.size();
    d.labels = labels.c_str();
    d.lhs = lhs.data();
    d.cons_off = off.data();
    d.term_scalar = ts.data();
    d.term_point = tp.data();
  }
};

extern "C" int32_t zkph_batch_verify_device(zkp_ctx* ctx, const zkph_statement* h, const uint8_t* tl, size_t tl_len,
                                            size_t N, const uint8_t* instance_enc, const uint8_t* common_enc,
                                            const uint8_t* commitments, const uint8_t* responses,
                                            const uint8_t* rho_seed32, uint8_t* coeff_out, uint8_t* points_out) {
  return guarded([&]() -> int32_t {
  const Statement& st = h->st;
  // batch-wide transcript prefix: Transcript::new(label), dom-sep, scalar labels (macros.rs:346-350)
  Transcript t(tl, tl_len);
  domain_sep(t, st.label);
  for (auto& s : st.secrets) append_scalar_var(t, s);
  uint32_t prefix[53];
  t.export_state(prefix);
  flat_stmt f(st);
  int32_t accept = 0;
  int64_t bad = -1;
  seed_arg sa(rho_seed32, 32);     // no seed: 32 fresh bytes from the OS CSPRNG
  if (!sa.ok) return EngineFailure;
  int32_t rc = zkp_batch_verify_proofs(ctx, &f.d, prefix, N, instance_enc, common_enc, commitments, responses, sa.p,
                                       &accept, &bad, coeff_out, points_out);
  if (rc == ZKP_ERR_POINT || rc == ZKP_ERR_SCALAR) return VerificationFailure;
  if (rc == ZKP_ERR_SIZE) return BatchSizeMismatch;
  if (rc != ZKP_OK) return EngineFailure;
  return accept ? PROOF_OK : VerificationFailure;
  });
}

extern "C" int32_t zkph_prove_many_device(zkp_ctx* ctx, const zkph_statement* h, const uint8_t* tl, size_t tl_len, size_t N,
                                          const uint8_t* secrets, const uint64_t* points, const uint8_t* entropy,
                                          uint8_t* encodings, uint8_t* commitments, uint8_t* responses) {
  return guarded([&]() -> int32_t {
  const Statement& st = h->st;
  // batch-wide transcript prefix: Transcript::new(label), dom-sep, scalar labels (macros.rs:206-214)
  Transcript t(tl, tl_len);
  domain_sep(t, st.label);
  for (auto& s : st.secrets) append_scalar_var(t, s);
  uint32_t prefix[53];
  t.export_state(prefix);
  flat_stmt f(st);
  std::vector<uint8_t> own_entropy;
  if (!entropy && N) {             // no entropy: 32 fresh bytes per proof from the OS CSPRNG (prover.rs:82 thread_rng)
    own_entropy.resize(N * 32);
    if (!os_random(own_entropy.data(), own_entropy.size())) return EngineFailure;
    entropy = own_entropy.data();
  }
  int32_t rc = zkp_prove_batch(ctx, &f.d, prefix, N, secrets, points, entropy, encodings, commitments, responses, nullptr);
  if (rc == ZKP_ERR_SIZE) return BatchSizeMismatch;
  if (rc != ZKP_OK) return EngineFailure;
  return PROOF_OK;
  });
}

extern "C" int32_t zkph_prove_many(zkp_ctx* ctx, const zkph_statement* h, const uint8_t* tl, size_t tl_len, size_t N,
                                   const uint8_t* secrets, const uint64_t* points, const uint8_t* entropy,
                                   int32_t threads, uint8_t* encodings, uint8_t* commitments, uint8_t* responses) {
  return guarded([&]() -> int32_t {
  const Statement& st = h->st;
  const size_t m = st.secrets.size(), p = st.num_points(), k = st.constraints.size();
  std::vector<Scalar> sec(N * m);
  parallel_for(N * m, threads, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; i++) sec[i] = load_scalar(secrets + 32 * i);
  });
  std::vector<uint8_t> own_entropy;
  if (!entropy && N) {
    own_entropy.resize(N * 32);
    if (!os_random(own_entropy.data(), own_entropy.size())) return EngineFailure;
    entropy = own_entropy.data();
  }
  std::vector<BatchableProof> proofs;
  std::vector<Enc> encs;
  ProofError e = stmt_prove_many(ctx, st, std::string((const char*)tl, tl_len), N, sec.data(), (const Limbs*)points,
                                 entropy, threads, &proofs, &encs);
  if (e != PROOF_OK) return e;
  memcpy(encodings, encs.data(), N * p * 32);
  parallel_for(N, threads, [&](size_t lo, size_t hi, int) {
    for (size_t j = lo; j < hi; j++) {
      for (size_t i = 0; i < k; i++) memcpy(commitments + (j * k + i) * 32, proofs[j].commitments[i].data(), 32);
      for (size_t i = 0; i < m; i++) proofs[j].responses[i].to_bytes(responses + (j * m + i) * 32);
    }
  });
  return PROOF_OK;
  });
}

extern "C" void zkph_scalar_mul(uint8_t* out32, const uint8_t* a32, const uint8_t* b32) {
  sc_mul(Scalar::from_bytes_mod_order(a32), Scalar::from_bytes_mod_order(b32)).to_bytes(out32);
}
extern "C" void zkph_scalar_from_wide(uint8_t* out32, const uint8_t* in64) {
  Scalar::from_bytes_mod_order_wide(in64).to_bytes(out32);
}
extern "C" void zkph_merlin_test_vector(uint8_t* out32) {
  Transcript t((const uint8_t*)"test protocol", 13);
  t.append_message((const uint8_t*)"step1", 5, (const uint8_t*)"some data", 9);
  uint8_t ch[32], big[1024];
  memset(big, 0x63, 1024);
  for (int i = 0; i < 32; i++) {
    t.challenge_bytes((const uint8_t*)"challenge", 9, ch, 32);
    t.append_message((const uint8_t*)"bigdata", 7, big, 1024);
    t.append_message((const uint8_t*)"challengedata", 13, ch, 32);
  }
  memcpy(out32, ch, 32);
}
extern "C" void zkph_rng_bytes(const uint8_t* seed, size_t seed_len, uint8_t* out, size_t n) {
  Rng r(seed, seed_len);
  r.bytes(out, n);
}
