#include "toolbox.hpp"

#include <exception>
#include <functional>
#include <thread>

namespace zkp_host {

// ---- helpers ---------------------------------------------------------------------------------------------
static inline const uint8_t* U8(const std::string& s) { return (const uint8_t*)s.data(); }

void parallel_for(size_t n, int threads, const std::function<void(size_t, size_t, int)>& fn) {
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads < 1) threads = 1;
  if ((size_t)threads > n) threads = n ? (int)n : 1;
  if (threads == 1) {
    fn(0, n, 0);
    return;
  }
  // an exception inside a worker (a failed allocation) is carried back to the caller instead of ending the process
  std::vector<std::thread> pool;
  std::vector<std::exception_ptr> failed(threads);
  size_t per = n / threads, rem = n % threads, lo = 0;
  for (int t = 0; t < threads; t++) {
    size_t cnt = per + ((size_t)t < rem ? 1 : 0);
    const size_t a = lo, b = lo + cnt;
    try {
      pool.emplace_back([&fn, &failed, a, b, t]() {
        try {
          fn(a, b, t);
        } catch (...) {
          failed[t] = std::current_exception();
        }
      });
    } catch (...) {   // no more threads: this range runs here
      try {
        fn(a, b, t);
      } catch (...) {
        failed[t] = std::current_exception();
      }
    }
    lo += cnt;
  }
  for (auto& th : pool) th.join();
  for (auto& f : failed)
    if (f) std::rethrow_exception(f);
}

Rng::Rng(const uint8_t* seed, size_t len) {
  memset(st_, 0, sizeof st_);
  uint8_t* b = (uint8_t*)st_;
  size_t pos = 0;
  for (size_t i = 0; i < len; i++) {
    b[pos++] ^= seed[i];
    if (pos == 136) { keccak_f1600(st_); pos = 0; }
  }
  b[pos] ^= 0x1F;
  b[135] ^= 0x80;
  keccak_f1600(st_);
  pos_ = 0;
}
void Rng::bytes(uint8_t* out, size_t n) {
  const uint8_t* b = (const uint8_t*)st_;
  for (size_t i = 0; i < n; i++) {
    if (pos_ == 136) { keccak_f1600(st_); pos_ = 0; }
    out[i] = b[pos_++];
  }
}
Scalar Rng::u128() {
  uint8_t b[16];
  bytes(b, 16);
  uint64_t lo, hi;
  memcpy(&lo, b, 8);
  memcpy(&hi, b + 8, 8);
  return Scalar::from_u128(lo, hi);
}

static bool is_identity_encoding(const Enc& e) {
  uint8_t o = 0;
  for (int i = 0; i < 32; i++) o |= e[i];
  return o == 0;
}

// ---- TranscriptProtocol -----------------------------------------------------------------------------------
void domain_sep(Transcript& t, const std::string& label) {
  static const char* P = "schnorrzkp/1.0/ristretto255";
  t.append_message((const uint8_t*)"dom-sep", 7, (const uint8_t*)P, strlen(P));
  t.append_message((const uint8_t*)"dom-sep", 7, U8(label), label.size());
}
void append_scalar_var(Transcript& t, const std::string& label) {
  t.append_message((const uint8_t*)"scvar", 5, U8(label), label.size());
}
void append_point_var(Transcript& t, const std::string& label, const Enc& enc) {
  t.append_message((const uint8_t*)"ptvar", 5, U8(label), label.size());
  t.append_message((const uint8_t*)"val", 3, enc.data(), 32);
}
bool validate_and_append_point_var(Transcript& t, const std::string& label, const Enc& enc) {
  if (is_identity_encoding(enc)) return false;
  append_point_var(t, label, enc);
  return true;
}
void append_blinding_commitment(Transcript& t, const std::string& label, const Enc& enc) {
  t.append_message((const uint8_t*)"blindcom", 8, U8(label), label.size());
  t.append_message((const uint8_t*)"val", 3, enc.data(), 32);
}
bool validate_and_append_blinding_commitment(Transcript& t, const std::string& label, const Enc& enc) {
  if (is_identity_encoding(enc)) return false;
  append_blinding_commitment(t, label, enc);
  return true;
}
Scalar get_challenge(Transcript& t, const std::string& label) {
  uint8_t b[64];
  t.challenge_bytes(U8(label), label.size(), b, 64);
  return Scalar::from_bytes_mod_order_wide(b);
}

// ---- Prover ------------------------------------------------------------------------------------------------
Prover::Prover(zkp_ctx* ctx, const std::string& proof_label, Transcript* transcript) : ctx_(ctx), transcript_(transcript) {
  domain_sep(*transcript_, proof_label);
}
int Prover::allocate_scalar(const std::string& label, const Scalar& assignment) {
  append_scalar_var(*transcript_, label);
  scalars_.push_back(assignment);
  return (int)scalars_.size() - 1;
}
int Prover::allocate_point(const std::string& label, const Limbs& assignment, Enc* enc_out, ProofError* err) {
  Enc enc;
  if (zkp_compress_batch(ctx_, assignment.data(), 1, enc.data()) != ZKP_OK) {
    if (err) *err = EngineFailure;
    return -1;
  }
  append_point_var(*transcript_, label, enc);
  points_.push_back(assignment);
  point_labels_.push_back(label);
  if (enc_out) *enc_out = enc;
  if (err) *err = PROOF_OK;
  return (int)points_.size() - 1;
}

static void draw_blindings(const Transcript& t, const std::vector<Scalar>& secrets, const uint8_t entropy[32],
                           std::vector<Scalar>* blindings) {
  TranscriptRngBuilder b = t.build_rng();
  for (const Scalar& s : secrets) {
    uint8_t sb[32];
    s.to_bytes(sb);
    b.rekey_with_witness_bytes((const uint8_t*)"", 0, sb, 32);
  }
  TranscriptRng trng = b.finalize(entropy);
  blindings->clear();
  for (size_t i = 0; i < secrets.size(); i++) {
    uint8_t wide[64];
    trng.fill_bytes(wide, 64);
    blindings->push_back(Scalar::from_bytes_mod_order_wide(wide));
  }
}

ProofError Prover::prove_impl(Rng& rng, Scalar* challenge, std::vector<Scalar>* responses, std::vector<Enc>* commitments) {
  uint8_t entropy[32];
  rng.bytes(entropy, 32);
  std::vector<Scalar> blindings;
  draw_blindings(*transcript_, scalars_, entropy, &blindings);
  last_blindings = blindings;
  // all constraints in ONE constant-time batched call (prover.rs:93-97 loops; the commitments are independent)
  std::vector<uint8_t> sc;
  std::vector<uint64_t> pts, offsets(1, 0);
  for (auto& c : constraints_) {
    for (auto& term : c.second) {
      uint8_t b[32];
      blindings[term.first].to_bytes(b);
      sc.insert(sc.end(), b, b + 32);
      pts.insert(pts.end(), points_[term.second].begin(), points_[term.second].end());
    }
    offsets.push_back(sc.size() / 32);
  }
  size_t k = constraints_.size();
  std::vector<uint8_t> out(k * 32 + 32);
  if (k && zkp_msm_ct_batched(ctx_, sc.data(), pts.data(), ZKP_POINTS_LIMBS51, offsets.data(), k, out.data()) != ZKP_OK)
    return EngineFailure;
  commitments->clear();
  for (size_t i = 0; i < k; i++) {
    Enc e;
    memcpy(e.data(), out.data() + 32 * i, 32);
    append_blinding_commitment(*transcript_, point_labels_[constraints_[i].first], e);
    commitments->push_back(e);
  }
  *challenge = get_challenge(*transcript_, "chal");
  responses->clear();
  for (size_t i = 0; i < scalars_.size(); i++) responses->push_back(sc_muladd(scalars_[i], *challenge, blindings[i]));
  return PROOF_OK;
}
ProofError Prover::prove_compact(Rng& rng, CompactProof* out) {
  std::vector<Enc> coms;
  return prove_impl(rng, &out->challenge, &out->responses, &coms);
}
ProofError Prover::prove_batchable(Rng& rng, BatchableProof* out) {
  Scalar c;
  return prove_impl(rng, &c, &out->responses, &out->commitments);
}

// ---- Verifier ----------------------------------------------------------------------------------------------
Verifier::Verifier(zkp_ctx* ctx, const std::string& proof_label, Transcript* transcript)
    : ctx_(ctx), transcript_(transcript), num_scalars_(0) {
  domain_sep(*transcript_, proof_label);
}
int Verifier::allocate_scalar(const std::string& label) {
  append_scalar_var(*transcript_, label);
  return num_scalars_++;
}
int Verifier::allocate_point(const std::string& label, const Enc& assignment, ProofError* err) {
  if (!validate_and_append_point_var(*transcript_, label, assignment)) {
    if (err) *err = VerificationFailure;
    return -1;
  }
  points_.push_back(assignment);
  point_labels_.push_back(label);
  if (err) *err = PROOF_OK;
  return (int)points_.size() - 1;
}

ProofError Verifier::verify_compact(const CompactProof& proof) {
  if ((int)proof.responses.size() != num_scalars_) return VerificationFailure;
  // "Decompress all parameters or fail verification" (verifier.rs:87-92)
  size_t np = points_.size();
  if (np) {
    std::vector<uint64_t> limbs(np * 20);
    std::vector<uint8_t> valid(np);
    if (zkp_decompress_batch(ctx_, points_[0].data(), np, limbs.data(), valid.data()) != ZKP_OK) return EngineFailure;
    for (uint8_t v : valid)
      if (!v) return VerificationFailure;
  }
  Scalar minus_c = sc_neg(proof.challenge);
  std::vector<uint8_t> sc, pts;
  std::vector<uint64_t> offsets(1, 0);
  for (auto& c : constraints_) {
    uint8_t b[32];
    for (auto& term : c.second) {
      proof.responses[term.first].to_bytes(b);
      sc.insert(sc.end(), b, b + 32);
      pts.insert(pts.end(), points_[term.second].begin(), points_[term.second].end());
    }
    minus_c.to_bytes(b);
    sc.insert(sc.end(), b, b + 32);
    pts.insert(pts.end(), points_[c.first].begin(), points_[c.first].end());
    offsets.push_back(sc.size() / 32);
  }
  size_t k = constraints_.size();
  std::vector<uint8_t> out(k * 32 + 32), valid(k + 1);
  if (k && zkp_msm_vartime_batched(ctx_, sc.data(), pts.data(), offsets.data(), k, out.data(), valid.data()) != ZKP_OK)
    return EngineFailure;
  for (size_t i = 0; i < k; i++) {
    if (!valid[i]) return VerificationFailure;
    Enc e;
    memcpy(e.data(), out.data() + 32 * i, 32);
    append_blinding_commitment(*transcript_, point_labels_[constraints_[i].first], e);
  }
  Scalar challenge = get_challenge(*transcript_, "chal");
  return challenge == proof.challenge ? PROOF_OK : VerificationFailure;
}

ProofError Verifier::verify_batchable(const BatchableProof& proof, Rng& rng) {
  if ((int)proof.responses.size() != num_scalars_) return VerificationFailure;
  if (proof.commitments.size() != constraints_.size()) return VerificationFailure;
  for (size_t i = 0; i < proof.commitments.size(); i++)
    if (!validate_and_append_blinding_commitment(*transcript_, point_labels_[constraints_[i].first], proof.commitments[i]))
      return VerificationFailure;
  Scalar minus_c = sc_neg(get_challenge(*transcript_, "chal"));
  size_t off = points_.size(), n = off + proof.commitments.size();
  std::vector<Scalar> coeffs(n, Scalar::zero());
  for (size_t i = 0; i < constraints_.size(); i++) {
    Scalar rho = rng.u128();
    coeffs[off + i] = sc_sub(coeffs[off + i], rho);
    coeffs[constraints_[i].first] = sc_add(coeffs[constraints_[i].first], sc_mul(rho, minus_c));
    for (auto& term : constraints_[i].second)
      coeffs[term.second] = sc_add(coeffs[term.second], sc_mul(rho, proof.responses[term.first]));
  }
  std::vector<uint8_t> sc(n * 32), pts(n * 32);
  for (size_t i = 0; i < n; i++) {
    coeffs[i].to_bytes(&sc[32 * i]);
    memcpy(&pts[32 * i], i < off ? points_[i].data() : proof.commitments[i - off].data(), 32);
  }
  uint8_t out[32];
  int32_t ident = 0;
  int64_t bad = -1;
  int32_t rc = zkp_msm_vartime(ctx_, sc.data(), pts.data(), n, out, &ident, &bad);
  if (rc == ZKP_ERR_POINT) return VerificationFailure;
  if (rc != ZKP_OK) return EngineFailure;
  return ident ? PROOF_OK : VerificationFailure;
}

// ---- BatchVerifier -----------------------------------------------------------------------------------------
BatchVerifier::BatchVerifier(zkp_ctx* ctx, const std::string& proof_label, size_t batch_size,
                             std::vector<Transcript>* transcripts, ProofError* err, int threads,
                             bool identical_transcripts)
    : ctx_(ctx), batch_size_(batch_size), transcripts_(transcripts), num_scalars_(0), threads_(threads),
      uniform_(identical_transcripts) {
  if (transcripts->size() != batch_size) {
    *err = BatchSizeMismatch;
    return;
  }
  script_.push_back(Op{0, proof_label, 0});
  *err = PROOF_OK;
}
int BatchVerifier::allocate_scalar(const std::string& label) {
  script_.push_back(Op{1, label, 0});
  return num_scalars_++;
}
BatchPointVar BatchVerifier::allocate_static_point(const std::string& label, const Enc& assignment, ProofError* err) {
  BatchPointVar v = {true, -1};
  if (is_identity_encoding(assignment) && batch_size_) {   // toolbox/mod.rs:191: the first transcript already fails
    *err = VerificationFailure;
    return v;
  }
  static_points_.push_back(assignment);
  static_point_labels_.push_back(label);
  v.idx = (int)static_points_.size() - 1;
  script_.push_back(Op{2, label, v.idx});
  *err = PROOF_OK;
  return v;
}
BatchPointVar BatchVerifier::allocate_instance_point(const std::string& label, const std::vector<Enc>& assignments,
                                                     ProofError* err) {
  BatchPointVar v = {false, -1};
  if (assignments.size() != batch_size_) {
    *err = BatchSizeMismatch;
    return v;
  }
  for (size_t j = 0; j < batch_size_; j++)
    if (is_identity_encoding(assignments[j])) {
      *err = VerificationFailure;
      return v;
    }
  instance_points_.push_back(assignments);
  instance_point_labels_.push_back(label);
  v.idx = (int)instance_points_.size() - 1;
  script_.push_back(Op{3, label, v.idx});
  *err = PROOF_OK;
  return v;
}
void BatchVerifier::replay(Transcript& t, size_t j, size_t from, size_t to) const {
  for (size_t o = from; o < to; o++) {
    const Op& op = script_[o];
    switch (op.kind) {
      case 0: domain_sep(t, op.label); break;
      case 1: append_scalar_var(t, op.label); break;
      case 2: append_point_var(t, op.label, static_points_[op.idx]); break;
      default: append_point_var(t, op.label, instance_points_[op.idx][j]); break;
    }
  }
}

ProofError BatchVerifier::batch_coeffs(const std::vector<BatchableProof>& proofs, Rng& rng, int threads,
                                       std::vector<uint8_t>* static_out, std::vector<uint8_t>* inst_out,
                                       std::vector<uint8_t>* pts_out) {
  const size_t N = batch_size_;
  if (proofs.size() != N) return BatchSizeMismatch;
  for (auto& p : proofs)
    if (p.commitments.size() != constraints_.size() || (int)p.responses.size() != num_scalars_) return VerificationFailure;
  const size_t num_s = static_points_.size(), num_i = instance_points_.size(), num_c = constraints_.size();
  // replay the recorded allocations (hashing the batch-wide prefix once when all transcripts started identical),
  // then the commitments and the challenges (batch_verifier.rs:75-167) -- independent per proof
  size_t prefix = 0;
  Transcript prefix_state;
  if (uniform_ && N) {
    while (prefix < script_.size() && script_[prefix].kind != 3) prefix++;
    prefix_state = (*transcripts_)[0];
    replay(prefix_state, 0, 0, prefix);
  }
  std::vector<Scalar> minus_c(N);
  std::vector<uint8_t> bad(N, 0);
  parallel_for(N, threads, [&](size_t lo, size_t hi, int) {
    for (size_t j = lo; j < hi; j++) {
      if (prefix) (*transcripts_)[j] = prefix_state;
      replay((*transcripts_)[j], j, prefix, script_.size());
      for (size_t i = 0; i < num_c; i++) {
        const BatchPointVar& lhs = constraints_[i].first;
        const std::string& label = lhs.is_static ? static_point_labels_[lhs.idx] : instance_point_labels_[lhs.idx];
        if (!validate_and_append_blinding_commitment((*transcripts_)[j], label, proofs[j].commitments[i])) bad[j] = 1;
      }
      minus_c[j] = sc_neg(get_challenge((*transcripts_)[j], "chal"));
    }
  });
  for (uint8_t b : bad)
    if (b) return VerificationFailure;
  // the random weights in the reference's draw order: for i in constraints, for j in batch (batch_verifier.rs:176-179)
  std::vector<Scalar> rho(num_c * N);
  for (size_t i = 0; i < num_c * N; i++) rho[i] = rng.u128();
  const size_t rows = num_i + num_c;
  std::vector<Scalar> inst(rows * N, Scalar::zero());
  int nthreads = threads <= 0 ? (int)std::thread::hardware_concurrency() : threads;
  if (nthreads < 1) nthreads = 1;
  std::vector<std::vector<Scalar>> static_part(nthreads, std::vector<Scalar>(num_s, Scalar::zero()));
  parallel_for(N, nthreads, [&](size_t lo, size_t hi, int tid) {
    std::vector<Scalar>& sp = static_part[tid];
    for (size_t j = lo; j < hi; j++) {
      for (size_t i = 0; i < num_c; i++) {
        const Scalar& r = rho[i * N + j];
        Scalar& com = inst[(num_i + i) * N + j];
        com = sc_sub(com, r);
        const BatchPointVar& lhs = constraints_[i].first;
        Scalar rc = sc_mul(r, minus_c[j]);
        if (lhs.is_static) sp[lhs.idx] = sc_add(sp[lhs.idx], rc);
        else inst[lhs.idx * N + j] = sc_add(inst[lhs.idx * N + j], rc);
        for (auto& term : constraints_[i].second) {
          Scalar rr = sc_mul(r, proofs[j].responses[term.first]);
          if (term.second.is_static) sp[term.second.idx] = sc_add(sp[term.second.idx], rr);
          else inst[term.second.idx * N + j] = sc_add(inst[term.second.idx * N + j], rr);
        }
      }
    }
  });
  static_out->assign(num_s * 32, 0);
  for (size_t s = 0; s < num_s; s++) {
    Scalar tot = Scalar::zero();
    for (auto& part : static_part) tot = sc_add(tot, part[s]);
    tot.to_bytes(&(*static_out)[32 * s]);
  }
  inst_out->resize(rows * N * 32);
  pts_out->resize(rows * N * 32);
  parallel_for(N, nthreads, [&](size_t lo, size_t hi, int) {
    for (size_t r = 0; r < rows; r++)
      for (size_t j = lo; j < hi; j++) {
        inst[r * N + j].to_bytes(&(*inst_out)[(r * N + j) * 32]);
        const Enc& e = r < num_i ? instance_points_[r][j] : proofs[j].commitments[r - num_i];
        memcpy(&(*pts_out)[(r * N + j) * 32], e.data(), 32);
      }
  });
  return PROOF_OK;
}

ProofError BatchVerifier::verify_batchable(const std::vector<BatchableProof>& proofs, Rng& rng, int threads) {
  std::vector<uint8_t> sc, ic, ip;
  ProofError e = batch_coeffs(proofs, rng, threads, &sc, &ic, &ip);
  if (e != PROOF_OK) return e;
  std::vector<uint8_t> sp(static_points_.size() * 32);
  for (size_t s = 0; s < static_points_.size(); s++) memcpy(&sp[32 * s], static_points_[s].data(), 32);
  int32_t accept = 0;
  int64_t bad = -1;
  int32_t rc = zkp_batch_verify(ctx_, sc.data(), sp.data(), static_points_.size(), ic.data(), ip.data(), rows(),
                                batch_size_, &accept, &bad);
  if (rc == ZKP_ERR_POINT) return VerificationFailure;
  if (rc != ZKP_OK) return EngineFailure;
  return accept ? PROOF_OK : VerificationFailure;
}

// ---- define_proof! mirror ----------------------------------------------------------------------------------
ProofError stmt_prove(zkp_ctx* ctx, const Statement& st, Transcript* t, const std::vector<Scalar>& secrets,
                      const std::vector<Limbs>& points, Rng& rng, CompactProof* compact, BatchableProof* batchable,
                      std::vector<Enc>* encodings) {
  if (secrets.size() != st.secrets.size() || points.size() != st.num_points()) return BatchSizeMismatch;
  Prover pr(ctx, st.label, t);
  for (size_t i = 0; i < secrets.size(); i++) pr.allocate_scalar(st.secrets[i], secrets[i]);
  if (encodings) encodings->clear();
  for (size_t i = 0; i < points.size(); i++) {
    Enc e;
    ProofError err;
    pr.allocate_point(st.point_name(i), points[i], &e, &err);
    if (err != PROOF_OK) return err;
    if (encodings) encodings->push_back(e);
  }
  for (auto& c : st.constraints) pr.constrain(c.first, c.second);
  if (compact) return pr.prove_compact(rng, compact);
  return pr.prove_batchable(rng, batchable);
}

static ProofError build_verifier(Verifier& v, const Statement& st, const std::vector<Enc>& points) {
  if (points.size() != st.num_points()) return BatchSizeMismatch;
  for (auto& s : st.secrets) v.allocate_scalar(s);
  for (size_t i = 0; i < points.size(); i++) {
    ProofError err;
    v.allocate_point(st.point_name(i), points[i], &err);
    if (err != PROOF_OK) return err;
  }
  for (auto& c : st.constraints) v.constrain(c.first, c.second);
  return PROOF_OK;
}
ProofError stmt_verify_compact(zkp_ctx* ctx, const Statement& st, Transcript* t, const std::vector<Enc>& points,
                               const CompactProof& proof) {
  Verifier v(ctx, st.label, t);
  ProofError e = build_verifier(v, st, points);
  return e != PROOF_OK ? e : v.verify_compact(proof);
}
ProofError stmt_verify_batchable(zkp_ctx* ctx, const Statement& st, Transcript* t, const std::vector<Enc>& points,
                                 const BatchableProof& proof, Rng& rng) {
  Verifier v(ctx, st.label, t);
  ProofError e = build_verifier(v, st, points);
  return e != PROOF_OK ? e : v.verify_batchable(proof, rng);
}
ProofError stmt_batch_verify(zkp_ctx* ctx, const Statement& st, std::vector<Transcript>* transcripts,
                             const std::vector<std::vector<Enc>>& instance, const std::vector<Enc>& common,
                             const std::vector<BatchableProof>& proofs, Rng& rng, int threads) {
  ProofError err;
  BatchVerifier bv(ctx, st.label, proofs.size(), transcripts, &err, threads);
  if (err != PROOF_OK) return err;
  if (instance.size() != st.instance.size() || common.size() != st.common.size()) return BatchSizeMismatch;
  for (auto& s : st.secrets) bv.allocate_scalar(s);
  std::vector<BatchPointVar> pv;
  for (size_t i = 0; i < instance.size(); i++) {
    pv.push_back(bv.allocate_instance_point(st.instance[i], instance[i], &err));
    if (err != PROOF_OK) return err;
  }
  for (size_t i = 0; i < common.size(); i++) {
    pv.push_back(bv.allocate_static_point(st.common[i], common[i], &err));
    if (err != PROOF_OK) return err;
  }
  for (auto& c : st.constraints) {
    BatchLinComb lc;
    for (auto& term : c.second) lc.push_back(std::make_pair(term.first, pv[term.second]));
    bv.constrain(pv[c.first], lc);
  }
  return bv.verify_batchable(proofs, rng, threads);
}

ProofError stmt_prove_many(zkp_ctx* ctx, const Statement& st, const std::string& transcript_label, size_t N,
                           const Scalar* secrets, const Limbs* points, const uint8_t* entropy, int threads,
                           std::vector<BatchableProof>* proofs, std::vector<Enc>* encodings) {
  const size_t m = st.secrets.size(), p = st.num_points(), k = st.constraints.size();
  // (1) every allocate_point compression of every proof in one device call (toolbox/mod.rs:180)
  encodings->resize(N * p);
  if (N && p && zkp_compress_batch(ctx, points[0].data(), N * p, (*encodings)[0].data()) != ZKP_OK) return EngineFailure;
  // (2) per proof: transcript up to the commitments, synthetic-nonce blindings (prover.rs:78-89)
  size_t terms = 0;
  for (auto& c : st.constraints) terms += c.second.size();
  std::vector<Transcript> tr(N);
  std::vector<Scalar> blind(N * m);
  std::vector<uint8_t> sc(N * terms * 32);
  std::vector<uint64_t> pts(N * terms * 20), offsets(N * k + 1);
  parallel_for(N, threads, [&](size_t lo, size_t hi, int) {
    for (size_t j = lo; j < hi; j++) {
      Transcript t(U8(transcript_label), transcript_label.size());
      domain_sep(t, st.label);
      for (size_t i = 0; i < m; i++) append_scalar_var(t, st.secrets[i]);
      for (size_t i = 0; i < p; i++) append_point_var(t, st.point_name(i), (*encodings)[j * p + i]);
      std::vector<Scalar> sec(secrets + j * m, secrets + (j + 1) * m), bl;
      draw_blindings(t, sec, entropy + 32 * j, &bl);
      size_t pos = j * terms;
      for (size_t ci = 0; ci < k; ci++) {
        offsets[j * k + ci] = pos;
        for (auto& term : st.constraints[ci].second) {
          bl[term.first].to_bytes(&sc[pos * 32]);
          memcpy(&pts[pos * 20], points[j * p + term.second].data(), 160);
          pos++;
        }
      }
      for (size_t i = 0; i < m; i++) blind[j * m + i] = bl[i];
      tr[j] = t;
    }
  });
  offsets[N * k] = N * terms;
  // (3) all N*k constant-time MSMs + compressions in one device call (prover.rs:93-103)
  std::vector<uint8_t> coms(N * k * 32 + 32);
  if (N && k && zkp_msm_ct_batched(ctx, sc.data(), pts.data(), ZKP_POINTS_LIMBS51, offsets.data(), N * k, coms.data()) != ZKP_OK)
    return EngineFailure;
  // (4) per proof: commitments into the transcript, challenge, responses (prover.rs:98-109)
  proofs->resize(N);
  parallel_for(N, threads, [&](size_t lo, size_t hi, int) {
    for (size_t j = lo; j < hi; j++) {
      BatchableProof& pr = (*proofs)[j];
      pr.commitments.resize(k);
      for (size_t ci = 0; ci < k; ci++) {
        memcpy(pr.commitments[ci].data(), &coms[(j * k + ci) * 32], 32);
        append_blinding_commitment(tr[j], st.point_name(st.constraints[ci].first), pr.commitments[ci]);
      }
      Scalar c = get_challenge(tr[j], "chal");
      pr.responses.resize(m);
      for (size_t i = 0; i < m; i++) pr.responses[i] = sc_muladd(secrets[j * m + i], c, blind[j * m + i]);
    }
  });
  return PROOF_OK;
}

}  // namespace zkp_host
