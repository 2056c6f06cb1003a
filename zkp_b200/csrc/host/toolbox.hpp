// Host-side mirror of the reference's constraint-system API, routed through the C ABI of the engine.
//
// Same names, argument meaning and error behaviour as
//   toolbox::SchnorrCS / TranscriptProtocol   /root/reference/src/toolbox/mod.rs:86-98, :102-228
//   toolbox::prover::Prover                   /root/reference/src/toolbox/prover.rs:23-142
//   toolbox::verifier::Verifier               /root/reference/src/toolbox/verifier.rs:26-183
//   toolbox::batch_verifier::BatchVerifier    /root/reference/src/toolbox/batch_verifier.rs:30-245
//   define_proof! generated module functions  /root/reference/src/macros.rs:206-370   (struct Statement below)
// Every multiscalar multiplication, compression and decompression goes through include/zkp_b200.h (CUDA);
// Merlin hashing, scalar arithmetic mod l and the coefficient fold stay on the host, as in the reference.
// The reference is Rust; with no Rust toolchain in this image the host side is C++ (INTEGRATION.md shows the
// Rust shim).  Randomness the reference takes from thread_rng is injected (Rng) so runs are reproducible.
#pragma once
#include <array>
#include <functional>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/zkp_b200.h"
#include "merlin.hpp"
#include "scalar.hpp"

namespace zkp_host {

enum ProofError { PROOF_OK = 0, VerificationFailure = 1, BatchSizeMismatch = 2, EngineFailure = 3 };

typedef std::array<uint8_t, 32> Enc;            // CompressedRistretto
typedef std::array<uint64_t, 20> Limbs;         // RistrettoPoint in FieldElement51 limb form (X,Y,Z,T)

struct CompactProof { Scalar challenge; std::vector<Scalar> responses; };              // proofs.rs:15-20
struct BatchableProof { std::vector<Enc> commitments; std::vector<Scalar> responses; };  // proofs.rs:27-32

// Deterministic stand-in for rand::thread_rng: SHAKE-256 of a seed (same stream as oracle.toolbox.SeededRng).
class Rng {
 public:
  Rng(const uint8_t* seed, size_t len);
  void bytes(uint8_t* out, size_t n);
  Scalar u128();

 private:
  uint64_t st_[25];
  size_t pos_;
};

// ---- TranscriptProtocol (toolbox/mod.rs:165-228) -----------------------------------------------------------
void domain_sep(Transcript& t, const std::string& label);
void append_scalar_var(Transcript& t, const std::string& label);
void append_point_var(Transcript& t, const std::string& label, const Enc& enc);       // after compress
bool validate_and_append_point_var(Transcript& t, const std::string& label, const Enc& enc);
void append_blinding_commitment(Transcript& t, const std::string& label, const Enc& enc);
bool validate_and_append_blinding_commitment(Transcript& t, const std::string& label, const Enc& enc);
Scalar get_challenge(Transcript& t, const std::string& label);

typedef std::vector<std::pair<int, int>> LinComb;  // (scalar var, point var)

// run fn(lo, hi, thread_index) over [0, n) on `threads` host threads (0 = all cores)
void parallel_for(size_t n, int threads, const std::function<void(size_t, size_t, int)>& fn);

class Prover {
 public:
  // the associated types of the SchnorrCS trait (toolbox/mod.rs:86-98), so generic statements can be written once
  typedef int ScalarVar;
  typedef int PointVar;
  typedef LinComb LC;
  Prover(zkp_ctx* ctx, const std::string& proof_label, Transcript* transcript);
  int allocate_scalar(const std::string& label, const Scalar& assignment);
  // compresses on the device (toolbox/mod.rs:180) and returns the encoding like the reference does
  int allocate_point(const std::string& label, const Limbs& assignment, Enc* enc_out, ProofError* err);
  void constrain(int lhs, const LinComb& lc) { constraints_.push_back(std::make_pair(lhs, lc)); }
  ProofError prove_compact(Rng& rng, CompactProof* out);
  ProofError prove_batchable(Rng& rng, BatchableProof* out);
  // exposed for parity tests: the blindings drawn by the last prove_* call
  std::vector<Scalar> last_blindings;

 private:
  ProofError prove_impl(Rng& rng, Scalar* challenge, std::vector<Scalar>* responses, std::vector<Enc>* commitments);
  zkp_ctx* ctx_;
  Transcript* transcript_;
  std::vector<Scalar> scalars_;
  std::vector<Limbs> points_;
  std::vector<std::string> point_labels_;
  std::vector<std::pair<int, LinComb>> constraints_;
};

class Verifier {
 public:
  typedef int ScalarVar;
  typedef int PointVar;
  typedef LinComb LC;
  Verifier(zkp_ctx* ctx, const std::string& proof_label, Transcript* transcript);
  int allocate_scalar(const std::string& label);
  int allocate_point(const std::string& label, const Enc& assignment, ProofError* err);
  void constrain(int lhs, const LinComb& lc) { constraints_.push_back(std::make_pair(lhs, lc)); }
  ProofError verify_compact(const CompactProof& proof);
  ProofError verify_batchable(const BatchableProof& proof, Rng& rng);

 private:
  zkp_ctx* ctx_;
  Transcript* transcript_;
  int num_scalars_;
  std::vector<Enc> points_;
  std::vector<std::string> point_labels_;
  std::vector<std::pair<int, LinComb>> constraints_;
};

struct BatchPointVar { bool is_static; int idx; };
typedef std::vector<std::pair<int, BatchPointVar>> BatchLinComb;

class BatchVerifier {
 public:
  typedef int ScalarVar;
  typedef BatchPointVar PointVar;
  typedef BatchLinComb LC;
  // transcripts.size() must equal batch_size, else *err = BatchSizeMismatch (batch_verifier.rs:72-74)
  // `threads` host threads work on the per-proof transcripts (0 = all cores); the reference is single-threaded
  BatchVerifier(zkp_ctx* ctx, const std::string& proof_label, size_t batch_size, std::vector<Transcript>* transcripts,
                ProofError* err, int threads = 1, bool identical_transcripts = false);
  int allocate_scalar(const std::string& label);
  BatchPointVar allocate_static_point(const std::string& label, const Enc& assignment, ProofError* err);
  BatchPointVar allocate_instance_point(const std::string& label, const std::vector<Enc>& assignments, ProofError* err);
  void constrain(BatchPointVar lhs, const BatchLinComb& lc) { constraints_.push_back(std::make_pair(lhs, lc)); }
  ProofError verify_batchable(const std::vector<BatchableProof>& proofs, Rng& rng, int threads = 0);
  // everything before the MSM (batch_verifier.rs:138-217), exposed for parity tests
  ProofError batch_coeffs(const std::vector<BatchableProof>& proofs, Rng& rng, int threads,
                          std::vector<uint8_t>* static_coeffs, std::vector<uint8_t>* inst_coeffs,
                          std::vector<uint8_t>* inst_points);
  size_t num_static() const { return static_points_.size(); }
  size_t rows() const { return instance_points_.size() + constraints_.size(); }
  const std::vector<Enc>& static_points() const { return static_points_; }

 private:
  zkp_ctx* ctx_;
  size_t batch_size_;
  std::vector<Transcript>* transcripts_;
  int num_scalars_;
  int threads_;
  // true while every transcript is known to hold the same state (caller passed identical copies and only
  // batch-wide data has been absorbed): such appends are hashed once and the state is copied
  bool uniform_;
  // allocation calls are recorded and replayed per transcript inside ONE parallel region at verification time
  // (the reference hashes eagerly, transcript by transcript; the transcripts' final states are identical)
  struct Op { int kind; std::string label; int idx; };   // kind: 0 dom-sep, 1 scalar, 2 static point, 3 instance point
  std::vector<Op> script_;
  void replay(Transcript& t, size_t j, size_t from, size_t to) const;
  std::vector<Enc> static_points_;
  std::vector<std::string> static_point_labels_;
  std::vector<std::vector<Enc>> instance_points_;
  std::vector<std::string> instance_point_labels_;
  std::vector<std::pair<BatchPointVar, BatchLinComb>> constraints_;
};

// ---- define_proof! mirror (macros.rs:74-370) ---------------------------------------------------------------
struct Statement {
  std::string name, label;
  std::vector<std::string> secrets, instance, common;
  // (lhs point name index, [(secret index, point index)]) with point indices over instance ++ common
  std::vector<std::pair<int, LinComb>> constraints;
  size_t num_points() const { return instance.size() + common.size(); }
  const std::string& point_name(size_t i) const { return i < instance.size() ? instance[i] : common[i - instance.size()]; }
};

// module::prove_compact / prove_batchable (macros.rs:261-278): points in allocation order instance ++ common
ProofError stmt_prove(zkp_ctx* ctx, const Statement& st, Transcript* t, const std::vector<Scalar>& secrets,
                      const std::vector<Limbs>& points, Rng& rng, CompactProof* compact, BatchableProof* batchable,
                      std::vector<Enc>* encodings);
// module::verify_compact / verify_batchable (macros.rs:314-333)
ProofError stmt_verify_compact(zkp_ctx* ctx, const Statement& st, Transcript* t, const std::vector<Enc>& points,
                               const CompactProof& proof);
ProofError stmt_verify_batchable(zkp_ctx* ctx, const Statement& st, Transcript* t, const std::vector<Enc>& points,
                                 const BatchableProof& proof, Rng& rng);
// module::batch_verify (macros.rs:336-370): instance[i] has one encoding per proof, common[i] one encoding
ProofError stmt_batch_verify(zkp_ctx* ctx, const Statement& st, std::vector<Transcript>* transcripts,
                             const std::vector<std::vector<Enc>>& instance, const std::vector<Enc>& common,
                             const std::vector<BatchableProof>& proofs, Rng& rng, int threads);

// N independent proofs of one statement in three batched device calls (the reference has no batch prover;
// this is N x module::prove_batchable with identical per-proof results): BASELINE configs[1].
// secrets[N][m], points[N][p] (limb form, allocation order), entropy[N][32]; transcript_label starts every transcript.
ProofError stmt_prove_many(zkp_ctx* ctx, const Statement& st, const std::string& transcript_label, size_t N,
                           const Scalar* secrets, const Limbs* points, const uint8_t* entropy, int threads,
                           std::vector<BatchableProof>* proofs, std::vector<Enc>* encodings /* N x p */);

}  // namespace zkp_host
