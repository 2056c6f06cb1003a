// C++ port of /root/reference/tests/dleq_using_constraint_api.rs against the host mirror of zkp's toolbox
// (zkp_b200/csrc/host/toolbox.hpp) and the CUDA engine behind it: the same generic statement over the constraint-system
// interface (SchnorrCS, toolbox/mod.rs:86-98), the same labels, allocation order and secrets, the same three tests
// (:41-78 compact, :81-118 batchable, :121-170 batch of 16), plus the rejections the reference's other test file checks
// (tests/sig_and_vrf_example.rs:200-221).  Group elements come from the engine: A = x*B and G = x*H through the
// constant-time MSM entry point, limb forms through zkp_decompress_batch.
// Build + run: tests/test_cpp_port.py (g++ -std=c++17, links zkp_b200/lib/libzkp_b200.so).  Exit code 0 = all passed.
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/zkp_b200.h"
#include "../../zkp_b200/csrc/host/toolbox.hpp"

using namespace zkp_host;

static zkp_ctx* ctx;
static int failures = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) { printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
  } while (0)

static Enc hex(const char* h) {
  Enc e;
  for (int i = 0; i < 32; i++) { unsigned v; sscanf(h + 2 * i, "%2x", &v); e[i] = (uint8_t)v; }
  return e;
}
// dalek_constants::RISTRETTO_BASEPOINT_POINT.compress() (RFC 9496 A.1) and
// RistrettoPoint::hash_from_bytes::<Sha512>(B.compress().as_bytes()) (value pinned by oracle/ristretto.py + libsodium)
static const Enc B_ENC = hex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76");
static const Enc H_ENC = hex("90ca11cd6c6227cb0abc39e2710c444ae6617ea81898e716353f3410d9656605");

static Limbs decompress(const Enc& e) {
  Limbs l;
  uint8_t valid = 0;
  CHECK(zkp_decompress_batch(ctx, e.data(), 1, l.data(), &valid) == ZKP_OK && valid == 1);
  return l;
}
// P * x  (RistrettoPoint * Scalar)
static Limbs mul(const Limbs& P, const Scalar& x) {
  uint8_t sb[32];
  x.to_bytes(sb);
  uint64_t off[2] = {0, 1};
  Enc out;
  CHECK(zkp_msm_ct_batched(ctx, sb, P.data(), ZKP_POINTS_LIMBS51, off, 1, out.data()) == ZKP_OK);
  return decompress(out);
}
static Transcript transcript(const char* label) { return Transcript((const uint8_t*)label, strlen(label)); }

// fn dleq_statement<CS: SchnorrCS>(cs, x, A, G, B, H)        (dleq_using_constraint_api.rs:28-38)
template <class CS>
static void dleq_statement(CS& cs, typename CS::ScalarVar x, typename CS::PointVar A, typename CS::PointVar G,
                           typename CS::PointVar B, typename CS::PointVar H) {
  cs.constrain(A, typename CS::LC{{x, B}});
  cs.constrain(G, typename CS::LC{{x, H}});
}

struct Proved { CompactProof compact; BatchableProof batchable; Enc cmpr_A, cmpr_G; };

static Proved prove(const char* tlabel, uint64_t xval, bool batchable, const char* seed) {
  const Limbs B = decompress(B_ENC), H = decompress(H_ENC);
  const Scalar x = Scalar::from_u128(xval, 0);
  const Limbs A = mul(B, x), G = mul(H, x);
  Transcript t = transcript(tlabel);
  Prover prover(ctx, "DLEQProof", &t);
  ProofError err = PROOF_OK;
  Proved out;
  Enc ignore;
  // committing var names to the transcript forces the ordering
  int var_x = prover.allocate_scalar("x", x);
  int var_B = prover.allocate_point("B", B, &ignore, &err);
  int var_H = prover.allocate_point("H", H, &ignore, &err);
  int var_A = prover.allocate_point("A", A, &out.cmpr_A, &err);
  int var_G = prover.allocate_point("G", G, &out.cmpr_G, &err);
  CHECK(err == PROOF_OK && ignore == H_ENC);
  dleq_statement(prover, var_x, var_A, var_G, var_B, var_H);
  Rng rng((const uint8_t*)seed, strlen(seed));
  if (batchable) CHECK(prover.prove_batchable(rng, &out.batchable) == PROOF_OK);
  else CHECK(prover.prove_compact(rng, &out.compact) == PROOF_OK);
  return out;
}

static ProofError verify(const char* tlabel, const char* proof_label, const Proved& p, bool batchable, const Enc& A_enc) {
  Transcript t = transcript(tlabel);
  Verifier verifier(ctx, proof_label, &t);
  ProofError err = PROOF_OK;
  int var_x = verifier.allocate_scalar("x");
  int var_B = verifier.allocate_point("B", B_ENC, &err);
  int var_H = verifier.allocate_point("H", H_ENC, &err);
  int var_A = verifier.allocate_point("A", A_enc, &err);
  int var_G = verifier.allocate_point("G", p.cmpr_G, &err);
  if (err != PROOF_OK) return err;
  dleq_statement(verifier, var_x, var_A, var_G, var_B, var_H);
  Rng rng((const uint8_t*)"verifier-rng", 12);
  return batchable ? verifier.verify_batchable(p.batchable, rng) : verifier.verify_compact(p.compact);
}

static void create_and_verify_compact_dleq() {
  Proved p = prove("DLEQTest", 89327492234ull, false, "rng-1");
  CHECK(p.compact.responses.size() == 1);
  CHECK(verify("DLEQTest", "DLEQProof", p, false, p.cmpr_A) == PROOF_OK);
  // wrong public point, wrong domain separator, wrong transcript label, tampered response -> Err
  CHECK(verify("DLEQTest", "DLEQProof", p, false, p.cmpr_G) == VerificationFailure);
  CHECK(verify("DLEQTest", "DLEQProoG", p, false, p.cmpr_A) == VerificationFailure);
  CHECK(verify("DLEQTesT", "DLEQProof", p, false, p.cmpr_A) == VerificationFailure);
  Proved q = p;
  q.compact.responses[0] = sc_add(q.compact.responses[0], Scalar::from_u128(1, 0));
  CHECK(verify("DLEQTest", "DLEQProof", q, false, p.cmpr_A) == VerificationFailure);
  q = p;
  q.compact.responses.push_back(Scalar::zero());   // verifier.rs:82-84
  CHECK(verify("DLEQTest", "DLEQProof", q, false, p.cmpr_A) == VerificationFailure);
}

static void create_and_verify_batchable_dleq() {
  Proved p = prove("DLEQTest", 89327492234ull, true, "rng-2");
  CHECK(p.batchable.commitments.size() == 2 && p.batchable.responses.size() == 1);
  CHECK(verify("DLEQTest", "DLEQProof", p, true, p.cmpr_A) == PROOF_OK);
  CHECK(verify("DLEQTest", "DLEQProof", p, true, p.cmpr_G) == VerificationFailure);
  CHECK(verify("DLEQTest", "DLEQProoG", p, true, p.cmpr_A) == VerificationFailure);
  Proved q = p;
  q.batchable.commitments[1] = q.batchable.commitments[0];
  CHECK(verify("DLEQTest", "DLEQProof", q, true, p.cmpr_A) == VerificationFailure);
  q = p;
  memset(q.batchable.commitments[0].data(), 0, 32);   // identity encoding (toolbox/mod.rs:215)
  CHECK(verify("DLEQTest", "DLEQProof", q, true, p.cmpr_A) == VerificationFailure);
}

static void create_and_batch_verify_batchable_dleq() {
  const size_t batch_size = 16;
  std::vector<BatchableProof> proofs;
  std::vector<Enc> cmpr_As, cmpr_Gs;
  for (size_t j = 0; j < batch_size; j++) {
    std::string seed = "rng-batch-" + std::to_string(j);
    Proved p = prove("DLEQBatchTest", (uint64_t)j + 89327492234ull, true, seed.c_str());
    proofs.push_back(p.batchable);
    cmpr_As.push_back(p.cmpr_A);
    cmpr_Gs.push_back(p.cmpr_G);
  }
  auto run = [&](const std::vector<BatchableProof>& pr, const std::vector<Enc>& As, size_t n_transcripts) {
    std::vector<Transcript> transcripts(n_transcripts, transcript("DLEQBatchTest"));
    ProofError err = PROOF_OK;
    BatchVerifier verifier(ctx, "DLEQProof", batch_size, &transcripts, &err);
    if (err != PROOF_OK) return err;
    int var_x = verifier.allocate_scalar("x");
    BatchPointVar var_B = verifier.allocate_static_point("B", B_ENC, &err);
    BatchPointVar var_H = verifier.allocate_static_point("H", H_ENC, &err);
    BatchPointVar var_A = verifier.allocate_instance_point("A", As, &err);
    if (err != PROOF_OK) return err;
    BatchPointVar var_G = verifier.allocate_instance_point("G", cmpr_Gs, &err);
    if (err != PROOF_OK) return err;
    dleq_statement(verifier, var_x, var_A, var_G, var_B, var_H);
    Rng rng((const uint8_t*)"batch-rng", 9);
    return verifier.verify_batchable(pr, rng);
  };
  CHECK(run(proofs, cmpr_As, batch_size) == PROOF_OK);
  // one bad proof in the batch rejects the batch; size errors are BatchSizeMismatch (batch_verifier.rs:72-74, :120-122, :138-148)
  std::vector<BatchableProof> bad = proofs;
  bad[7].responses[0] = sc_add(bad[7].responses[0], Scalar::from_u128(1, 0));
  CHECK(run(bad, cmpr_As, batch_size) == VerificationFailure);
  std::vector<Enc> swapped = cmpr_As;
  std::swap(swapped[2], swapped[3]);
  CHECK(run(proofs, swapped, batch_size) == VerificationFailure);
  CHECK(run(proofs, cmpr_As, batch_size - 1) == BatchSizeMismatch);
  std::vector<Enc> fewer(cmpr_As.begin(), cmpr_As.end() - 1);
  CHECK(run(proofs, fewer, batch_size) == BatchSizeMismatch);
  std::vector<BatchableProof> fewer_proofs(proofs.begin(), proofs.end() - 1);
  CHECK(run(fewer_proofs, cmpr_As, batch_size) == BatchSizeMismatch);
}

int main() {
  if (zkp_ctx_create(&ctx, 0) != ZKP_OK) {
    printf("no CUDA device: this test needs the engine\n");
    return 2;
  }
  create_and_verify_compact_dleq();
  create_and_verify_batchable_dleq();
  create_and_batch_verify_batchable_dleq();
  zkp_ctx_destroy(ctx);
  printf(failures ? "%d check(s) FAILED\n" : "all reference tests passed (%d failures)\n", failures);
  return failures ? 1 : 0;
}
