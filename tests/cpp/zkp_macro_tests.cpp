// C++ port of /root/reference/tests/zkp.rs against the define_proof! mirror of the host side
// (zkp_b200/csrc/host/toolbox.hpp `Statement` + stmt_prove / stmt_verify_* / stmt_batch_verify = the functions the macro
// generates, /root/reference/src/macros.rs:261-370) and the wire format of src/proofs.rs (host/wire.cpp):
//   define_proof! {dleq, "DLEQ Example Proof", (x), (A, B, H), (G) : A = (x * G), B = (x * H) }      (tests/zkp.rs:28)
//   create_and_verify_compact (:31-69), create_and_verify_batchable (:72-112), create_batch_and_batch_verify (:115-175)
// with the same inputs (H = hash_from_bytes::<Sha512>("A VRF input, for instance"), x = 1/89327492234, the four messages,
// x_i = 89327492234 * (i + 1)), the bincode round trip between prover and verifier, and Err for tampered inputs.
// Group elements come from the engine (constant-time MSM for x*G, x*H; zkp_decompress_batch for limb forms); the
// hash-to-group encodings are constants pinned by oracle/ristretto.py against RFC 9496 and libsodium.
// Build + run: tests/test_cpp_port.py.  Exit code 0 = all passed.
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/zkp_b200.h"
#include "../../include/zkp_b200_host.h"
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
static const Enc G_ENC = hex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76");   // RISTRETTO_BASEPOINT_COMPRESSED
static const Enc H_VRF = hex("8062d869a1a967d6a60604a3ec8d0316cef712e094f4cb991de60a9f52555068");   // "A VRF input, for instance"
static const char* MESSAGES[4] = {"One message", "Another message", "A third message", "A fourth message"};
static const Enc H_MSG[4] = {hex("d00abb78261cfa1877d84e7089a90438d2a6aa55a7aa17692499a52e052af444"),
                             hex("7c9bcc52b97705875dc6bd12d692db896c53a58a768a228f6a1b24ef34611f68"),
                             hex("74612243c13ee91af2ae1bcf64144367a32c087319917610845742e78bbe5e31"),
                             hex("acb7811f5bab4e2326c4c41baaba72d6cf1827103d95bf4f37f87c45a3c00649")};
// Scalar::from(89327492234u64).invert()
static const Enc X_INV = hex("0227157482356a105d6545d92cc257e68b4eb62f121e622276bccfcd6cf57e07");

static Limbs decompress(const Enc& e) {
  Limbs l;
  uint8_t valid = 0;
  CHECK(zkp_decompress_batch(ctx, e.data(), 1, l.data(), &valid) == ZKP_OK && valid == 1);
  return l;
}
static Limbs mul(const Limbs& P, const Scalar& x) {
  uint8_t sb[32];
  x.to_bytes(sb);
  uint64_t off[2] = {0, 1};
  Enc out;
  CHECK(zkp_msm_ct_batched(ctx, sb, P.data(), ZKP_POINTS_LIMBS51, off, 1, out.data()) == ZKP_OK);
  return decompress(out);
}
static Transcript transcript(const char* label) { return Transcript((const uint8_t*)label, strlen(label)); }

// define_proof! {dleq, "DLEQ Example Proof", (x), (A, B, H), (G) : A = (x * G), B = (x * H) }
static Statement dleq() {
  Statement st;
  st.name = "dleq";
  st.label = "DLEQ Example Proof";
  st.secrets = {"x"};
  st.instance = {"A", "B", "H"};
  st.common = {"G"};
  // point indices over instance ++ common: A 0, B 1, H 2, G 3
  st.constraints = {{0, {{0, 3}}}, {1, {{0, 2}}}};
  return st;
}

// bincode::serialize + bincode::deserialize of a CompactProof / BatchableProof
static CompactProof roundtrip(const CompactProof& p) {
  const size_t m = p.responses.size();
  std::vector<uint8_t> resp(32 * m), wire(zkph_compact_proof_size(m));
  uint8_t chal[32];
  p.challenge.to_bytes(chal);
  for (size_t i = 0; i < m; i++) p.responses[i].to_bytes(&resp[32 * i]);
  CHECK(zkph_compact_proof_serialize(chal, resp.data(), m, wire.data()) == 0);
  CHECK(wire.size() == 32 + 8 + 32 * m);
  std::vector<uint8_t> r2(32 * m);
  uint8_t c2[32];
  size_t m_out = 0, used = 0;
  CHECK(zkph_compact_proof_parse(wire.data(), wire.size(), m, c2, r2.data(), &m_out, &used) == 0);
  CHECK(m_out == m && used == wire.size());
  CompactProof q;
  CHECK(Scalar::from_canonical_bytes(&q.challenge, c2));
  q.responses.resize(m);
  for (size_t i = 0; i < m; i++) CHECK(Scalar::from_canonical_bytes(&q.responses[i], &r2[32 * i]));
  return q;
}
static BatchableProof roundtrip(const BatchableProof& p) {
  const size_t k = p.commitments.size(), m = p.responses.size();
  std::vector<uint8_t> com(32 * k), resp(32 * m), wire(zkph_batchable_proof_size(k, m));
  for (size_t i = 0; i < k; i++) memcpy(&com[32 * i], p.commitments[i].data(), 32);
  for (size_t i = 0; i < m; i++) p.responses[i].to_bytes(&resp[32 * i]);
  CHECK(zkph_batchable_proof_serialize(com.data(), k, resp.data(), m, wire.data()) == 0);
  std::vector<uint8_t> c2(32 * k), r2(32 * m);
  int64_t bad = 0;
  CHECK(zkph_batchable_proofs_parse(wire.data(), wire.size(), 1, k, m, c2.data(), r2.data(), 1, &bad) == 0);
  BatchableProof q;
  q.commitments.resize(k);
  q.responses.resize(m);
  for (size_t i = 0; i < k; i++) memcpy(q.commitments[i].data(), &c2[32 * i], 32);
  for (size_t i = 0; i < m; i++) CHECK(Scalar::from_canonical_bytes(&q.responses[i], &r2[32 * i]));
  return q;
}

struct Proved { CompactProof compact; BatchableProof batchable; std::vector<Enc> points; };   // points: A, B, H, G

static Proved prove(const Scalar& x, const Enc& H_enc, bool batchable, const char* seed) {
  const Statement st = dleq();
  const Limbs G = decompress(G_ENC), H = decompress(H_enc);
  const Limbs A = mul(G, x), B = mul(H, x);
  Transcript t = transcript("DLEQTest");
  Rng rng((const uint8_t*)seed, strlen(seed));
  Proved out;
  // dleq::ProveAssignments { x, A, B, G, H } -> allocation order secrets, instance (A, B, H), common (G)
  ProofError e = stmt_prove(ctx, st, &t, {x}, {A, B, H, G}, rng, batchable ? nullptr : &out.compact,
                            batchable ? &out.batchable : nullptr, &out.points);
  CHECK(e == PROOF_OK && out.points.size() == 4 && out.points[2] == H_enc && out.points[3] == G_ENC);
  return out;
}

static void create_and_verify_compact() {
  Scalar x;
  CHECK(Scalar::from_canonical_bytes(&x, X_INV.data()));
  CHECK(sc_mul(x, Scalar::from_u128(89327492234ull, 0)) == Scalar::from_u128(1, 0));   // it is the inverse
  Proved p = prove(x, H_VRF, false, "zkp-rs-compact");
  CompactProof parsed = roundtrip(p.compact);
  Transcript t = transcript("DLEQTest");
  CHECK(stmt_verify_compact(ctx, dleq(), &t, {p.points[0], p.points[1], H_VRF, G_ENC}, parsed) == PROOF_OK);
  Transcript t2 = transcript("DLEQTest");
  CHECK(stmt_verify_compact(ctx, dleq(), &t2, {p.points[1], p.points[0], H_VRF, G_ENC}, parsed) == VerificationFailure);
  Transcript t3 = transcript("DLEQTesu");
  CHECK(stmt_verify_compact(ctx, dleq(), &t3, {p.points[0], p.points[1], H_VRF, G_ENC}, parsed) == VerificationFailure);
}

static void create_and_verify_batchable() {
  Scalar x;
  CHECK(Scalar::from_canonical_bytes(&x, X_INV.data()));
  Proved p = prove(x, H_VRF, true, "zkp-rs-batchable");
  BatchableProof parsed = roundtrip(p.batchable);
  Rng vr((const uint8_t*)"v", 1);
  Transcript t = transcript("DLEQTest");
  CHECK(stmt_verify_batchable(ctx, dleq(), &t, {p.points[0], p.points[1], H_VRF, G_ENC}, parsed, vr) == PROOF_OK);
  parsed.responses[0] = sc_add(parsed.responses[0], Scalar::from_u128(1, 0));
  Transcript t2 = transcript("DLEQTest");
  CHECK(stmt_verify_batchable(ctx, dleq(), &t2, {p.points[0], p.points[1], H_VRF, G_ENC}, parsed, vr) == VerificationFailure);
}

static void create_batch_and_batch_verify() {
  std::vector<BatchableProof> proofs;
  std::vector<Enc> pubkeys, vrf_outputs, hs;
  for (int i = 0; i < 4; i++) {
    const Scalar x = sc_mul(Scalar::from_u128(89327492234ull, 0), Scalar::from_u128((uint64_t)(i + 1), 0));
    Proved p = prove(x, H_MSG[i], true, MESSAGES[i]);
    proofs.push_back(roundtrip(p.batchable));
    pubkeys.push_back(p.points[0]);
    vrf_outputs.push_back(p.points[1]);
    hs.push_back(H_MSG[i]);
  }
  auto run = [&](const std::vector<BatchableProof>& pr, const std::vector<Enc>& A) {
    std::vector<Transcript> transcripts(4, transcript("DLEQTest"));
    Rng rng((const uint8_t*)"batch", 5);
    // dleq::BatchVerifyAssignments { A: pubkeys, B: vrf_outputs, H: ..., G: BASEPOINT_COMPRESSED }
    return stmt_batch_verify(ctx, dleq(), &transcripts, {A, vrf_outputs, hs}, {G_ENC}, pr, rng, 1);
  };
  CHECK(run(proofs, pubkeys) == PROOF_OK);
  std::vector<Enc> wrong = pubkeys;
  std::swap(wrong[0], wrong[1]);
  CHECK(run(proofs, wrong) == VerificationFailure);
  std::vector<BatchableProof> bad = proofs;
  bad[2].commitments[0] = bad[2].commitments[1];
  CHECK(run(bad, pubkeys) == VerificationFailure);
  bad = proofs;
  bad.pop_back();
  CHECK(run(bad, pubkeys) == BatchSizeMismatch);
}

int main() {
  if (zkp_ctx_create(&ctx, 0) != ZKP_OK) {
    printf("no CUDA device: this test needs the engine\n");
    return 2;
  }
  create_and_verify_compact();
  create_and_verify_batchable();
  create_batch_and_batch_verify();
  zkp_ctx_destroy(ctx);
  printf(failures ? "%d check(s) FAILED\n" : "all reference tests passed (%d failures)\n", failures);
  return failures ? 1 : 0;
}
