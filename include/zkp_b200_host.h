/* zkp_b200_host.h -- flat C face of the host-side mirror of zkp's toolbox (zkp_b200/csrc/host/ .cpp sources).
 *
 * The C++ classes there (Prover, Verifier, BatchVerifier, Statement = the define_proof! mirror) follow
 * /root/reference/src/toolbox/{prover,verifier,batch_verifier,mod}.rs and /root/reference/src/macros.rs:206-370;
 * this header only makes them callable from ctypes for the parity tests and the benches.  All MSM / compress /
 * decompress work inside goes through the device entry points of zkp_b200.h.
 *
 * A statement is described by flat arrays (what define_proof!{name, label, (secrets), (instance), (common) : ...}
 * expands to): NUL-separated label strings and constraints in CSR form; point indices run over instance ++ common.
 * Return values are the reference's ProofError: 0 Ok, 1 VerificationFailure, 2 BatchSizeMismatch, 3 engine failure.
 *
 * RANDOMNESS.  The reference draws from rand::thread_rng(): 32 bytes of prover entropy per proof (prover.rs:82) and the
 * 128-bit weights of verify_batchable / batch_verify (verifier.rs:153, batch_verifier.rs:179).  Every `rng_seed`,
 * `rho_seed32` and `entropy` argument below may be NULL: the library then draws fresh bytes from the OS CSPRNG
 * (getrandom) for that call -- the default a deployment wants.  A seed that IS passed is used as is, for reproducible
 * tests and benches; it must then itself be fresh, secret CSPRNG output, per call and per shard of a sharded batch:
 * weights the prover can predict (a constant, reused or public seed) let a forger pick responses whose weighted errors
 * cancel, which voids batch soundness, and a repeated prover seed repeats nonces, which reveals witnesses.
 */
#ifndef ZKP_B200_HOST_H
#define ZKP_B200_HOST_H
#include <stddef.h>
#include <stdint.h>
#include "zkp_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct zkph_statement zkph_statement;

/* labels: `n_secrets + n_instance + n_common` NUL-terminated strings back to back (secrets, instance, common).
 * constraint i: lhs point index lhs[i]; terms cons_off[i]..cons_off[i+1]-1 of (term_scalar[], term_point[]).  */
zkph_statement* zkph_statement_new(const char* name, const char* label, const char* labels, int32_t n_secrets,
                                   int32_t n_instance, int32_t n_common, int32_t n_constraints, const int32_t* lhs,
                                   const int32_t* cons_off, const int32_t* term_scalar, const int32_t* term_point);
void zkph_statement_free(zkph_statement* st);

/* module::prove_compact / prove_batchable (macros.rs:261-278) for ONE proof.
 * secrets[m][32], points[p][20 u64] (limb form, allocation order), rng_seed = stand-in for thread_rng.
 * Outputs: encodings[p][32]; compact: challenge[32] + responses[m][32]; batchable: commitments[k][32] + responses. */
int32_t zkph_prove(zkp_ctx* ctx, const zkph_statement* st, const uint8_t* transcript_label, size_t tl_len,
                   const uint8_t* secrets, const uint64_t* points, const uint8_t* rng_seed, size_t seed_len,
                   int32_t batchable, uint8_t* encodings, uint8_t* challenge, uint8_t* commitments, uint8_t* responses,
                   uint8_t* blindings_out /* may be NULL */);

/* module::verify_compact (macros.rs:314-322) */
int32_t zkph_verify_compact(zkp_ctx* ctx, const zkph_statement* st, const uint8_t* transcript_label, size_t tl_len,
                            const uint8_t* points_enc, const uint8_t* challenge, const uint8_t* responses,
                            size_t n_responses);
/* module::verify_batchable (macros.rs:325-333) */
int32_t zkph_verify_batchable(zkp_ctx* ctx, const zkph_statement* st, const uint8_t* transcript_label, size_t tl_len,
                              const uint8_t* points_enc, const uint8_t* commitments, size_t n_commitments,
                              const uint8_t* responses, size_t n_responses, const uint8_t* rng_seed, size_t seed_len);

/* module::batch_verify (macros.rs:336-370): instance_enc[n_instance][N][32], common_enc[n_common][32],
 * commitments[N][k][32], responses[N][m][32].  coeff_out (optional) receives what the reference feeds the MSM:
 * static_coeffs[num_s][32] ++ row-major instance matrix[(n_instance+k)][N][32]; points_out likewise.          */
int32_t zkph_batch_verify(zkp_ctx* ctx, const zkph_statement* st, const uint8_t* transcript_label, size_t tl_len,
                          size_t N, const uint8_t* instance_enc, const uint8_t* common_enc, const uint8_t* commitments,
                          const uint8_t* responses, const uint8_t* rng_seed, size_t seed_len, int32_t threads,
                          uint8_t* coeff_out, uint8_t* points_out, double* host_seconds /* hashing+fold, may be NULL */);

/* module::batch_verify with the reference's own signature: one transcript handle per proof (each with its own prior
 * state; every handle is left advanced, like the reference's `&mut Transcript`s).  n_transcripts != n_proofs is
 * BatchSizeMismatch (batch_verifier.rs:72-74).                                                                         */
typedef struct zkph_transcript zkph_transcript;
int32_t zkph_batch_verify_t(zkp_ctx* ctx, const zkph_statement* st, zkph_transcript* const* transcripts,
                            size_t n_transcripts, size_t n_proofs, const uint8_t* instance_enc, const uint8_t* common_enc,
                            const uint8_t* commitments, const uint8_t* responses, const uint8_t* rng_seed, size_t seed_len,
                            int32_t threads);

/* module::batch_verify with the per-proof work on the DEVICE (SURVEY.md 8f rows f1 + f2): the host hashes only the
 * batch-wide transcript prefix (user transcript + dom-sep + scalar labels) and hands everything else to
 * zkp_batch_verify_proofs.  Same inputs as zkph_batch_verify; the weights are derived per proof from rho_seed[32]
 * (see zkp_b200.h), so the MSM inputs equal zkph_batch_verify's only when that one is given the same derivation.  */
int32_t zkph_batch_verify_device(zkp_ctx* ctx, const zkph_statement* st, const uint8_t* transcript_label, size_t tl_len,
                                 size_t N, const uint8_t* instance_enc, const uint8_t* common_enc,
                                 const uint8_t* commitments, const uint8_t* responses, const uint8_t* rho_seed32,
                                 uint8_t* coeff_out, uint8_t* points_out);

/* N independent batchable proofs of one statement in three batched device calls (BASELINE configs[1]):
 * secrets[N][m][32], points[N][p][20 u64], entropy[N][32] -> encodings[N][p][32], commitments[N][k][32],
 * responses[N][m][32].  Per-proof results equal zkph_prove(batchable) with the same entropy.                 */
int32_t zkph_prove_many(zkp_ctx* ctx, const zkph_statement* st, const uint8_t* transcript_label, size_t tl_len, size_t N,
                        const uint8_t* secrets, const uint64_t* points, const uint8_t* entropy, int32_t threads,
                        uint8_t* encodings, uint8_t* commitments, uint8_t* responses);

/* The same N proofs with the per-proof transcript / nonce / response work on the DEVICE (zkp_prove_batch, SURVEY.md 8f-1
 * for the prover): the host hashes only the batch-wide transcript prefix.  Byte-identical to zkph_prove_many.          */
int32_t zkph_prove_many_device(zkp_ctx* ctx, const zkph_statement* st, const uint8_t* transcript_label, size_t tl_len,
                               size_t N, const uint8_t* secrets, const uint64_t* points, const uint8_t* entropy,
                               uint8_t* encodings, uint8_t* commitments, uint8_t* responses);

/* Wire format of /root/reference/src/proofs.rs as the reference's tests serialize it (bincode 1.x defaults,
 * tests/zkp.rs:53-54, :96-97):  CompactProof = challenge[32] | u64 m | responses[m][32];
 * BatchableProof = u64 k | commitments[k][32] | u64 m | responses[m][32]  (little-endian; scalars must be canonical).
 * Return codes: 0 ok, 1 VerificationFailure, 2 BatchSizeMismatch, ZKPH_MALFORMED for input bincode would refuse.   */
#define ZKPH_MALFORMED 4
size_t zkph_compact_proof_size(size_t m);
size_t zkph_batchable_proof_size(size_t k, size_t m);
int32_t zkph_compact_proof_serialize(const uint8_t* challenge, const uint8_t* responses, size_t m, uint8_t* out);
int32_t zkph_batchable_proof_serialize(const uint8_t* commitments, size_t k, const uint8_t* responses, size_t m,
                                       uint8_t* out);
int32_t zkph_compact_proof_parse(const uint8_t* buf, size_t len, size_t m_cap, uint8_t* challenge, uint8_t* responses,
                                 size_t* m_out, size_t* consumed);
/* N concatenated serialized BatchableProofs of one statement -> the SoA arrays of zkph_batch_verify(_device):
 * commitments[N][k][32], responses[N][m][32].  first_bad (optional) = index of the offending proof.             */
int32_t zkph_batchable_proofs_parse(const uint8_t* buf, size_t len, size_t N, size_t k, size_t m, uint8_t* commitments,
                                    uint8_t* responses, int32_t threads, int64_t* first_bad);

/* merlin::Transcript handles, for callers that drive the transcript themselves the way the reference's API takes
 * `&mut Transcript` (e.g. /root/reference/tests/sig_and_vrf_example.rs:86-125: messages are appended before proving).
 * The *_t entry points below mutate the handle exactly like the reference mutates its transcript.               */
zkph_transcript* zkph_transcript_new(const uint8_t* label, size_t len);
zkph_transcript* zkph_transcript_clone(const zkph_transcript* t);
void zkph_transcript_free(zkph_transcript* t);
void zkph_transcript_append_message(zkph_transcript* t, const uint8_t* label, size_t llen, const uint8_t* msg, size_t mlen);
void zkph_transcript_export_state(const zkph_transcript* t, uint32_t out53[53]);   /* the prefix_state of zkp_b200.h */
void zkph_transcript_challenge_bytes(zkph_transcript* t, const uint8_t* label, size_t llen, uint8_t* out, size_t n);
int32_t zkph_prove_t(zkp_ctx* ctx, const zkph_statement* st, zkph_transcript* t, const uint8_t* secrets,
                     const uint64_t* points, const uint8_t* rng_seed, size_t seed_len, int32_t batchable,
                     uint8_t* encodings, uint8_t* challenge, uint8_t* commitments, uint8_t* responses,
                     uint8_t* blindings_out);
int32_t zkph_verify_compact_t(zkp_ctx* ctx, const zkph_statement* st, zkph_transcript* t, const uint8_t* points_enc,
                              const uint8_t* challenge, const uint8_t* responses, size_t n_responses);
int32_t zkph_verify_batchable_t(zkp_ctx* ctx, const zkph_statement* st, zkph_transcript* t, const uint8_t* points_enc,
                                const uint8_t* commitments, size_t n_commitments, const uint8_t* responses,
                                size_t n_responses, const uint8_t* rng_seed, size_t seed_len);

/* host primitives exposed for parity tests */
void zkph_scalar_mul(uint8_t* out32, const uint8_t* a32, const uint8_t* b32);
void zkph_scalar_from_wide(uint8_t* out32, const uint8_t* in64);
void zkph_merlin_test_vector(uint8_t* out32);
void zkph_rng_bytes(const uint8_t* seed, size_t seed_len, uint8_t* out, size_t n);

#ifdef __cplusplus
}
#endif
#endif
