/* zkp_b200.h -- C ABI of the B200-native ristretto255 multi-scalar-multiplication engine.
 *
 * The reference (dalek-cryptography/zkp) has no FFI: its hot path is a Rust trait call on a foreign type,
 *     RistrettoPoint::multiscalar_mul / vartime_multiscalar_mul / optional_multiscalar_mul,
 *     CompressedRistretto::decompress, RistrettoPoint::compress, IsIdentity::is_identity
 * (curve25519-dalek ^2, /root/reference/Cargo.toml:27).  Each entry point below names the reference call
 * site(s) whose work it replaces; INTEGRATION.md shows the Rust `extern "C"` block and the nine-line patch of
 * zkp's `toolbox` that binds them.
 *
 * Conventions
 *   - scalars: 32 bytes little-endian, canonical (< l = 2^252 + 27742317777372353535851937790883648493),
 *     i.e. curve25519_dalek::scalar::Scalar::as_bytes().
 *   - compressed points: 32-byte ristretto255 encodings (CompressedRistretto::as_bytes()).
 *   - limb-form points: 4 x 5 x uint64, the in-memory layout of dalek's u64-backend
 *     RistrettoPoint(EdwardsPoint{X,Y,Z,T: FieldElement51}), radix 2^51, limbs < 2^54.
 *   - all buffers are caller-owned; nothing is retained after a call returns.  Host entry points take host
 *     pointers (pinned memory recommended) and are synchronous.  `_dev` entry points take device pointers that
 *     are 16-byte aligned, enqueue on the context's stream and do NOT synchronise.
 *   - a context is bound to one device and one stream; calls on one context must be serialised by the caller
 *     (the reference's Prover/Verifier are !Sync as well, /root/reference/src/toolbox/prover.rs:23-31).
 *   - no entry point aborts or throws across the boundary; every failure is a status code.
 */
#ifndef ZKP_B200_H
#define ZKP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zkp_ctx zkp_ctx;

/* status codes */
#define ZKP_OK 0                 /* success                                                                  */
#define ZKP_ERR_POINT 1          /* some encoding failed to decompress  == `None` of optional_multiscalar_mul,
                                    mapped to ProofError::VerificationFailure at verifier.rs:166,
                                    batch_verifier.rs:228                                                    */
#define ZKP_ERR_SIZE 2           /* inconsistent sizes / null or misaligned pointers == ProofError::BatchSizeMismatch
                                    (batch_verifier.rs:72-74,120-122,138-140)                                */
#define ZKP_ERR_SCALAR 3         /* a scalar is not canonical (>= l)                                         */
#define ZKP_ERR_NOGPU (-1)       /* no CUDA device / CUDA runtime failure while creating the context         */
#define ZKP_ERR_CUDA (-2)        /* CUDA failure during a call; zkp_last_error() has the text                */
#define ZKP_ERR_NOMEM (-3)       /* workspace allocation failed                                              */

/* flags for point inputs of zkp_msm_ct_batched */
#define ZKP_POINTS_COMPRESSED 0  /* 32-byte encodings  */
#define ZKP_POINTS_LIMBS51 1     /* 4x5xuint64 limb form */

int32_t zkp_device_count(void);
/* Create a context on CUDA device `device` with its own non-blocking stream. */
int32_t zkp_ctx_create(zkp_ctx** out, int32_t device);
void zkp_ctx_destroy(zkp_ctx* ctx);
/* Enqueue on the caller's stream (a cudaStream_t) from now on; NULL restores the context's own stream. */
int32_t zkp_ctx_set_stream(zkp_ctx* ctx, void* cuda_stream);
/* Tunables: "window" (Pippenger window width c, 0 = choose from n), "window_cap" (upper bound of the automatic
 * choice), "chunk" (max sorted entries per accumulate work item, 0 = auto), "profile" (0/1, see
 * zkp_ctx_stage_ms; forces the stages to run back to back), "overlap" (0/1, default 0: run the digit sort on a
 * second stream concurrently with decompression; measured neutral), "chunk_terms" (terms per H2D chunk of the host-input MSM,
 * default 2^18), "dual_stream" (0/1, default 1: the chunk kernels alternate between two streams), "ramp_chunks" (0/1,
 * default 0: growing chunk sizes; measured slower), "phase1_percent" (share of the points decompressed under the digit
 * histogram, default 54), "bv_chunk_terms" (slab size of zkp_batch_verify_proofs: 4 x this / rows proofs, default 3 * 2^17),
 * "bv_phase1_rows" (rows of a slab decompressed under the histogram: 0 = chosen per slab from the arrival of the copies,
 * default; n > 0 = fixed; -1 = alternating, for tests), "bv_prep_stream" (0/1, default 1: the front-end kernel of a slab as a
 * resident grid on its own high-priority stream), "bv_prep_blocks" / "bv_prep_smem_kb" (that grid: blocks per SM, default 2,
 * and the unused dynamic shared memory that caps its residency, default 64), "scatter_batch" (0/1, default 1), "fused_sort",
 * "balance", "ingest_variant", "bv_compiled", "share_static_tables", "prove_chunk", "coop_max_msms" (ablation
 * switches documented in DESIGN.md).  Environment: ZKP_BV_TIMELINE=1 makes zkp_batch_verify_proofs print, on stderr, when
 * every copy, front-end kernel and ingestion launch of the call ended (timed CUDA events; a development aid). */
int32_t zkp_ctx_set_option(zkp_ctx* ctx, const char* key, int64_t value);
int32_t zkp_ctx_synchronize(zkp_ctx* ctx);
const char* zkp_last_error(zkp_ctx* ctx);
/* Diagnostic: with option "profile"=1 every zkp_msm_vartime[_dev] call synchronises and records the device
 * time of its stages: 0 decompress, 1 recode+histogram, 2 scan, 3 scatter, 4 bucket accumulate, 5 bucket
 * reduce, 6 finish (ms).  stage 100 / 101 return the window width / chunk length last used.          */
double zkp_ctx_stage_ms(zkp_ctx* ctx, int32_t stage);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t zkp_ctx_launch_count(zkp_ctx* ctx);

/* CompressedRistretto::decompress over a batch.
 * Replaces the per-point loop at /root/reference/src/toolbox/verifier.rs:87-92 (and the lazy
 * `.map(|pt| pt.decompress())` at verifier.rs:164, batch_verifier.rs:226 when used stand-alone).
 * limbs_out[n][4][5] receives (X,Y,Z=1,T) in FieldElement51 form; valid_out[n] is 1 where decoding succeeded
 * (limbs are the identity where it failed).  Returns ZKP_OK even if some points are invalid.              */
int32_t zkp_decompress_batch(zkp_ctx* ctx, const uint8_t* enc, size_t n, uint64_t* limbs_out, uint8_t* valid_out);

/* RistrettoPoint::compress over a batch.
 * Replaces /root/reference/src/toolbox/mod.rs:180 (append_point_var) and :204 (append_blinding_commitment). */
int32_t zkp_compress_batch(zkp_ctx* ctx, const uint64_t* limbs_in, size_t n, uint8_t* enc_out);

/* One variable-time MSM over compressed points: out32 = compress( sum_i scalars[i] * decompress(points[i]) ).
 * Replaces RistrettoPoint::optional_multiscalar_mul(...) with `.map(|pt| pt.decompress())` at
 * /root/reference/src/toolbox/batch_verifier.rs:219-228 and verifier.rs:162-166, and
 * vartime_multiscalar_mul at verifier.rs:97-106.  The identity test of batch_verifier.rs:230 /
 * verifier.rs:168 is `out32 == 32 zero bytes` (also returned in *is_identity when non-NULL).
 * Returns ZKP_ERR_POINT (and the first failing index in *first_bad when non-NULL) where the reference
 * returns None.  n == 0 yields the identity.                                                              */
int32_t zkp_msm_vartime(zkp_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out32,
                        int32_t* is_identity, int64_t* first_bad);

/* Same, device-resident and asynchronous.  d_result (device, 48 bytes, 16-aligned) receives
 *   bytes 0..31 the encoding, int32 @32 status (ZKP_OK / ZKP_ERR_POINT / ZKP_ERR_SCALAR),
 *   int32 @36 is_identity, int64 @40 first_bad (-1 if none).                                              */
int32_t zkp_msm_vartime_dev(zkp_ctx* ctx, const void* d_scalars, const void* d_points, size_t n, void* d_result);

/* M independent variable-time MSMs described CSR-style: MSM j uses terms offsets[j] .. offsets[j+1]-1.
 * Replaces the per-proof loops around vartime_multiscalar_mul (/root/reference/src/toolbox/verifier.rs:96-110)
 * and optional_multiscalar_mul (verifier.rs:162) when many proofs are verified individually.
 * out[M][32]; valid[M] = 0 where a point of that MSM failed to decode (== None).                          */
int32_t zkp_msm_vartime_batched(zkp_ctx* ctx, const uint8_t* scalars, const uint8_t* points,
                                const uint64_t* offsets, size_t M, uint8_t* out, uint8_t* valid);

/* M independent CONSTANT-TIME MSMs followed by compression: out[j] = compress( sum scalars * points ).
 * Replaces RistrettoPoint::multiscalar_mul + append_blinding_commitment's compress at
 * /root/reference/src/toolbox/prover.rs:93-103.  No branch or memory address depends on the scalars.
 * `point_format` selects ZKP_POINTS_LIMBS51 (the prover holds RistrettoPoints, prover.rs:27) or
 * ZKP_POINTS_COMPRESSED.  Compressed inputs that fail to decode give ZKP_ERR_POINT.                        */
int32_t zkp_msm_ct_batched(zkp_ctx* ctx, const uint8_t* scalars, const void* points, int32_t point_format,
                           const uint64_t* offsets, size_t M, uint8_t* out);

/* BatchVerifier::verify_batchable from the MSM on (/root/reference/src/toolbox/batch_verifier.rs:208-234):
 * given the coefficient vector and the point rows exactly as the reference builds them,
 *   static_coeffs[num_s], static_points[num_s],
 *   instance_coeffs = row-major Matrix[(rows) x batch]  (util.rs:35-37, batch_verifier.rs:174,222),
 *   instance_points[rows][batch]                        (batch_verifier.rs:208-217; rows = num_i + num_c),
 * compute the combined MSM and the identity test.  *accept = 1 iff the sum is the ristretto identity.
 * Returns ZKP_ERR_POINT when a point fails to decode (the reference returns VerificationFailure, :228).     */
int32_t zkp_batch_verify(zkp_ctx* ctx, const uint8_t* static_coeffs, const uint8_t* static_points, size_t num_s,
                         const uint8_t* instance_coeffs, const uint8_t* instance_points, size_t rows, size_t batch,
                         int32_t* accept, int64_t* first_bad);

/* SURVEY.md section 8f rows f1 + f2 ("next"): BatchVerifier::verify_batchable with the per-proof work on the device --
 * Merlin transcripts and challenges (/root/reference/src/toolbox/batch_verifier.rs:115-167), random weights (:179),
 * coefficient fold (:176-206) -- followed by the same combined MSM and identity test (:219-234).
 * The statement is the define_proof! expansion (macros.rs:336-370): transcript order = instance points in order,
 * then common (static) points, then one blinding commitment per constraint labelled with its lhs variable.      */
typedef struct {
  int32_t m, ni, nc, k;        /* secrets, instance points, common points, constraints                         */
  const char* labels;          /* ni + nc NUL-terminated point labels back to back (instance first)             */
  const int32_t* lhs;          /* [k] lhs point index over instance ++ common                                   */
  const int32_t* cons_off;     /* [k+1] term ranges                                                             */
  const int32_t* term_scalar;  /* secret index per term                                                         */
  const int32_t* term_point;   /* point index per term, over instance ++ common                                 */
} zkp_statement_desc;
/* prefix_state: 53 little-endian uint32 = STROBE state (25 lanes as lo,hi), pos, pos_begin, cur_flags, of the
 * transcript every proof starts from (user transcript + dom-sep + scalar labels; zkp_b200_host.h produces it).
 * instance_enc[ni][N][32], common_enc[nc][32], commitments[N][k][32], responses[N][m][32], rho_seed[32]: the
 * weights are rho_(i,j) = bytes [16i,16i+16) of SHAKE256(rho_seed || le64(j)) (stand-in for thread_rng).
 * rho_seed may be NULL: the library then draws it from the OS CSPRNG for this call (the deployment default).  A seed
 * that is passed must itself be fresh, secret CSPRNG output per call and per shard of a sharded batch -- with weights
 * a prover can predict (constant, reused or public seed) responses can be chosen so that the weighted errors cancel,
 * and the batch check is void.
 * coeff_out / points_out (optional, n = nc + (ni+k)*N rows of 32 bytes) receive the MSM inputs for parity tests.
 * Returns ZKP_ERR_POINT for an identity or undecodable encoding (VerificationFailure), ZKP_ERR_SCALAR for a
 * non-canonical response, ZKP_ERR_SIZE if the statement exceeds 40 point variables per kind or 64 constraints.  */
int32_t zkp_batch_verify_proofs(zkp_ctx* ctx, const zkp_statement_desc* st, const uint32_t* prefix_state, size_t N,
                                const uint8_t* instance_enc, const uint8_t* common_enc, const uint8_t* commitments,
                                const uint8_t* responses, const uint8_t* rho_seed, int32_t* accept, int64_t* first_bad,
                                uint8_t* coeff_out, uint8_t* points_out);

/* Single-verdict mode over shards (one BatchVerifier batch whose proofs -- the columns of the instance matrix of
 * /root/reference/src/toolbox/batch_verifier.rs:174 -- are cut over several GPUs or calls): zkp_batch_verify_partial takes
 * the same arguments as zkp_batch_verify for ONE shard (the static coefficients may all go to one shard) and returns
 * the shard's MSM sum as an extended point in FieldElement51 limb form (X, Y, Z, T: 20 x u64) instead of a verdict;
 * zkp_partials_verdict adds `count` such points and applies the identity test of batch_verifier.rs:230 to the sum
 * (enc_out32 optional: the encoding of the sum).  The sums cross ranks with one all-gather of 160 bytes per rank.     */
int32_t zkp_batch_verify_partial(zkp_ctx* ctx, const uint8_t* static_coeffs, const uint8_t* static_points, size_t num_s,
                                 const uint8_t* instance_coeffs, const uint8_t* instance_points, size_t rows, size_t batch,
                                 uint64_t* partial_limbs_out, int64_t* first_bad);
int32_t zkp_partials_verdict(zkp_ctx* ctx, const uint64_t* partial_limbs, size_t count, int32_t* accept,
                             uint8_t* enc_out32);

/* The same with device-resident data, asynchronous on the context's stream (what a rank of a sharded batch runs between
 * its MSM and the verdict): zkp_msm_vartime_partial_dev = zkp_msm_vartime_dev that also leaves the sum as 160 bytes of
 * FieldElement51 limbs at d_partial_limbs; after an all-gather of those 160-byte records (NCCL, same stream),
 * zkp_partials_verdict_dev writes the 48-byte result record of zkp_msm_vartime_dev for their sum to d_result
 * (is_identity = the verdict of batch_verifier.rs:230).  All pointers 16-byte aligned device memory.                  */
int32_t zkp_msm_vartime_partial_dev(zkp_ctx* ctx, const void* d_scalars, const void* d_points, size_t n, void* d_result,
                                    void* d_partial_limbs);
int32_t zkp_partials_verdict_dev(zkp_ctx* ctx, const void* d_partial_limbs, size_t count, void* d_result);

/* Batch proving with the per-proof work of Prover::prove_impl (/root/reference/src/toolbox/prover.rs:76-112) on the
 * device: allocate_point compressions (toolbox/mod.rs:180), transcript replay, the synthetic-nonce blindings of
 * prover.rs:78-89 (TranscriptRng rekeyed with every secret and finalized with entropy[j], the stand-in for thread_rng),
 * the k constant-time MSMs + compress per proof (prover.rs:93-103), challenge and responses s*c + b (:106-109).
 * secrets[N][m][32] canonical, points[N][p][20 u64] (p = ni + nc, FieldElement51 limbs X,Y,Z,T, the statement's
 * allocation order instance ++ common), entropy[N][32] -> encodings_out[N][p][32], commitments_out[N][k][32],
 * responses_out[N][m][32]; blindings_out[N][m][32] optional (parity tests).  prefix_state as for
 * zkp_batch_verify_proofs.  Proof j equals what the reference's prover produces from the same transcript, secrets and
 * entropy.  entropy may be NULL: 32 fresh bytes per proof then come from the OS CSPRNG (the deployment default); entropy
 * that is passed must be fresh and secret (a repeated value repeats nonces and reveals the witnesses).  Batches of at
 * least 2^15 proofs run as slices of 2^14 ("prove_pipe_chunk") alternating between two workspaces and streams, so the
 * copies of one slice overlap the kernels of the next; outputs are written only after the slice's checks have passed.
 * Returns ZKP_ERR_SCALAR for a non-canonical secret.                                                               */
int32_t zkp_prove_batch(zkp_ctx* ctx, const zkp_statement_desc* st, const uint32_t* prefix_state, size_t N,
                        const uint8_t* secrets, const uint64_t* points, const uint8_t* entropy, uint8_t* encodings_out,
                        uint8_t* commitments_out, uint8_t* responses_out, uint8_t* blindings_out);

/* Test hook (no GPU needed): compiles the batch-verification transcript script of `st` (what zkp_batch_verify_proofs
 * hands to the device) and executes it on the host for ONE proof (instance_enc[ni][32], commitments[k][32]):
 * challenge_out64 receives the 64 challenge bytes, *n_blocks_out (optional) the number of Keccak-f blocks.          */
int32_t zkp_selftest_bv_script(const zkp_statement_desc* st, const uint32_t* prefix_state, const uint8_t* instance_enc,
                               const uint8_t* common_enc, const uint8_t* commitments, uint8_t* challenge_out64,
                               int32_t* n_blocks_out);

/* Test hook (no GPU needed): the chunk boundaries the host-input pipeline of zkp_msm_vartime / zkp_batch_verify uses for
 * n terms with the "chunk_terms", "phase1_percent" and "ramp_chunks" options given: bounds_out[0..count] (cap entries
 * available), *k1_out = chunks in the first ingestion phase.  Returns the chunk count, -1 on bad arguments.           */
int64_t zkp_selftest_chunk_schedule(size_t n, size_t chunk, int32_t phase1_percent, int32_t ramp, size_t* bounds_out,
                                    size_t cap, size_t* k1_out);

/* Self-test of the device hashing code: out32 receives Merlin's published conformance vector (a8c933f5...). */
int32_t zkp_selftest_hash(zkp_ctx* ctx, uint8_t* out32);

/* Device-side micro-benchmark of the field multiplier variants (roofline calibration):
 * runs `iters` dependent multiplies (or squarings) per thread over a full grid and returns the measured
 * rate in operations per second in *ops_per_sec.  kind: 0 = fe_mul (8x32 saturated), 1 = fe_sq,
 * 2 = 5x51-limb multiply (u64 products), 3 = 10x25.5-limb multiply, 4 / 5 = 32 wide multiplies per
 * iteration without / with carry chains (raw IMAD.WIDE issue rate; ops counted per iteration), 6 / 7 =
 * fe_mul / fe_sq with the shift-add reduction variant, 8 / 9 = fe_mul / fe_sq with the variable-time tail
 * (cold-branch carry fold; what decompression and bucket accumulation run), 10 / 11 = signed mixed point
 * addition (7 M, the bucket-accumulation step) with constant-time / variable-time tails.                  */
int32_t zkp_bench_field(zkp_ctx* ctx, int32_t kind, int32_t iters, double* ops_per_sec);

/* Diagnostic: half of the warps run the integer fe_sq chain, the other half a DFMA chain (160 per fe_sq).
 * mode 0 both, 1 integer warps only, 2 FP64 warps only; *ms_out = kernel time for `iters` squarings per thread. */
int32_t zkp_bench_dual(zkp_ctx* ctx, int32_t mode, int32_t iters, double* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* ZKP_B200_H */
