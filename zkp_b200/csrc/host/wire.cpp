// Wire format of the reference's proof types (/root/reference/src/proofs.rs:14-32) as its tests serialize them
// (`bincode::serialize` / `deserialize`, /root/reference/tests/zkp.rs:53-54, :96-97; bincode 1.x default options:
// little-endian, fixed-width integers, `Vec<T>` = u64 length then the items; curve25519-dalek 2.x serializes `Scalar`
// and `CompressedRistretto` as 32-byte tuples without a length prefix and REJECTS non-canonical scalars when
// deserializing [ext]):
//   CompactProof   = challenge[32] | u64 m | responses[m][32]                       32 + 8 + 32 m bytes
//   BatchableProof = u64 k | commitments[k][32] | u64 m | responses[m][32]          8 + 32 k + 8 + 32 m bytes
// parsing a stream of serialized proofs straight into the SoA arrays the batch entry
// points take (commitments[N][k][32], responses[N][m][32]).
#include <string.h>
#include <thread>
#include <vector>

#include "../../../include/zkp_b200_host.h"
#include "scalar.hpp"
#include "toolbox.hpp"

using namespace zkp_host;

static inline uint64_t rd_u64(const uint8_t* p) {
  uint64_t v = 0;
  for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
  return v;
}
static inline void wr_u64(uint8_t* p, uint64_t v) {
  for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i));
}
static bool canonical(const uint8_t* s32) {
  Scalar t;
  return Scalar::from_canonical_bytes(&t, s32);
}

extern "C" size_t zkph_compact_proof_size(size_t m) { return 32 + 8 + 32 * m; }
extern "C" size_t zkph_batchable_proof_size(size_t k, size_t m) { return 8 + 32 * k + 8 + 32 * m; }

extern "C" int32_t zkph_compact_proof_serialize(const uint8_t* challenge, const uint8_t* responses, size_t m, uint8_t* out) {
  if (!challenge || !out || (m && !responses)) return ZKPH_MALFORMED;
  memcpy(out, challenge, 32);
  wr_u64(out + 32, m);
  if (m) memcpy(out + 40, responses, 32 * m);
  return PROOF_OK;
}
extern "C" int32_t zkph_batchable_proof_serialize(const uint8_t* commitments, size_t k, const uint8_t* responses, size_t m,
                                                  uint8_t* out) {
  if (!out || (k && !commitments) || (m && !responses)) return ZKPH_MALFORMED;
  wr_u64(out, k);
  if (k) memcpy(out + 8, commitments, 32 * k);
  wr_u64(out + 8 + 32 * k, m);
  if (m) memcpy(out + 16 + 32 * k, responses, 32 * m);
  return PROOF_OK;
}

// One CompactProof: *m_out receives the number of responses found; at most m_cap are copied out.
extern "C" int32_t zkph_compact_proof_parse(const uint8_t* buf, size_t len, size_t m_cap, uint8_t* challenge,
                                            uint8_t* responses, size_t* m_out, size_t* consumed) {
  if (!buf || len < 40) return ZKPH_MALFORMED;
  if (!canonical(buf)) return ZKPH_MALFORMED;
  const uint64_t m = rd_u64(buf + 32);
  if (m > (len - 40) / 32) return ZKPH_MALFORMED;
  for (uint64_t i = 0; i < m; i++)
    if (!canonical(buf + 40 + 32 * i)) return ZKPH_MALFORMED;
  if (m_out) *m_out = (size_t)m;
  if (consumed) *consumed = 40 + 32 * (size_t)m;
  if (m > m_cap) return VerificationFailure;   // a proof with the wrong number of responses cannot verify (verifier.rs:82-84)
  if (challenge) memcpy(challenge, buf, 32);
  if (responses && m) memcpy(responses, buf + 40, 32 * (size_t)m);
  return PROOF_OK;
}

// N serialized BatchableProofs of ONE statement (k commitments, m responses each), concatenated: -> SoA.
// A proof whose counts differ from (k, m) makes the batch VerificationFailure (batch_verifier.rs:142-148; BatchSizeMismatch
// there is only proofs.len() != batch_size, :138-140); truncated input, trailing bytes or a non-canonical scalar make it
// ZKPH_MALFORMED (the reference's bincode::deserialize fails).
extern "C" int32_t zkph_batchable_proofs_parse(const uint8_t* buf, size_t len, size_t N, size_t k, size_t m,
                                               uint8_t* commitments, uint8_t* responses, int32_t threads,
                                               int64_t* first_bad) {
  if (first_bad) *first_bad = -1;
  if (N && (!buf || (k && !commitments) || (m && !responses))) return ZKPH_MALFORMED;
  const size_t one = zkph_batchable_proof_size(k, m);
  if (N && len / N == one && len % N == 0) try {
    // the common case: every proof has the expected shape, so proof j sits at j * one -- parse in parallel
    int nthr = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nthr < 1) nthr = 1;
    std::vector<int64_t> bad((size_t)nthr + 1, -1);
    std::vector<int32_t> code(bad.size(), PROOF_OK);
    parallel_for(N, nthr, [&](size_t lo, size_t hi, int tid) {
      for (size_t j = lo; j < hi; j++) {
        const uint8_t* p = buf + j * one;
        int32_t rc = PROOF_OK;
        if (rd_u64(p) != k || rd_u64(p + 8 + 32 * k) != m) rc = ZKPH_MALFORMED + 100;   // shape differs: slow path decides
        else
          for (size_t i = 0; i < m && rc == PROOF_OK; i++)
            if (!canonical(p + 16 + 32 * k + 32 * i)) rc = ZKPH_MALFORMED;
        if (rc != PROOF_OK) {
          if (bad[tid] < 0) { bad[tid] = (int64_t)j; code[tid] = rc; }
          return;
        }
        if (k) memcpy(commitments + j * k * 32, p + 8, 32 * k);
        if (m) memcpy(responses + j * m * 32, p + 16 + 32 * k, 32 * m);
      }
    });
    int64_t fb = -1;
    int32_t rc = PROOF_OK;
    for (size_t t = 0; t < bad.size(); t++)
      if (bad[t] >= 0 && (fb < 0 || bad[t] < fb)) { fb = bad[t]; rc = code[t]; }
    if (rc == PROOF_OK) return PROOF_OK;
    if (rc == ZKPH_MALFORMED) { if (first_bad) *first_bad = fb; return rc; }
    // fall through to the sequential walk for the exact error of a batch with irregular shapes
  } catch (...) {
    // no memory for the worker bookkeeping: the sequential walk below needs none
  }
  size_t off = 0;
  for (size_t j = 0; j < N; j++) {
    if (first_bad) *first_bad = (int64_t)j;
    if (len - off < 8) return ZKPH_MALFORMED;
    const uint64_t kk = rd_u64(buf + off);
    if (kk > (len - off - 8) / 32) return ZKPH_MALFORMED;
    size_t q = off + 8 + 32 * (size_t)kk;
    if (len - q < 8) return ZKPH_MALFORMED;
    const uint64_t mm = rd_u64(buf + q);
    if (mm > (len - q - 8) / 32) return ZKPH_MALFORMED;
    for (uint64_t i = 0; i < mm; i++)
      if (!canonical(buf + q + 8 + 32 * i)) return ZKPH_MALFORMED;
    if (kk != k || mm != m) return VerificationFailure;   // batch_verifier.rs:142-148
    if (k) memcpy(commitments + j * k * 32, buf + off + 8, 32 * k);
    if (m) memcpy(responses + j * m * 32, buf + q + 8, 32 * m);
    off = q + 8 + 32 * (size_t)mm;
  }
  if (off != len) { if (first_bad) *first_bad = (int64_t)N; return ZKPH_MALFORMED; }   // trailing bytes
  if (first_bad) *first_bad = -1;
  return PROOF_OK;
}
