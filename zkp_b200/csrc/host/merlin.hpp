// Merlin transcripts (STROBE-128 over Keccak-f[1600]) on the host.
//
// Replaces the `merlin = "2"` crate (/root/reference/Cargo.toml:21) [ext] as used by the reference's
// TranscriptProtocol (/root/reference/src/toolbox/mod.rs:165-228) and by the prover's synthetic-nonce RNG
// (/root/reference/src/toolbox/prover.rs:78-89).  Hashing stays on the host (north_star); the byte stream must be
// reproduced exactly so that challenges agree with the reference.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace zkp_host {

void keccak_f1600(uint64_t st[25]);

class Strobe128 {
 public:
  Strobe128() {}
  explicit Strobe128(const uint8_t* protocol_label, size_t len);
  void meta_ad(const uint8_t* data, size_t len, bool more);
  void ad(const uint8_t* data, size_t len, bool more);
  void prf(uint8_t* out, size_t len, bool more);
  void key(const uint8_t* data, size_t len, bool more);
  // 53 little-endian words: 25 lanes as (lo, hi), pos, pos_begin, cur_flags -- the layout zkp_batch_verify_proofs takes
  void export_state(uint32_t out[53]) const;

 private:
  static const int R = 166;
  enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
  union {
    uint64_t lanes[25];
    uint8_t bytes[200];
  } st_;
  uint8_t pos_ = 0, pos_begin_ = 0, cur_flags_ = 0;
  void run_f();
  void absorb(const uint8_t* data, size_t len);
  void overwrite(const uint8_t* data, size_t len);
  void squeeze(uint8_t* out, size_t len);
  void begin_op(uint8_t flags, bool more);
};

class TranscriptRng {
 public:
  explicit TranscriptRng(const Strobe128& s) : strobe_(s) {}
  void fill_bytes(uint8_t* dest, size_t len);

 private:
  Strobe128 strobe_;
};

class TranscriptRngBuilder {
 public:
  explicit TranscriptRngBuilder(const Strobe128& s) : strobe_(s) {}
  void rekey_with_witness_bytes(const uint8_t* label, size_t llen, const uint8_t* witness, size_t wlen);
  // `entropy32` stands in for the 32 bytes the reference draws from thread_rng (prover.rs:82)
  TranscriptRng finalize(const uint8_t entropy32[32]);

 private:
  Strobe128 strobe_;
};

class Transcript {
 public:
  Transcript() {}
  Transcript(const uint8_t* label, size_t len);
  void append_message(const uint8_t* label, size_t llen, const uint8_t* msg, size_t mlen);
  void challenge_bytes(const uint8_t* label, size_t llen, uint8_t* dest, size_t dlen);
  TranscriptRngBuilder build_rng() const { return TranscriptRngBuilder(strobe_); }
  void export_state(uint32_t out[53]) const { strobe_.export_state(out); }

 private:
  Strobe128 strobe_;
};

}  // namespace zkp_host
