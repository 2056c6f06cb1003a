#include "merlin.hpp"

namespace zkp_host {

static inline uint64_t rol(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

void keccak_f1600(uint64_t a[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
      0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
      0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
      0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  for (int rnd = 0; rnd < 24; rnd++) {
    uint64_t c[5], d[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(a[x + 5 * y], ROT[x + 5 * y]);
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= RC[rnd];
  }
}

Strobe128::Strobe128(const uint8_t* protocol_label, size_t len) {
  memset(st_.bytes, 0, 200);
  const uint8_t init[6] = {1, R + 2, 1, 0, 1, 96};
  memcpy(st_.bytes, init, 6);
  memcpy(st_.bytes + 6, "STROBEv1.0.2", 12);
  keccak_f1600(st_.lanes);
  pos_ = pos_begin_ = cur_flags_ = 0;
  meta_ad(protocol_label, len, false);
}
void Strobe128::run_f() {
  st_.bytes[pos_] ^= pos_begin_;
  st_.bytes[pos_ + 1] ^= 0x04;
  st_.bytes[R + 1] ^= 0x80;
  keccak_f1600(st_.lanes);
  pos_ = 0;
  pos_begin_ = 0;
}
void Strobe128::absorb(const uint8_t* data, size_t len) {
  for (size_t i = 0; i < len; i++) {
    st_.bytes[pos_] ^= data[i];
    if (++pos_ == R) run_f();
  }
}
void Strobe128::overwrite(const uint8_t* data, size_t len) {
  for (size_t i = 0; i < len; i++) {
    st_.bytes[pos_] = data[i];
    if (++pos_ == R) run_f();
  }
}
void Strobe128::squeeze(uint8_t* out, size_t len) {
  for (size_t i = 0; i < len; i++) {
    out[i] = st_.bytes[pos_];
    st_.bytes[pos_] = 0;
    if (++pos_ == R) run_f();
  }
}
void Strobe128::begin_op(uint8_t flags, bool more) {
  if (more) return;  // continuing the current operation (flags must match; callers guarantee it)
  uint8_t old_begin = pos_begin_;
  pos_begin_ = pos_ + 1;
  cur_flags_ = flags;
  uint8_t hdr[2] = {old_begin, flags};
  absorb(hdr, 2);
  if ((flags & (FLAG_C | FLAG_K)) && pos_ != 0) run_f();
}
void Strobe128::meta_ad(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_M | FLAG_A, more); absorb(d, n); }
void Strobe128::ad(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_A, more); absorb(d, n); }
void Strobe128::prf(uint8_t* out, size_t n, bool more) { begin_op(FLAG_I | FLAG_A | FLAG_C, more); squeeze(out, n); }
void Strobe128::key(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_A | FLAG_C, more); overwrite(d, n); }

static inline void u32le(uint8_t out[4], size_t n) {
  out[0] = (uint8_t)n; out[1] = (uint8_t)(n >> 8); out[2] = (uint8_t)(n >> 16); out[3] = (uint8_t)(n >> 24);
}

Transcript::Transcript(const uint8_t* label, size_t len) : strobe_((const uint8_t*)"Merlin v1.0", 11) {
  append_message((const uint8_t*)"dom-sep", 7, label, len);
}
void Transcript::append_message(const uint8_t* label, size_t llen, const uint8_t* msg, size_t mlen) {
  uint8_t l4[4];
  u32le(l4, mlen);
  strobe_.meta_ad(label, llen, false);
  strobe_.meta_ad(l4, 4, true);
  strobe_.ad(msg, mlen, false);
}
void Transcript::challenge_bytes(const uint8_t* label, size_t llen, uint8_t* dest, size_t dlen) {
  uint8_t l4[4];
  u32le(l4, dlen);
  strobe_.meta_ad(label, llen, false);
  strobe_.meta_ad(l4, 4, true);
  strobe_.prf(dest, dlen, false);
}
void TranscriptRngBuilder::rekey_with_witness_bytes(const uint8_t* label, size_t llen, const uint8_t* w, size_t wlen) {
  uint8_t l4[4];
  u32le(l4, wlen);
  strobe_.meta_ad(label, llen, false);
  strobe_.meta_ad(l4, 4, true);
  strobe_.key(w, wlen, false);
}
TranscriptRng TranscriptRngBuilder::finalize(const uint8_t entropy32[32]) {
  strobe_.meta_ad((const uint8_t*)"rng", 3, false);
  strobe_.key(entropy32, 32, false);
  return TranscriptRng(strobe_);
}
void TranscriptRng::fill_bytes(uint8_t* dest, size_t len) {
  uint8_t l4[4];
  u32le(l4, len);
  strobe_.meta_ad(l4, 4, false);
  strobe_.prf(dest, len, false);
}

}  // namespace zkp_host
