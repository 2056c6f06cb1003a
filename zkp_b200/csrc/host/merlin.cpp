#include "merlin.hpp"
#include <stdio.h>
#include <stdlib.h>

namespace zkp_host {

static inline uint64_t rol(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

void keccak_f1600(uint64_t a[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
      0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
      0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
      0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  // fully unrolled round: lanes in locals a<xy> = a[x + 5y]
  uint64_t a00 = a[0], a10 = a[1], a20 = a[2], a30 = a[3], a40 = a[4];
  uint64_t a01 = a[5], a11 = a[6], a21 = a[7], a31 = a[8], a41 = a[9];
  uint64_t a02 = a[10], a12 = a[11], a22 = a[12], a32 = a[13], a42 = a[14];
  uint64_t a03 = a[15], a13 = a[16], a23 = a[17], a33 = a[18], a43 = a[19];
  uint64_t a04 = a[20], a14 = a[21], a24 = a[22], a34 = a[23], a44 = a[24];
  for (int rnd = 0; rnd < 24; rnd++) {
    uint64_t c0 = a00 ^ a01 ^ a02 ^ a03 ^ a04, c1 = a10 ^ a11 ^ a12 ^ a13 ^ a14, c2 = a20 ^ a21 ^ a22 ^ a23 ^ a24;
    uint64_t c3 = a30 ^ a31 ^ a32 ^ a33 ^ a34, c4 = a40 ^ a41 ^ a42 ^ a43 ^ a44;
    uint64_t d0 = c4 ^ rol(c1, 1), d1 = c0 ^ rol(c2, 1), d2 = c1 ^ rol(c3, 1), d3 = c2 ^ rol(c4, 1), d4 = c3 ^ rol(c0, 1);
    // theta + rho + pi: b[y][2x+3y] = rol(a[x][y] ^ d[x], ROT[x][y])
    uint64_t b00 = a00 ^ d0,           b13 = rol(a01 ^ d0, 36), b21 = rol(a02 ^ d0, 3),  b34 = rol(a03 ^ d0, 41), b42 = rol(a04 ^ d0, 18);
    uint64_t b02 = rol(a10 ^ d1, 1),  b10 = rol(a11 ^ d1, 44), b23 = rol(a12 ^ d1, 10), b31 = rol(a13 ^ d1, 45), b44 = rol(a14 ^ d1, 2);
    uint64_t b04 = rol(a20 ^ d2, 62), b12 = rol(a21 ^ d2, 6),  b20 = rol(a22 ^ d2, 43), b33 = rol(a23 ^ d2, 15), b41 = rol(a24 ^ d2, 61);
    uint64_t b01 = rol(a30 ^ d3, 28), b14 = rol(a31 ^ d3, 55), b22 = rol(a32 ^ d3, 25), b30 = rol(a33 ^ d3, 21), b43 = rol(a34 ^ d3, 56);
    uint64_t b03 = rol(a40 ^ d4, 27), b11 = rol(a41 ^ d4, 20), b24 = rol(a42 ^ d4, 39), b32 = rol(a43 ^ d4, 8),  b40 = rol(a44 ^ d4, 14);
    // chi
    a00 = b00 ^ (~b10 & b20); a10 = b10 ^ (~b20 & b30); a20 = b20 ^ (~b30 & b40); a30 = b30 ^ (~b40 & b00); a40 = b40 ^ (~b00 & b10);
    a01 = b01 ^ (~b11 & b21); a11 = b11 ^ (~b21 & b31); a21 = b21 ^ (~b31 & b41); a31 = b31 ^ (~b41 & b01); a41 = b41 ^ (~b01 & b11);
    a02 = b02 ^ (~b12 & b22); a12 = b12 ^ (~b22 & b32); a22 = b22 ^ (~b32 & b42); a32 = b32 ^ (~b42 & b02); a42 = b42 ^ (~b02 & b12);
    a03 = b03 ^ (~b13 & b23); a13 = b13 ^ (~b23 & b33); a23 = b23 ^ (~b33 & b43); a33 = b33 ^ (~b43 & b03); a43 = b43 ^ (~b03 & b13);
    a04 = b04 ^ (~b14 & b24); a14 = b14 ^ (~b24 & b34); a24 = b24 ^ (~b34 & b44); a34 = b34 ^ (~b44 & b04); a44 = b44 ^ (~b04 & b14);
    a00 ^= RC[rnd];
  }
  a[0] = a00; a[1] = a10; a[2] = a20; a[3] = a30; a[4] = a40;
  a[5] = a01; a[6] = a11; a[7] = a21; a[8] = a31; a[9] = a41;
  a[10] = a02; a[11] = a12; a[12] = a22; a[13] = a32; a[14] = a42;
  a[15] = a03; a[16] = a13; a[17] = a23; a[18] = a33; a[19] = a43;
  a[20] = a04; a[21] = a14; a[22] = a24; a[23] = a34; a[24] = a44;
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
void Strobe128::export_state(uint32_t out[53]) const {
  for (int i = 0; i < 25; i++) {
    out[2 * i] = (uint32_t)st_.lanes[i];
    out[2 * i + 1] = (uint32_t)(st_.lanes[i] >> 32);
  }
  out[50] = pos_;
  out[51] = pos_begin_;
  out[52] = cur_flags_;
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

// ZKP_DEBUG_TRANSCRIPT=1 in the environment: everything fed to (and drawn from) a host transcript goes to stderr, like
// the reference's `debug-transcript` feature (Cargo.toml:35 -> merlin/debug-transcript, README.md:76-79).  For comparing
// transcripts with another implementation byte by byte; it prints challenges, so never in production.
static bool debug_transcript() {
  static const bool on = [] {
    const char* e = getenv("ZKP_DEBUG_TRANSCRIPT");
    return e && *e && *e != '0';
  }();
  return on;
}
static void debug_line(const char* op, const uint8_t* label, size_t llen, const uint8_t* data, size_t dlen) {
  fprintf(stderr, "[zkp transcript] %s label=\"", op);
  for (size_t i = 0; i < llen; i++) fputc(label[i] >= 32 && label[i] < 127 ? label[i] : '.', stderr);
  fprintf(stderr, "\" len=%zu data=", dlen);
  for (size_t i = 0; i < dlen; i++) fprintf(stderr, "%02x", data[i]);
  fputc('\n', stderr);
}

Transcript::Transcript(const uint8_t* label, size_t len) : strobe_((const uint8_t*)"Merlin v1.0", 11) {
  append_message((const uint8_t*)"dom-sep", 7, label, len);
}
void Transcript::append_message(const uint8_t* label, size_t llen, const uint8_t* msg, size_t mlen) {
  if (debug_transcript()) debug_line("append_message", label, llen, msg, mlen);
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
  if (debug_transcript()) debug_line("challenge_bytes", label, llen, dest, dlen);
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
