// Keccak-f[1600], STROBE-128 and the Merlin transcript operations for device threads (one transcript per thread).
//
// SURVEY.md section 8f-1 ("next" row): the reference hashes every proof's transcript on the host
// (/root/reference/src/toolbox/batch_verifier.rs:75-77, :92-94, :105-107, :125-128, :152-167 through merlin ^2 [ext]);
// once the MSM runs on the GPU that hashing is the end-to-end limit, and the transcripts of a batch are
// independent.  This header restates merlin 2.0 `strobe.rs` / `transcript.rs` for the device so that
// zkp_batch_verify_proofs can derive the challenges next to the MSM.  The host mirror (csrc/host/merlin.cpp) and
// the oracle (oracle/merlin.py) implement the same byte stream; tests compare all three.
// Compiles for the host with -DZKP_HOST_EMUL like the arithmetic headers.
#pragma once
#include <stdint.h>
#include "fe.cuh"

namespace zkp {

ZKP_DEV uint64_t rol64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

ZKP_DEV void keccak_f1600_dev(uint64_t* a) {
  const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
      0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
      0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
      0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  uint64_t a00 = a[0], a10 = a[1], a20 = a[2], a30 = a[3], a40 = a[4];
  uint64_t a01 = a[5], a11 = a[6], a21 = a[7], a31 = a[8], a41 = a[9];
  uint64_t a02 = a[10], a12 = a[11], a22 = a[12], a32 = a[13], a42 = a[14];
  uint64_t a03 = a[15], a13 = a[16], a23 = a[17], a33 = a[18], a43 = a[19];
  uint64_t a04 = a[20], a14 = a[21], a24 = a[22], a34 = a[23], a44 = a[24];
#if ZKP_DEVICE_ASM
#pragma unroll 1
#endif
  for (int rnd = 0; rnd < 24; rnd++) {
    uint64_t c0 = a00 ^ a01 ^ a02 ^ a03 ^ a04, c1 = a10 ^ a11 ^ a12 ^ a13 ^ a14, c2 = a20 ^ a21 ^ a22 ^ a23 ^ a24;
    uint64_t c3 = a30 ^ a31 ^ a32 ^ a33 ^ a34, c4 = a40 ^ a41 ^ a42 ^ a43 ^ a44;
    uint64_t d0 = c4 ^ rol64(c1, 1), d1 = c0 ^ rol64(c2, 1), d2 = c1 ^ rol64(c3, 1), d3 = c2 ^ rol64(c4, 1),
             d4 = c3 ^ rol64(c0, 1);
    uint64_t b00 = a00 ^ d0, b13 = rol64(a01 ^ d0, 36), b21 = rol64(a02 ^ d0, 3), b34 = rol64(a03 ^ d0, 41),
             b42 = rol64(a04 ^ d0, 18);
    uint64_t b02 = rol64(a10 ^ d1, 1), b10 = rol64(a11 ^ d1, 44), b23 = rol64(a12 ^ d1, 10), b31 = rol64(a13 ^ d1, 45),
             b44 = rol64(a14 ^ d1, 2);
    uint64_t b04 = rol64(a20 ^ d2, 62), b12 = rol64(a21 ^ d2, 6), b20 = rol64(a22 ^ d2, 43), b33 = rol64(a23 ^ d2, 15),
             b41 = rol64(a24 ^ d2, 61);
    uint64_t b01 = rol64(a30 ^ d3, 28), b14 = rol64(a31 ^ d3, 55), b22 = rol64(a32 ^ d3, 25), b30 = rol64(a33 ^ d3, 21),
             b43 = rol64(a34 ^ d3, 56);
    uint64_t b03 = rol64(a40 ^ d4, 27), b11 = rol64(a41 ^ d4, 20), b24 = rol64(a42 ^ d4, 39), b32 = rol64(a43 ^ d4, 8),
             b40 = rol64(a44 ^ d4, 14);
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

// STROBE-128 state of one transcript: 200 state bytes as 25 lanes + (pos, pos_begin, cur_flags)
struct strobe_t {
  uint64_t st[25];
  uint32_t pos, pos_begin, cur_flags;
};

// strobe_run_f is compiled out of line on the device (code size: ~3.6k instructions, several call sites)
#if ZKP_DEVICE_ASM
#define ZKP_DEV_NOINLINE __device__ __noinline__
#else
#define ZKP_DEV_NOINLINE static
#endif

#define ZKP_STROBE_R 166u
#define ZKP_FLAG_I 1u
#define ZKP_FLAG_A 2u
#define ZKP_FLAG_C 4u
#define ZKP_FLAG_M 16u
#define ZKP_FLAG_K 32u

// Byte view of the state (little-endian lanes; char aliasing).  STROBE is byte-oriented and the rate (166) is not a
// multiple of 8, so the state is addressed as bytes in local memory; an earlier variant that assembled bytes into the
// lanes with variable 64-bit shifts produced wrong states on the device (though not in host emulation) and was dropped.
// tests/test_gpu_toolbox.py::test_device_merlin_selftest guards this code with Merlin's conformance vector.
ZKP_DEV uint8_t* st_bytes(strobe_t& s) { return (uint8_t*)s.st; }

// Kept out of line on the device: the permutation is ~3.6k instructions reached from several call sites.
ZKP_DEV_NOINLINE void strobe_run_f(strobe_t& s) {
  uint8_t* b = st_bytes(s);
  b[s.pos] ^= (uint8_t)s.pos_begin;
  b[s.pos + 1] ^= 0x04;
  b[ZKP_STROBE_R + 1] ^= 0x80;
  keccak_f1600_dev(s.st);
  s.pos = 0;
  s.pos_begin = 0;
}
ZKP_DEV void strobe_absorb(strobe_t& s, const uint8_t* d, uint32_t n) {
  uint8_t* b = st_bytes(s);
  for (uint32_t i = 0; i < n; i++) {
    b[s.pos] ^= d[i];
    if (++s.pos == ZKP_STROBE_R) strobe_run_f(s);
  }
}
ZKP_DEV void strobe_squeeze(strobe_t& s, uint8_t* out, uint32_t n) {
  uint8_t* b = st_bytes(s);
  for (uint32_t i = 0; i < n; i++) {
    out[i] = b[s.pos];
    b[s.pos] = 0;
    if (++s.pos == ZKP_STROBE_R) strobe_run_f(s);
  }
}
ZKP_DEV void strobe_begin_op(strobe_t& s, uint32_t flags, bool more) {
  if (more) return;
  uint8_t hdr[2];
  hdr[0] = (uint8_t)s.pos_begin;
  hdr[1] = (uint8_t)flags;
  s.pos_begin = s.pos + 1;
  s.cur_flags = flags;
  strobe_absorb(s, hdr, 2);
  if ((flags & (ZKP_FLAG_C | ZKP_FLAG_K)) && s.pos != 0) strobe_run_f(s);
}
ZKP_DEV void strobe_meta_ad(strobe_t& s, const uint8_t* d, uint32_t n, bool more) {
  strobe_begin_op(s, ZKP_FLAG_M | ZKP_FLAG_A, more);
  strobe_absorb(s, d, n);
}
ZKP_DEV void strobe_ad(strobe_t& s, const uint8_t* d, uint32_t n, bool more) {
  strobe_begin_op(s, ZKP_FLAG_A, more);
  strobe_absorb(s, d, n);
}
ZKP_DEV void strobe_prf(strobe_t& s, uint8_t* out, uint32_t n, bool more) {
  strobe_begin_op(s, ZKP_FLAG_I | ZKP_FLAG_A | ZKP_FLAG_C, more);
  strobe_squeeze(s, out, n);
}

// merlin Transcript::append_message / challenge_bytes on a strobe state
ZKP_DEV void transcript_append(strobe_t& s, const uint8_t* label, uint32_t llen, const uint8_t* msg, uint32_t mlen) {
  uint8_t l4[4];
  l4[0] = (uint8_t)mlen; l4[1] = (uint8_t)(mlen >> 8); l4[2] = (uint8_t)(mlen >> 16); l4[3] = (uint8_t)(mlen >> 24);
  strobe_meta_ad(s, label, llen, false);
  strobe_meta_ad(s, l4, 4, true);
  strobe_ad(s, msg, mlen, false);
}
ZKP_DEV void transcript_challenge(strobe_t& s, const uint8_t* label, uint32_t llen, uint8_t* out, uint32_t n) {
  uint8_t l4[4];
  l4[0] = (uint8_t)n; l4[1] = (uint8_t)(n >> 8); l4[2] = (uint8_t)(n >> 16); l4[3] = (uint8_t)(n >> 24);
  strobe_meta_ad(s, label, llen, false);
  strobe_meta_ad(s, l4, 4, true);
  strobe_prf(s, out, n, false);
}

// STROBE KEY operation (flags A|C: the state bytes are overwritten, not xored) -- merlin strobe.rs `key` [ext]
ZKP_DEV void strobe_overwrite(strobe_t& s, const uint8_t* d, uint32_t n) {
  uint8_t* b = st_bytes(s);
  for (uint32_t i = 0; i < n; i++) {
    b[s.pos] = d[i];
    if (++s.pos == ZKP_STROBE_R) strobe_run_f(s);
  }
}
ZKP_DEV void strobe_key(strobe_t& s, const uint8_t* d, uint32_t n, bool more) {
  strobe_begin_op(s, ZKP_FLAG_A | ZKP_FLAG_C, more);
  strobe_overwrite(s, d, n);
}
// merlin TranscriptRngBuilder::rekey_with_witness_bytes / finalize and TranscriptRng::fill_bytes [ext]
ZKP_DEV void rng_rekey_with_witness(strobe_t& s, const uint8_t* label, uint32_t llen, const uint8_t* w, uint32_t wlen) {
  uint8_t l4[4];
  l4[0] = (uint8_t)wlen; l4[1] = (uint8_t)(wlen >> 8); l4[2] = (uint8_t)(wlen >> 16); l4[3] = (uint8_t)(wlen >> 24);
  strobe_meta_ad(s, label, llen, false);
  strobe_meta_ad(s, l4, 4, true);
  strobe_key(s, w, wlen, false);
}
ZKP_DEV void rng_finalize(strobe_t& s, const uint8_t* entropy32) {
  const uint8_t L_RNG[3] = {'r', 'n', 'g'};
  strobe_meta_ad(s, L_RNG, 3, false);
  strobe_key(s, entropy32, 32, false);
}
ZKP_DEV void rng_fill_bytes(strobe_t& s, uint8_t* out, uint32_t n) {
  uint8_t l4[4];
  l4[0] = (uint8_t)n; l4[1] = (uint8_t)(n >> 8); l4[2] = (uint8_t)(n >> 16); l4[3] = (uint8_t)(n >> 24);
  strobe_meta_ad(s, l4, 4, false);
  strobe_prf(s, out, n, false);
}

// SHAKE-256 of a short message (< 136 bytes), first `n` output bytes: the per-proof weight generator
ZKP_DEV void shake256_short(uint8_t* out, uint32_t n, const uint8_t* msg, uint32_t len) {
  uint64_t st[25];
#pragma unroll
  for (int i = 0; i < 25; i++) st[i] = 0;
  for (uint32_t i = 0; i < len; i++) st[i >> 3] ^= (uint64_t)msg[i] << ((i & 7) * 8);
  st[len >> 3] ^= 0x1FULL << ((len & 7) * 8);
  st[16] ^= 0x80ULL << 56;  // byte 135
  keccak_f1600_dev(st);
  uint32_t pos = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (pos == 136) {
      keccak_f1600_dev(st);
      pos = 0;
    }
    out[i] = (uint8_t)(st[pos >> 3] >> ((pos & 7) * 8));
    pos++;
  }
}

}  // namespace zkp
