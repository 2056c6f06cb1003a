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


// Three-input logic on 32-bit halves, spelled out: the front-end kernels are bound by the ALU pipe (profiles/
// r02_front_end_ncu.md) and a Keccak round written on 64-bit lanes compiled to 156 LOP3 + 27 register moves where
// 122 LOP3 do -- theta's D = C[x-1] ^ rol(C[x+1], 1) is never formed: every lane takes A ^ C[x-1] ^ rol(C[x+1], 1) in ONE
// LOP3 per half (the compiler, left alone, factors D out again), and halves that are not register pairs need no moves.
#if ZKP_DEVICE_ASM
ZKP_DEV uint32_t kx3(uint32_t a, uint32_t b, uint32_t c) {      // a ^ b ^ c
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
ZKP_DEV uint32_t kchi(uint32_t a, uint32_t b, uint32_t c) {     // a ^ (~b & c)
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xD2;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
ZKP_DEV uint32_t kfsl(uint32_t lo, uint32_t hi, int s) { return __funnelshift_l(lo, hi, s); }   // high word of (hi:lo) << s
#else
ZKP_DEV uint32_t kx3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }
ZKP_DEV uint32_t kchi(uint32_t a, uint32_t b, uint32_t c) { return a ^ (~b & c); }
ZKP_DEV uint32_t kfsl(uint32_t lo, uint32_t hi, int s) { return (hi << s) | (lo >> (32 - s)); }
#endif

ZKP_DEV void keccak_f1600_dev(uint64_t* a) {
  const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
      0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
      0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
      0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  // rho offsets of lane (x, y) at index x + 5 y; pi sends lane (x, y) to (y, 2x + 3y)
  const int RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  uint32_t l[25], h[25];
#pragma unroll
  for (int i = 0; i < 25; i++) { l[i] = (uint32_t)a[i]; h[i] = (uint32_t)(a[i] >> 32); }
#if ZKP_DEVICE_ASM
#pragma unroll 1
#endif
  for (int rnd = 0; rnd < 24; rnd++) {
    uint32_t cl[5], ch[5], rl[5], rh[5], bl[25], bh[25];
#pragma unroll
    for (int x = 0; x < 5; x++) {
      cl[x] = kx3(kx3(l[x], l[x + 5], l[x + 10]), l[x + 15], l[x + 20]);
      ch[x] = kx3(kx3(h[x], h[x + 5], h[x + 10]), h[x + 15], h[x + 20]);
    }
#pragma unroll
    for (int x = 0; x < 5; x++) {   // rol64(C[x], 1)
      rl[x] = kfsl(ch[x], cl[x], 1);
      rh[x] = kfsl(cl[x], ch[x], 1);
    }
#pragma unroll
    for (int y = 0; y < 5; y++) {
#pragma unroll
      for (int x = 0; x < 5; x++) {
        const int i = x + 5 * y, dst = y + 5 * ((2 * x + 3 * y) % 5), r = RHO[i];
        const uint32_t tl = kx3(l[i], cl[(x + 4) % 5], rl[(x + 1) % 5]);
        const uint32_t th = kx3(h[i], ch[(x + 4) % 5], rh[(x + 1) % 5]);
        if (r == 0) {
          bl[dst] = tl; bh[dst] = th;
        } else if (r < 32) {
          bl[dst] = kfsl(th, tl, r); bh[dst] = kfsl(tl, th, r);
        } else {   // rotation by 32 + (r - 32): the halves trade places first (no lane rotates by exactly 32)
          bl[dst] = kfsl(tl, th, r - 32); bh[dst] = kfsl(th, tl, r - 32);
        }
      }
    }
#pragma unroll
    for (int y = 0; y < 5; y++) {
#pragma unroll
      for (int x = 0; x < 5; x++) {
        const int i = x + 5 * y, i1 = (x + 1) % 5 + 5 * y, i2 = (x + 2) % 5 + 5 * y;
        l[i] = kchi(bl[i], bl[i1], bl[i2]);
        h[i] = kchi(bh[i], bh[i1], bh[i2]);
      }
    }
    l[0] ^= (uint32_t)RC[rnd];
    h[0] ^= (uint32_t)(RC[rnd] >> 32);
  }
#pragma unroll
  for (int i = 0; i < 25; i++) a[i] = (uint64_t)l[i] | ((uint64_t)h[i] << 32);
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
