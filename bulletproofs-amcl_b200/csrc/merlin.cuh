// Merlin v1.0 transcripts over STROBE-128 / Keccak-f[1600] as host+device code.
//
// The reference drives `merlin::Transcript` through its TranscriptProtocol trait
// (/root/reference/src/transcript.rs:12-61).  For SINGLE proofs the transcript stays on the host
// (host/merlin.hpp).  For SLABS of independent proofs (batch verification and batch
// proving) the B transcripts are B independent Keccak streams: one thread per proof replays them where the proof
// bytes already are, and only challenges leave the kernel.  Same framing as host/merlin.hpp (STROBE-128/1.0.2,
// rate 166, Merlin's meta-AD/AD/PRF operations); pinned by Merlin's `equivalence_simple` known answer through the
// host build of this header (tests/host_check2.cpp).
#pragma once
#include "fp.cuh"

namespace bp {

BP_HD uint64_t mrl_rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

// Keccak-f[1600], 24 rounds; one round unrolled, the round loop kept rolled (code size: this runs next to field code)
static BP_HD_COLD void keccak_f1600_hd(uint64_t* a) {
  const uint64_t RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
                           0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
                           0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
                           0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
                           0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
                           0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  uint64_t s[25];
#pragma unroll
  for (int i = 0; i < 25; i++) s[i] = a[i];
#pragma unroll 1
  for (int r = 0; r < 24; r++) {
    uint64_t c[5], b[25];
#pragma unroll
    for (int x = 0; x < 5; x++) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
#pragma unroll
    for (int x = 0; x < 5; x++) {
      const uint64_t d = c[(x + 4) % 5] ^ mrl_rotl64(c[(x + 1) % 5], 1);
#pragma unroll
      for (int y = 0; y < 5; y++) s[x + 5 * y] ^= d;
    }
    // rho + pi: B[y + 5*((2x + 3y) % 5)] = rot(A[x + 5y], r[x + 5y])
    b[0] = s[0];
    b[10] = mrl_rotl64(s[1], 1);   b[20] = mrl_rotl64(s[2], 62);  b[5] = mrl_rotl64(s[3], 28);   b[15] = mrl_rotl64(s[4], 27);
    b[16] = mrl_rotl64(s[5], 36);  b[1] = mrl_rotl64(s[6], 44);   b[11] = mrl_rotl64(s[7], 6);   b[21] = mrl_rotl64(s[8], 55);
    b[6] = mrl_rotl64(s[9], 20);   b[7] = mrl_rotl64(s[10], 3);   b[17] = mrl_rotl64(s[11], 10); b[2] = mrl_rotl64(s[12], 43);
    b[12] = mrl_rotl64(s[13], 25); b[22] = mrl_rotl64(s[14], 39); b[23] = mrl_rotl64(s[15], 41); b[8] = mrl_rotl64(s[16], 45);
    b[18] = mrl_rotl64(s[17], 15); b[3] = mrl_rotl64(s[18], 21);  b[13] = mrl_rotl64(s[19], 8);  b[14] = mrl_rotl64(s[20], 18);
    b[24] = mrl_rotl64(s[21], 2);  b[9] = mrl_rotl64(s[22], 61);  b[19] = mrl_rotl64(s[23], 56); b[4] = mrl_rotl64(s[24], 14);
#pragma unroll
    for (int y = 0; y < 5; y++)
#pragma unroll
      for (int x = 0; x < 5; x++) s[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    s[0] ^= RC[r];
  }
#pragma unroll
  for (int i = 0; i < 25; i++) a[i] = s[i];
}

// exported STROBE state: 200 state bytes (little-endian lanes) | pos | pos_begin | cur_flags  (host/merlin.hpp export_state)
static const int MERLIN_STATE_BYTES = 203;

// STROBE-128 duplex restricted to the three operations Merlin uses
struct StrobeHD {
  static constexpr uint32_t R = 166;
  static constexpr uint8_t FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_M = 16;
  uint64_t st[25];
  uint32_t pos, pos_begin;

  BP_HD void xor_byte(uint32_t i, uint8_t v) { st[i >> 3] ^= (uint64_t)v << (8 * (i & 7)); }
  BP_HD uint8_t get_byte(uint32_t i) const { return (uint8_t)(st[i >> 3] >> (8 * (i & 7))); }
  BP_HD void clr_byte(uint32_t i) { st[i >> 3] &= ~((uint64_t)0xff << (8 * (i & 7))); }

  BP_HD void load(const uint8_t* in) {
    for (int l = 0; l < 25; l++) {
      uint64_t v = 0;
      for (int k = 7; k >= 0; k--) v = (v << 8) | in[8 * l + k];
      st[l] = v;
    }
    pos = in[200]; pos_begin = in[201];
  }
  BP_HD void store(uint8_t* out) const {
    for (int l = 0; l < 25; l++)
      for (int k = 0; k < 8; k++) out[8 * l + k] = (uint8_t)(st[l] >> (8 * k));
    out[200] = (uint8_t)pos; out[201] = (uint8_t)pos_begin; out[202] = 0;
  }
  BP_HD void run_f() {
    xor_byte(pos, (uint8_t)pos_begin);
    xor_byte(pos + 1, 0x04);
    xor_byte(R + 1, 0x80);
    keccak_f1600_hd(st);
    pos = 0; pos_begin = 0;
  }
  BP_HD void absorb1(uint8_t b) { xor_byte(pos, b); if (++pos == R) run_f(); }
  BP_HD void absorb(const uint8_t* d, uint32_t n) { for (uint32_t i = 0; i < n; i++) absorb1(d[i]); }
  BP_HD void squeeze(uint8_t* out, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) { out[i] = get_byte(pos); clr_byte(pos); if (++pos == R) run_f(); }
  }
  BP_HD void begin_op(uint8_t flags) {
    const uint8_t old_begin = (uint8_t)pos_begin;
    pos_begin = pos + 1;
    absorb1(old_begin);
    absorb1(flags);
    if ((flags & FLAG_C) && pos != 0) run_f();
  }
  // Merlin framing: meta-AD(label) || meta-AD(len as u32 LE, more) then AD(message) or PRF(n)
  BP_HD void frame(const char* label, uint32_t label_len, uint32_t n) {
    begin_op(FLAG_M | FLAG_A);
    for (uint32_t i = 0; i < label_len; i++) absorb1((uint8_t)label[i]);
    absorb1((uint8_t)n); absorb1((uint8_t)(n >> 8)); absorb1((uint8_t)(n >> 16)); absorb1((uint8_t)(n >> 24));
  }
  // Transcript::append_message(label, msg) with the message given as an optional one-byte prefix (the 0x04 tag of
  // G1::to_bytes) followed by n bytes
  BP_HD void append_message(const char* label, uint32_t label_len, const uint8_t* msg, uint32_t n, int tag = -1) {
    frame(label, label_len, n + (tag >= 0 ? 1u : 0u));
    begin_op(FLAG_A);
    if (tag >= 0) absorb1((uint8_t)tag);
    absorb(msg, n);
  }
  BP_HD void append_u64(const char* label, uint32_t label_len, uint64_t x) {
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
    append_message(label, label_len, b, 8);
  }
  BP_HD void challenge_bytes(const char* label, uint32_t label_len, uint8_t* out, uint32_t n) {
    frame(label, label_len, n);
    begin_op(FLAG_I | FLAG_A | FLAG_C);
    squeeze(out, n);
  }
};

#define MRL_LIT(s) (s), (uint32_t)(sizeof(s) - 1)

}  // namespace bp
