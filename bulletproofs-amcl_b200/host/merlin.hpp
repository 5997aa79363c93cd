// Merlin v1.0 transcripts over STROBE-128 / Keccak-f[1600], host side (C++).
//
// The reference drives `merlin::Transcript` through its TranscriptProtocol trait
// (/root/reference/src/transcript.rs:12-61); per the design the transcript and challenge derivation stay on
// the host.  The merlin crate is not vendored, so this restates the published STROBE-128/1.0.2 framing
//  and is pinned by Merlin's `equivalence_simple` known answer (tests/test_host_lib.py).
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

namespace bph {

inline uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

// Keccak-f[1600], 24 rounds.  State lanes live in locals and every loop has constant bounds, so the compiler keeps
// the 25 lanes in registers and unrolls theta / rho-pi / chi completely.
inline void keccak_f1600(uint64_t st[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
      0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
      0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
      0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  // rotation offsets r[x + 5y] and the pi permutation: B[y + 5*((2x + 3y) % 5)] = rot(A[x + 5y])
  constexpr int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  uint64_t a[25];
#pragma GCC unroll 25
  for (int i = 0; i < 25; i++) a[i] = st[i];
  for (int r = 0; r < 24; r++) {
    uint64_t c[5], d[5], b[25];
#pragma GCC unroll 5
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma GCC unroll 5
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
#pragma GCC unroll 25
    for (int i = 0; i < 25; i++) {
      const int x = i % 5, y = i / 5;
      b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(a[i] ^ d[x], ROT[i]);
    }
#pragma GCC unroll 25
    for (int i = 0; i < 25; i++) {
      const int x = i % 5, y5 = i - x;
      a[i] = b[i] ^ (~b[y5 + (x + 1) % 5] & b[y5 + (x + 2) % 5]);
    }
    a[0] ^= RC[r];
  }
#pragma GCC unroll 25
  for (int i = 0; i < 25; i++) st[i] = a[i];
}

// SHAKE256 (used for the deterministic synthetic blinding streams of tests/bench)
inline void shake256(const uint8_t* msg, size_t len, uint8_t* out, size_t outlen) {
  const size_t rate = 136;
  uint64_t st[25];
  memset(st, 0, sizeof st);
  uint8_t* b = reinterpret_cast<uint8_t*>(st);
  size_t pos = 0;
  for (size_t i = 0; i < len; i++) { b[pos++] ^= msg[i]; if (pos == rate) { keccak_f1600(st); pos = 0; } }
  b[pos] ^= 0x1f; b[rate - 1] ^= 0x80;
  keccak_f1600(st);
  pos = 0;
  for (size_t i = 0; i < outlen; i++) { if (pos == rate) { keccak_f1600(st); pos = 0; } out[i] = b[pos++]; }
}

class Strobe128 {
 public:
  explicit Strobe128(const char* protocol_label) {
    memset(st_, 0, sizeof st_);
    uint8_t* b = bytes();
    const uint8_t init[6] = {1, R + 2, 1, 0, 1, 96};
    memcpy(b, init, 6);
    memcpy(b + 6, "STROBEv1.0.2", 12);
    keccak_f1600(st_);
    meta_ad(reinterpret_cast<const uint8_t*>(protocol_label), strlen(protocol_label), false);
  }
  void meta_ad(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_M | FLAG_A, more); absorb(d, n); }
  void ad(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_A, more); absorb(d, n); }
  void prf(uint8_t* out, size_t n, bool more) { begin_op(FLAG_I | FLAG_A | FLAG_C, more); squeeze(out, n); }
  // 200 state bytes | pos | pos_begin | cur_flags: where the device-side replay of a slab's transcripts starts
  // (csrc/merlin.cuh StrobeHD::load)
  void export_state(uint8_t out[203]) const {
    memcpy(out, st_, 200);
    out[200] = pos_; out[201] = pos_begin_; out[202] = cur_flags_;
  }

 private:
  static constexpr uint8_t R = 166;
  static constexpr uint8_t FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32;
  uint64_t st_[25];
  uint8_t pos_ = 0, pos_begin_ = 0, cur_flags_ = 0;
  uint8_t* bytes() { return reinterpret_cast<uint8_t*>(st_); }
  void run_f() {
    uint8_t* b = bytes();
    b[pos_] ^= pos_begin_; b[pos_ + 1] ^= 0x04; b[R + 1] ^= 0x80;
    keccak_f1600(st_);
    pos_ = 0; pos_begin_ = 0;
  }
  void absorb(const uint8_t* d, size_t n) {
    uint8_t* b = bytes();
    for (size_t i = 0; i < n; i++) { b[pos_++] ^= d[i]; if (pos_ == R) run_f(); }
  }
  void squeeze(uint8_t* out, size_t n) {
    uint8_t* b = bytes();
    for (size_t i = 0; i < n; i++) { out[i] = b[pos_]; b[pos_] = 0; pos_++; if (pos_ == R) run_f(); }
  }
  void begin_op(uint8_t flags, bool more) {
    if (more) return;                    // continuation of the current operation (same flags)
    uint8_t old_begin = pos_begin_;
    pos_begin_ = pos_ + 1;
    cur_flags_ = flags;
    uint8_t hdr[2] = {old_begin, flags};
    absorb(hdr, 2);
    if ((flags & (FLAG_C | FLAG_K)) && pos_ != 0) run_f();
  }
};

// merlin::Transcript (v1.0 framing)
class Transcript {
 public:
  explicit Transcript(const std::string& label) : strobe_("Merlin v1.0") { append_message("dom-sep", label); }
  void append_message(const char* label, const uint8_t* msg, size_t len) {
    strobe_.meta_ad(reinterpret_cast<const uint8_t*>(label), strlen(label), false);
    uint8_t l4[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
    strobe_.meta_ad(l4, 4, true);
    strobe_.ad(msg, len, false);
  }
  void append_message(const char* label, const std::string& s) { append_message(label, reinterpret_cast<const uint8_t*>(s.data()), s.size()); }
  void append_message(const char* label, const std::vector<uint8_t>& v) { append_message(label, v.data(), v.size()); }
  void append_u64(const char* label, uint64_t x) {
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
    append_message(label, b, 8);
  }
  void challenge_bytes(const char* label, uint8_t* out, size_t n) {
    strobe_.meta_ad(reinterpret_cast<const uint8_t*>(label), strlen(label), false);
    uint8_t l4[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_.meta_ad(l4, 4, true);
    strobe_.prf(out, n, false);
  }
  // TranscriptProtocol domain separators (transcript.rs:30-45)
  void innerproduct_domain_sep(uint64_t n) { append_message("dom-sep", std::string("ipp v1")); append_u64("n", n); }
  void r1cs_domain_sep() { append_message("dom-sep", std::string("r1cs v1")); }
  void r1cs_1phase_domain_sep() { append_message("dom-sep", std::string("r1cs-1phase")); }
  void r1cs_2phase_domain_sep() { append_message("dom-sep", std::string("r1cs-2phase")); }
  void export_state(uint8_t out[203]) const { strobe_.export_state(out); }

 private:
  Strobe128 strobe_;
};

}  // namespace bph
