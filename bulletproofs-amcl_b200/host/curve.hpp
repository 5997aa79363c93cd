// Host-side value types of the reference's public API: FieldElement and G1 (amcl_wrapper types as the
// reference uses them, /root/reference/src/ipp.rs:13-20, src/r1cs/proof.rs:26-58).
//
// FieldElement is a residue mod r kept in Montgomery form on 64-bit limbs (csrc/host_fp.h, same radix as
// the device).  G1 on the host is only ever a *value to hash or ship*: its canonical affine bytes
// X||Y (MODBYTES big endian each; identity = AMCL's (0,1)).  Every group operation runs on the device.
#pragma once
#include <stdint.h>
#include <string.h>

#include <stdio.h>
#include <stdlib.h>

#include <sys/random.h>

#include <chrono>
#include <string>
#include <vector>

#include "../../include/bpgpu.h"
#include "../csrc/host_fp.h"
#include "merlin.hpp"

namespace bph {

struct Bls381 {
  using FrP = bp::BlsFr;
  using FqP = bp::BlsFq;
  static constexpr unsigned CURVE_B = 4;
  static constexpr int ID = BPGPU_BLS12_381;
  static constexpr int MODBYTES = 48;
};
struct Bn254 {
  using FrP = bp::BnFr;
  using FqP = bp::BnFq;
  static constexpr unsigned CURVE_B = 2;
  static constexpr int ID = BPGPU_BN254;
  static constexpr int MODBYTES = 32;
};

// BPH_TRACE=1 in the environment prints the wall time between milestones of prove / verify to stderr
struct Trace {
  bool on;
  const char* what;
  std::chrono::steady_clock::time_point t0, last;
  explicit Trace(const char* w) : on(getenv("BPH_TRACE") != nullptr), what(w) { t0 = last = std::chrono::steady_clock::now(); }
  void mark(const char* name) {
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[bph %s] %-22s %8.3f ms\n", what, name, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
  }
  ~Trace() {
    if (on) fprintf(stderr, "[bph %s] total %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
};

// R1CSError (errors.rs:7-28) as the C ABI's status codes
enum : int {
  OK = BPGPU_OK,
  E_LEN = BPGPU_E_LEN,
  E_NOT_POW2 = BPGPU_E_NOT_POW2,
  E_INVALID_GENERATORS_LENGTH = BPGPU_E_GENS_LEN,
  E_VERIFICATION = BPGPU_E_VERIFY,
  E_FORMAT = BPGPU_E_FORMAT,
  E_CUDA = BPGPU_E_CUDA,
  E_ARG = BPGPU_E_ARG,
  E_MISSING_ASSIGNMENT = -8,
  E_GADGET = -9,
  E_ENTROPY = -11,
};

template <class C>
struct FieldElement {
  using H = bp::host::HFp<typename C::FrP>;
  static constexpr int MB = C::MODBYTES;
  H v;

  static FieldElement zero() { return {H::zero()}; }
  static FieldElement one() { return {H::one()}; }
  static FieldElement minus_one() { return {H::zero() - H::one()}; }
  static FieldElement from_u64(uint64_t x) { return {H::from_u64(x)}; }
  // FieldElement::from(&[u8; MODBYTES]): big-endian integer reduced mod r (transcript.rs:55-60).
  // On BLS12-381 the 48-byte value can exceed 2^256: value = hi * 2^256 + lo.
  static FieldElement from_bytes(const uint8_t* be) {
    H lo = H::from_be(be, MB);                       // low 32 bytes, reduced
    if (MB == 32) return {lo};
    uint8_t hb[32];
    memset(hb, 0, sizeof hb);
    memcpy(hb + 32 - (MB - 32), be, MB - 32);
    H hi = H::from_be(hb, 32);                       // Montgomery form of hi
    return {hi.to_mont() + lo};                      // (hi * 2^256) + lo, 2^256 being the Montgomery radix
  }
  void to_bytes(uint8_t* out) const { v.to_be(out, MB); }
  std::vector<uint8_t> to_bytes() const { std::vector<uint8_t> o(MB); v.to_be(o.data(), MB); return o; }

  bool is_zero() const { return v.is_zero(); }
  bool operator==(const FieldElement& o) const { return v == o.v; }
  bool operator!=(const FieldElement& o) const { return !(v == o.v); }
  friend FieldElement operator+(const FieldElement& a, const FieldElement& b) { return {a.v + b.v}; }
  friend FieldElement operator-(const FieldElement& a, const FieldElement& b) { return {a.v - b.v}; }
  friend FieldElement operator*(const FieldElement& a, const FieldElement& b) { return {a.v * b.v}; }
  FieldElement negation() const { return {v.neg()}; }
  FieldElement square() const { return {v.sqr()}; }
  FieldElement inverse() const { return {v.inv()}; }   // inverse of 0 is 0, as in AMCL
};

template <class C>
inline void append_fe(std::vector<uint8_t>& buf, const FieldElement<C>& x) {
  size_t o = buf.size();
  buf.resize(o + C::MODBYTES);
  x.to_bytes(buf.data() + o);
}

template <class C>
struct G1 {
  static constexpr int MB = C::MODBYTES;
  uint8_t xy[2 * C::MODBYTES];
  static G1 identity() { G1 p; memset(p.xy, 0, sizeof p.xy); p.xy[2 * MB - 1] = 1; return p; }
  static G1 from_xy(const uint8_t* b) { G1 p; memcpy(p.xy, b, sizeof p.xy); return p; }
  bool is_identity() const { return *this == identity(); }
  // a point as AMCL's ECP::frombytes accepts it (coordinates < p, on the curve, or (0, 1)): checked for every point that
  // arrives from outside before it is hashed into a transcript or used in an MSM
  bool is_valid() const { return bp::host::g1_xy_is_valid<typename C::FqP>(xy, MB, C::CURVE_B); }
  static bool xy_valid(const uint8_t* b) { return bp::host::g1_xy_is_valid<typename C::FqP>(b, MB, C::CURVE_B); }
  bool operator==(const G1& o) const { return memcmp(xy, o.xy, sizeof xy) == 0; }
  bool operator!=(const G1& o) const { return !(*this == o); }
  // G1::to_bytes(): 0x04 || X || Y
  std::vector<uint8_t> to_bytes() const { std::vector<uint8_t> o(1 + sizeof xy); o[0] = 4; memcpy(o.data() + 1, xy, sizeof xy); return o; }
};

// TranscriptProtocol (transcript.rs:12-61) over the host Merlin implementation
template <class C>
struct TranscriptProtocol {
  static void commit_scalar(Transcript& t, const char* label, const FieldElement<C>& s) { t.append_message(label, s.to_bytes()); }
  static void commit_point(Transcript& t, const char* label, const G1<C>& p) { t.append_message(label, p.to_bytes()); }
  static FieldElement<C> challenge_scalar(Transcript& t, const char* label) {
    uint8_t buf[C::MODBYTES];
    t.challenge_bytes(label, buf, sizeof buf);
    return FieldElement<C>::from_bytes(buf);
  }
};

// Source of the prover's blinding scalars and the verifier's batching scalar
// (FieldElement::random(): prover.rs:336-341,389-402,490-494; verifier.rs:392).
// One counter-mode stream in both modes: scalar_i = be_int(SHAKE256(key || le64(i))[..MODBYTES]) mod r.
//   mode 1 (tests, oracle/r1cs.py make_rng): key = seed_le64 || tag -- the deterministic stream of the parity tests;
//   mode 0: key = 32 bytes of OS entropy || "os".
// Counter mode lets the long draws (the blinding VECTORS s_L, s_R) be generated on the device, element i by thread i
// (fill_device -> bpgpu_fr_random), at the place in the stream where the same number of next() calls would have been.
// best-effort wipe of secret host memory (keys, blindings, witness) that the optimiser may not drop
inline void secure_zero(void* p, size_t n) {
  volatile uint8_t* v = static_cast<volatile uint8_t*>(p);
  for (size_t i = 0; i < n; i++) v[i] = 0;
}

template <class C>
class Rng {
 public:
  // OS entropy through getrandom(2); ok() is false if the kernel could not supply it (callers return BPH_E_ENTROPY:
  // no blinding is ever derived from a partially filled key)
  Rng() {
    size_t got = 0;
    while (got < 32) {
      ssize_t r = getrandom(key_ + got, 32 - got, 0);
      if (r <= 0) { ok_ = false; break; }
      got += (size_t)r;
    }
    key_[32] = 'o'; key_[33] = 's';
    key_len_ = 34;
  }
  Rng(uint64_t seed, const std::string& tag) {
    for (int i = 0; i < 8; i++) key_[key_len_++] = (uint8_t)(seed >> (8 * i));
    for (char ch : tag) if (key_len_ < 48) key_[key_len_++] = (uint8_t)ch;
  }
  ~Rng() { secure_zero(key_, sizeof key_); }
  bool ok() const { return ok_; }
  FieldElement<C> next() {
    uint8_t m[64], out[C::MODBYTES];
    memcpy(m, key_, key_len_);
    for (int i = 0; i < 8; i++) m[key_len_ + i] = (uint8_t)(ctr_ >> (8 * i));
    ctr_++;
    shake256(m, key_len_ + 8, out, sizeof out);
    return FieldElement<C>::from_bytes(out);
  }
  // stream position, for callers that generate a run of draws elsewhere (the batched prover's device kernels)
  const uint8_t* key() const { return key_; }
  size_t key_len() const { return key_len_; }
  uint64_t counter() const { return ctr_; }
  void skip(uint64_t n) { ctr_ += n; }
  // the next n draws as a device-resident vector (FieldElementVector::random); *out is owned by the caller
  int fill_device(bpgpu_ctx* ctx, size_t n, bpgpu_scalars** out) {
    int rc = bpgpu_fr_random(ctx, key_, key_len_, ctr_, n, out);
    if (!rc) ctr_ += n;
    return rc;
  }

 private:
  uint8_t key_[56] = {0};
  size_t key_len_ = 0;
  uint64_t ctr_ = 0;
  bool ok_ = true;
};

}  // namespace bph
