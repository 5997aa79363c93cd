// Host-side value types of the reference's public API: FieldElement and G1 (amcl_wrapper types as the
// reference uses them, /root/reference/src/ipp.rs:13-20, src/r1cs/proof.rs:26-58).
//
// FieldElement is a residue mod r kept in Montgomery form on 64-bit limbs (csrc/host_fp.h, same radix as
// the device).  G1 on the host is only ever a *value to hash or ship*: its canonical affine bytes
// X||Y (MODBYTES big endian each; identity = AMCL's (0,1)).  Every group operation runs on the device.
#pragma once
#include <stdint.h>
#include <string.h>

#include <random>
#include <vector>

#include "../../include/bpgpu.h"
#include "../csrc/host_fp.h"
#include "merlin.hpp"

namespace bph {

struct Bls381 {
  using FrP = bp::BlsFr;
  static constexpr int ID = BPGPU_BLS12_381;
  static constexpr int MODBYTES = 48;
};
struct Bn254 {
  using FrP = bp::BnFr;
  static constexpr int ID = BPGPU_BN254;
  static constexpr int MODBYTES = 32;
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
// mode 0: OS entropy expanded through SHAKE256; mode 1: the deterministic stream the oracle and the tests use,
// scalar_i = SHAKE256(seed_le64 || tag || i_le64)[..MODBYTES] mod r  (oracle/r1cs.py make_rng).
template <class C>
class Rng {
 public:
  Rng() : deterministic_(false), seed_(0), tag_("os") {
    std::random_device rd;
    seed_ = ((uint64_t)rd() << 32) ^ rd();
    for (int i = 0; i < 8; i++) os_key_[i] = rd();
  }
  Rng(uint64_t seed, const std::string& tag) : deterministic_(true), seed_(seed), tag_(tag) {}
  FieldElement<C> next() {
    std::vector<uint8_t> m;
    for (int i = 0; i < 8; i++) m.push_back((uint8_t)(seed_ >> (8 * i)));
    m.insert(m.end(), tag_.begin(), tag_.end());
    for (int i = 0; i < 8; i++) m.push_back((uint8_t)(ctr_ >> (8 * i)));
    if (!deterministic_) m.insert(m.end(), (const uint8_t*)os_key_, (const uint8_t*)os_key_ + sizeof os_key_);
    ctr_++;
    uint8_t out[C::MODBYTES];
    shake256(m.data(), m.size(), out, sizeof out);
    return FieldElement<C>::from_bytes(out);
  }

 private:
  bool deterministic_;
  uint64_t seed_;
  std::string tag_;
  uint64_t ctr_ = 0;
  uint32_t os_key_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

}  // namespace bph
