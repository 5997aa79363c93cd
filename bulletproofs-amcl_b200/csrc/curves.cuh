// Curve traits and the byte-level decoding of field elements and points, as host+device code (no CUDA runtime
// dependency: the host build of these headers is what the CPU unit tests of the device logic compile).
//
// Encodings are amcl_wrapper's: a scalar is FieldElement::to_bytes() = MODBYTES big endian, a point
// is G1::to_bytes()[1..] = X || Y, MODBYTES big endian each, the identity AMCL's (0, 1).
#pragma once
#include "../../include/bpgpu.h"
#include "ec.cuh"

namespace bp {

struct Bls {
  using Fq = Fp<BlsFq>;
  using Fr = Fp<BlsFr>;
  using FqParams = BlsFq;
  using FrParams = BlsFr;
  static constexpr int ID = BPGPU_BLS12_381;
  static constexpr int MODBYTES = 48;
  static constexpr int SCALAR_BITS = 255;
  static constexpr int CURVE_B = 4;            // y^2 = x^3 + 4
};
struct Bn {
  using Fq = Fp<BnFq>;
  using Fr = Fp<BnFr>;
  using FqParams = BnFq;
  using FrParams = BnFr;
  static constexpr int ID = BPGPU_BN254;
  static constexpr int MODBYTES = 32;
  static constexpr int SCALAR_BITS = 254;
  static constexpr int CURVE_B = 2;            // AMCL's (Nogami) BN254: y^2 = x^3 + 2
};

// big-endian bytes -> N little-endian limbs of the integer's low 32N bits
template <int N>
BP_HD void hd_be_to_limbs(const uint8_t* be, int nbytes, uint32_t* out) {
  for (int i = 0; i < N; i++) {
    const uint8_t* p = be + nbytes - 4 * (i + 1);
    out[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
  }
}
template <int N>
BP_HD void hd_limbs_to_be(const uint32_t* in, int nbytes, uint8_t* be) {
  for (int i = 0; i < nbytes - 4 * N; i++) be[i] = 0;
  for (int i = 0; i < N; i++) {
    uint8_t* p = be + nbytes - 4 * (i + 1);
    p[0] = (uint8_t)(in[i] >> 24); p[1] = (uint8_t)(in[i] >> 16); p[2] = (uint8_t)(in[i] >> 8); p[3] = (uint8_t)in[i];
  }
}

// limbs < modulus ?
template <class F>
BP_HD bool hd_lt_modulus(const F& x) {
  for (int i = F::N - 1; i >= 0; i--) {
    const uint32_t p = F::Params::P(i);
    if (x.v[i] < p) return true;
    if (x.v[i] > p) return false;
  }
  return false;
}

// x mod p for an arbitrary N-limb integer: both scalar fields and both base fields have 2^(32N) / p < 10
template <class F>
BP_HD void hd_canonicalise(F& x) {
#pragma unroll 1
  for (int k = 0; k < 10; k++) F::reduce_once(x.v);
}

// FieldElement::from(&[u8; MODBYTES]): the big-endian integer reduced mod r (transcript.rs:55-60), Montgomery form.
// On BLS12-381 the 48-byte value can exceed 2^256: value = hi * 2^256 + lo with 2^256 the Montgomery radix.
template <class Curve>
BP_HD typename Curve::Fr fr_from_be_wide(const uint8_t* be) {
  using Fr = typename Curve::Fr;
  constexpr int MB = Curve::MODBYTES;
  uint32_t limbs[MB / 4];
  hd_be_to_limbs<MB / 4>(be, MB, limbs);
  Fr lo;
  for (int k = 0; k < 8; k++) lo.v[k] = limbs[k];
  hd_canonicalise(lo);
  Fr res = lo * Fr::r2();
  if (MB > 32) {
    Fr hi = Fr::zero();
    bool any = false;
    for (int k = 8; k < MB / 4; k++) { hi.v[k - 8] = limbs[k]; any = any || limbs[k] != 0; }
    if (any) res = res + (hi * Fr::r2()) * Fr::r2();     // hi < 2^128 < r; hi*R*R/R = hi*R, times R2/R again = hi*2^256 in Montgomery form
  }
  return res;
}

// canonical scalar on the wire: value < r (what FieldElement::to_bytes() always emits)
template <class Curve>
BP_HD bool fr_be_is_canonical(const uint8_t* be) {
  using Fr = typename Curve::Fr;
  constexpr int MB = Curve::MODBYTES;
  for (int k = 0; k < MB - 32; k++) if (be[k]) return false;
  Fr v;
  hd_be_to_limbs<8>(be, MB, v.v);
  return hd_lt_modulus(v);
}

// Montgomery form of the curve constant b
template <class Curve>
BP_HD typename Curve::Fq curve_b_mont() {
  using Fq = typename Curve::Fq;
  Fq b = Fq::one().dbl();
  if (Curve::CURVE_B == 4) b = b.dbl();
  return b;
}

// X || Y -> affine point in Montgomery form, as AMCL's ECP::frombytes / ECP::new_bigs decide validity: a coordinate >= p
// or a pair off the curve is not a point (AMCL maps those to infinity; callers here reject them with BPGPU_E_FORMAT, see
// DESIGN.md "untrusted points").  (0, 1) is AMCL's encoding of the identity and becomes the device's (0, 0).
template <class Curve>
BP_HD bool g1_from_be_checked(const uint8_t* xy, Affine<typename Curve::Fq>* out) {
  using Fq = typename Curve::Fq;
  constexpr int MB = Curve::MODBYTES;
  Fq x, y;
  hd_be_to_limbs<Fq::N>(xy, MB, x.v);
  hd_be_to_limbs<Fq::N>(xy + MB, MB, y.v);
  if (!hd_lt_modulus(x) || !hd_lt_modulus(y)) return false;
  bool y_one = y.v[0] == 1;
  for (int k = 1; k < Fq::N; k++) y_one = y_one && y.v[k] == 0;
  if (x.is_zero() && y_one) { *out = Affine<Fq>::inf(); return true; }
  x = x.to_mont(); y = y.to_mont();
  if (y.sqr() != x.sqr() * x + curve_b_mont<Curve>()) return false;
  out->x = x; out->y = y;
  return true;
}

// ---- GLV on BLS12-381 G1: phi(x, y) = (beta * x, y) = lambda * (x, y) with lambda = z^2 - 1 (z the curve parameter),
// lambda^2 + lambda + 1 = 0 mod r.  Since r = z^4 - z^2 + 1 ~ lambda^2, plain division k = k2 * lambda + k1 already gives
// 0 <= k1 < lambda < 2^128 and k2 <= r / lambda < 2^128: k * P = k1 * P + k2 * phi(P) with two 128-bit scalars, which halves
// the doubling chain of a Straus / Horner evaluation (batch.cu).  Constants checked against the oracle in
// tests/test_host_core.py (phi(G) == lambda * G; k1 + k2 * lambda == k).
BP_HD uint32_t glv_bls_lambda(int i) { const uint32_t v[4] = {0xffffffffu, 0x00000000u, 0x0001a402u, 0xac45a401u}; return v[i]; }
BP_HD uint32_t glv_bls_mu(int i) { const uint32_t v[5] = {0xf6cfee30u, 0x63f6e522u, 0xe01faaddu, 0x7c6becf1u, 0x00000001u}; return v[i]; }   // floor(2^256 / lambda)
// beta (canonical limbs); Montgomery form by to_mont()
BP_HD Fp<BlsFq> glv_bls_beta() {
  const uint32_t v[12] = {0x0000aaacu, 0x8bfd0000u, 0x4f49fffdu, 0x409427ebu, 0x0fb85f9bu, 0x897d2965u,
                          0x89759ad4u, 0xaa0d857du, 0x63d4de85u, 0xec024086u, 0x397fe699u, 0x1a0111eau};
  Fp<BlsFq> b;
  for (int i = 0; i < 12; i++) b.v[i] = v[i];
  return b.to_mont();
}
// k (canonical, < r, 8 limbs) -> out[0..4) = k1, out[4..8) = k2
BP_HD void glv_bls_split(const uint32_t* k, uint32_t* out) {
  // q = floor(k * mu / 2^256): 8 x 5 limb product, limbs 8..12
  uint32_t prod[13];
  for (int i = 0; i < 13; i++) prod[i] = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < 5; j++) {
      const uint64_t t = (uint64_t)k[i] * glv_bls_mu(j) + prod[i + j] + carry;
      prod[i + j] = (uint32_t)t;
      carry = t >> 32;
    }
    prod[i + 5] = (uint32_t)carry;
  }
  uint32_t q[4] = {prod[8], prod[9], prod[10], prod[11]};          // < 2^128 (prod[12] = 0 for k < r)
  // rem = k - q * lambda  (< 3 * lambda: 5 limbs are enough, computed modulo 2^160)
  uint32_t ql[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < 4 && i + j < 5; j++) {
      const uint64_t t = (uint64_t)q[i] * glv_bls_lambda(j) + ql[i + j] + carry;
      ql[i + j] = (uint32_t)t;
      carry = t >> 32;
    }
    if (i + 4 < 5) ql[i + 4] = (uint32_t)carry;
  }
  uint32_t rem[5];
  uint64_t borrow = 0;
  for (int i = 0; i < 5; i++) {
    const uint64_t t = (uint64_t)k[i] - ql[i] - borrow;
    rem[i] = (uint32_t)t;
    borrow = (t >> 32) & 1;
  }
  for (int it = 0; it < 3; it++) {                                  // at most 2 corrections (measured: 1)
    bool ge = rem[4] != 0;
    if (!ge) {
      ge = true;
      for (int i = 3; i >= 0; i--) {
        if (rem[i] > glv_bls_lambda(i)) break;
        if (rem[i] < glv_bls_lambda(i)) { ge = false; break; }
      }
    }
    if (!ge) break;
    borrow = 0;
    for (int i = 0; i < 5; i++) {
      const uint64_t t = (uint64_t)rem[i] - (i < 4 ? glv_bls_lambda(i) : 0u) - borrow;
      rem[i] = (uint32_t)t;
      borrow = (t >> 32) & 1;
    }
    for (int i = 0; i < 4; i++) { if (++q[i] != 0) break; }
  }
  for (int i = 0; i < 4; i++) { out[i] = rem[i]; out[4 + i] = q[i]; }
}

// beta of the curve's GLV endomorphism in Montgomery form (only BLS12-381 uses it; BN254's balanced split needs a lattice)
template <class Curve> BP_HD typename Curve::Fq glv_beta();
template <> BP_HD Bls::Fq glv_beta<Bls>() { return glv_bls_beta(); }
template <> BP_HD Bn::Fq glv_beta<Bn>() { return Bn::Fq::zero(); }

}  // namespace bp
