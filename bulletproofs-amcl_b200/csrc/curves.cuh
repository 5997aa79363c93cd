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

}  // namespace bp
