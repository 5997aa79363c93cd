// Host-side (CPU) Montgomery arithmetic on 64-bit limbs, same Montgomery radix as the device
// (2^(32*N32) = 2^(64*NL)), so device values can be consumed without conversion.
//
// Used by the host layer for the parts of the path that are serial by nature and tiny:
//   * the Horner combination of the <= 64 per-window sums an MSM leaves on the device and the
//     final affine normalisation (one field inversion) -- a dependent chain of ~256 point
//     doublings that one GPU thread needs milliseconds for and one CPU core ~0.1 ms;
//   * Fr arithmetic on challenges (u^-1, y^-1, x^k ...) in the transcript-driven orchestration
//     (/root/reference/src/ipp.rs:112-113, src/r1cs/prover.rs:438-463,508-546).
// This is product code, not the oracle: the oracle (oracle/c) has its own independent field code.
#pragma once
#include <stdint.h>
#include <string.h>

#include "field_params.h"

namespace bp {
namespace host {

template <class P>
struct HFp {
  static constexpr int NL = P::N / 2;
  uint64_t v[NL];

  static constexpr uint64_t pl(int i) { return (uint64_t)P::P(2 * i) | ((uint64_t)P::P(2 * i + 1) << 32); }
  static constexpr uint64_t inv64() {
    // -p^-1 mod 2^64 by Newton iteration from the 32-bit constant
    uint64_t p0 = pl(0), x = (uint64_t)(0u - P::INV);  // x = p^-1 mod 2^32
    x *= 2 - p0 * x;                                    // mod 2^64
    return 0 - x;
  }
  static HFp zero() { HFp r; memset(r.v, 0, sizeof r.v); return r; }
  static HFp one() { HFp r; for (int i = 0; i < NL; i++) r.v[i] = (uint64_t)P::R1(2 * i) | ((uint64_t)P::R1(2 * i + 1) << 32); return r; }
  static HFp r2() { HFp r; for (int i = 0; i < NL; i++) r.v[i] = (uint64_t)P::R2(2 * i) | ((uint64_t)P::R2(2 * i + 1) << 32); return r; }
  static HFp from_u64(uint64_t x) { HFp r = zero(); r.v[0] = x; return r.to_mont(); }

  bool is_zero() const { uint64_t o = 0; for (int i = 0; i < NL; i++) o |= v[i]; return o == 0; }
  bool operator==(const HFp& b) const { uint64_t o = 0; for (int i = 0; i < NL; i++) o |= v[i] ^ b.v[i]; return o == 0; }
  bool operator!=(const HFp& b) const { return !(*this == b); }

  static bool geq_p(const uint64_t* a) {
    for (int i = NL - 1; i >= 0; i--) { uint64_t p = pl(i); if (a[i] > p) return true; if (a[i] < p) return false; }
    return true;
  }
  static void sub_p(uint64_t* a) {
    unsigned __int128 br = 0;
    for (int i = 0; i < NL; i++) { unsigned __int128 t = (unsigned __int128)a[i] - pl(i) - (uint64_t)br; a[i] = (uint64_t)t; br = (t >> 64) & 1; }
  }
  friend HFp operator+(const HFp& a, const HFp& b) {
    HFp r; unsigned __int128 c = 0;
    for (int i = 0; i < NL; i++) { c += (unsigned __int128)a.v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
    if (c || geq_p(r.v)) sub_p(r.v);
    return r;
  }
  friend HFp operator-(const HFp& a, const HFp& b) {
    HFp r; unsigned __int128 br = 0;
    for (int i = 0; i < NL; i++) { unsigned __int128 t = (unsigned __int128)a.v[i] - b.v[i] - (uint64_t)br; r.v[i] = (uint64_t)t; br = (t >> 64) & 1; }
    if (br) { unsigned __int128 c = 0; for (int i = 0; i < NL; i++) { c += (unsigned __int128)r.v[i] + pl(i); r.v[i] = (uint64_t)c; c >>= 64; } }
    return r;
  }
  HFp neg() const { return zero() - *this; }
  HFp dbl() const { return *this + *this; }
  // Montgomery product, multiplication and reduction fused limb by limb (the "no-carry" form: the top bit of every modulus
  // here is clear, so t[NL-1] = c1 + c2 cannot overflow and no extra accumulator limb is needed).  Constants are
  // compile-time, loops have constant bounds and are unrolled: the NL accumulator limbs stay in registers.
  friend HFp operator*(const HFp& a, const HFp& b) {
    static_assert((P::P(P::N - 1) >> 31) == 0, "top bit of the modulus must be clear");
    constexpr uint64_t INV = inv64();
    uint64_t t[NL] = {0};
#pragma GCC unroll 8
    for (int i = 0; i < NL; i++) {
      const uint64_t bi = b.v[i];
      unsigned __int128 c = (unsigned __int128)a.v[0] * bi + t[0];
      const uint64_t t0 = (uint64_t)c;
      uint64_t c1 = (uint64_t)(c >> 64);
      const uint64_t m = t0 * INV;
      unsigned __int128 d = (unsigned __int128)m * pl(0) + t0;
      uint64_t c2 = (uint64_t)(d >> 64);
#pragma GCC unroll 8
      for (int j = 1; j < NL; j++) {
        c = (unsigned __int128)a.v[j] * bi + t[j] + c1;
        c1 = (uint64_t)(c >> 64);
        d = (unsigned __int128)m * pl(j) + (uint64_t)c + c2;
        c2 = (uint64_t)(d >> 64);
        t[j - 1] = (uint64_t)d;
      }
      t[NL - 1] = c1 + c2;
    }
    if (geq_p(t)) sub_p(t);
    HFp r;
#pragma GCC unroll 8
    for (int i = 0; i < NL; i++) r.v[i] = t[i];
    return r;
  }
  HFp sqr() const { return (*this) * (*this); }
  HFp to_mont() const { return (*this) * r2(); }
  HFp from_mont() const { HFp o = zero(); o.v[0] = 1; return (*this) * o; }
  // Inverse, 0 -> 0 (FieldElement::inverse of zero is zero in AMCL): binary extended Euclid on the Montgomery
  // representative m = aR (HAC 14.61: ~2 * bits shift / subtract steps on NL limbs -- a few microseconds where the
  // Fermat ladder below needs bits squarings + bits/2 products), then (aR)^-1 * R^3 / R = a^-1 R.  Variable time: the
  // host inverts challenges and the Z coordinates of results that are about to be published.
  HFp inv() const {
    if (is_zero()) return zero();
    uint64_t u[NL], v[NL], x1[NL], x2[NL];
    for (int i = 0; i < NL; i++) { u[i] = this->v[i]; v[i] = pl(i); x1[i] = 0; x2[i] = 0; }
    x1[0] = 1;
    auto is_one = [](const uint64_t* a) { uint64_t o = a[0] ^ 1; for (int i = 1; i < NL; i++) o |= a[i]; return o == 0; };
    auto shr1 = [](uint64_t* a, uint64_t top) { for (int i = 0; i < NL - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 63); a[NL - 1] = (a[NL - 1] >> 1) | (top << 63); };
    auto half_mod = [&](uint64_t* x) {            // x / 2 mod p for x < p
      uint64_t top = 0;
      if (x[0] & 1) { unsigned __int128 c = 0; for (int i = 0; i < NL; i++) { c += (unsigned __int128)x[i] + pl(i); x[i] = (uint64_t)c; c >>= 64; } top = (uint64_t)c; }
      shr1(x, top);
    };
    auto geq = [](const uint64_t* a, const uint64_t* b) { for (int i = NL - 1; i >= 0; i--) { if (a[i] > b[i]) return true; if (a[i] < b[i]) return false; } return true; };
    auto sub = [](uint64_t* a, const uint64_t* b) { unsigned __int128 br = 0; for (int i = 0; i < NL; i++) { unsigned __int128 t = (unsigned __int128)a[i] - b[i] - (uint64_t)br; a[i] = (uint64_t)t; br = (t >> 64) & 1; } return (uint64_t)br; };
    auto sub_mod = [&](uint64_t* a, const uint64_t* b) {   // a - b mod p for a, b < p
      if (sub(a, b)) { unsigned __int128 c = 0; for (int i = 0; i < NL; i++) { c += (unsigned __int128)a[i] + pl(i); a[i] = (uint64_t)c; c >>= 64; } }
    };
    while (!is_one(u) && !is_one(v)) {
      while (!(u[0] & 1)) { shr1(u, 0); half_mod(x1); }
      while (!(v[0] & 1)) { shr1(v, 0); half_mod(x2); }
      if (geq(u, v)) { sub(u, v); sub_mod(x1, x2); } else { sub(v, u); sub_mod(x2, x1); }
    }
    HFp w;
    for (int i = 0; i < NL; i++) w.v[i] = is_one(u) ? x1[i] : x2[i];
    static const HFp R3 = r2() * r2();             // R^2 * R^2 / R
    return w * R3;
  }
  // Fermat ladder (kept as the cross-check of inv() in the CPU tests)
  HFp inv_fermat() const {
    uint64_t e[NL];
    for (int i = 0; i < NL; i++) e[i] = pl(i);
    e[0] -= 2;
    HFp acc = one();
    for (int i = NL - 1; i >= 0; i--)
      for (int b = 63; b >= 0; b--) { acc = acc.sqr(); if ((e[i] >> b) & 1) acc = acc * (*this); }
    return acc;
  }
  // canonical big-endian bytes <-> Montgomery
  static HFp from_be(const uint8_t* be, int nbytes) {
    HFp t;
    for (int i = 0; i < NL; i++) {
      uint64_t w = 0; const uint8_t* p = be + nbytes - 8 * (i + 1);
      for (int k = 0; k < 8; k++) w = (w << 8) | p[k];
      t.v[i] = w;
    }
    while (geq_p(t.v)) sub_p(t.v);
    return t.to_mont();
  }
  void to_be(uint8_t* be, int nbytes) const {
    HFp t = from_mont();
    memset(be, 0, nbytes);
    for (int i = 0; i < NL; i++) { uint8_t* p = be + nbytes - 8 * (i + 1); for (int k = 0; k < 8; k++) p[k] = (uint8_t)(t.v[i] >> (56 - 8 * k)); }
  }
};

// Is X || Y (big endian, `modbytes` each) a point as AMCL's ECP::frombytes / ECP::new_bigs decide it: both coordinates
// < p and y^2 = x^3 + b, or the identity's encoding (0, 1)?  Host twin of csrc/curves.cuh g1_from_be_checked: the host
// layer validates every untrusted point (proof points, commitments, bases) before its bytes are hashed or shipped.
template <class P>
inline bool g1_xy_is_valid(const uint8_t* xy, int modbytes, unsigned curve_b) {
  using F = HFp<P>;
  uint64_t raw[2][F::NL];
  for (int c = 0; c < 2; c++) {
    const uint8_t* be = xy + c * modbytes;
    for (int i = 0; i < F::NL; i++) {
      uint64_t w = 0; const uint8_t* p = be + modbytes - 8 * (i + 1);
      for (int k = 0; k < 8; k++) w = (w << 8) | p[k];
      raw[c][i] = w;
    }
    if (F::geq_p(raw[c])) return false;
  }
  uint64_t xz = 0, y1 = raw[1][0] ^ 1;
  for (int i = 0; i < F::NL; i++) { xz |= raw[0][i]; if (i) y1 |= raw[1][i]; }
  if (xz == 0 && y1 == 0) return true;
  F x, y;
  for (int i = 0; i < F::NL; i++) { x.v[i] = raw[0][i]; y.v[i] = raw[1][i]; }
  x = x.to_mont(); y = y.to_mont();
  return y.sqr() == x.sqr() * x + F::from_u64(curve_b);
}

// XYZZ point on the host; layout-compatible with the device's XYZZ<Fp<P>> (little-endian limbs)
template <class P>
struct HXYZZ {
  using F = HFp<P>;
  F x, y, zz, zzz;
  bool is_inf() const { return zz.is_zero(); }
  static HXYZZ inf() { HXYZZ p; p.x = p.y = p.zz = p.zzz = F::zero(); return p; }
  void dbl() {
    if (is_inf()) return;
    if (y.is_zero()) { *this = inf(); return; }
    F U = y.dbl(), V = U.sqr(), W = U * V, S = x * V, X2 = x.sqr(), M = X2.dbl() + X2;
    F X3 = M.sqr() - S.dbl();
    F Y3 = M * (S - X3) - W * y;
    zz = V * zz; zzz = W * zzz; x = X3; y = Y3;
  }
  void add(const HXYZZ& q) {
    if (q.is_inf()) return;
    if (is_inf()) { *this = q; return; }
    F U1 = x * q.zz, U2 = q.x * zz, S1 = y * q.zzz, S2 = q.y * zzz;
    F Pd = U2 - U1, R = S2 - S1;
    if (Pd.is_zero()) { if (R.is_zero()) dbl(); else *this = inf(); return; }
    F PP = Pd.sqr(), PPP = Pd * PP, Q = U1 * PP;
    F X3 = R.sqr() - PPP - Q.dbl();
    F Y3 = R * (Q - X3) - S1 * PPP;
    zz = zz * q.zz * PP; zzz = zzz * q.zzz * PPP; x = X3; y = Y3;
  }
  // X||Y big endian, identity = AMCL's (0, 1)
  void to_xy_be(int modbytes, uint8_t* out) const {
    if (is_inf()) { memset(out, 0, 2 * modbytes); out[2 * modbytes - 1] = 1; return; }
    F i3 = zzz.inv(), i1 = i3 * zz;
    F ax = x * i1.sqr(), ay = y * i3;
    ax.to_be(out, modbytes); ay.to_be(out + modbytes, modbytes);
  }
};

}  // namespace host
}  // namespace bp
