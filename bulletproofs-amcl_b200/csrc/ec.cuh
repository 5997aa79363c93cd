// G1 group law for short-Weierstrass curves with a = 0 (BLS12-381: b = 4, AMCL BN254: b = 2).
//
// Device replacement for the AMCL ECP operations the reference reaches through amcl_wrapper's
// G1 (`+`, `binary_scalar_mul`, `*`; /root/reference/src/ipp.rs:119-129,185-187,
// src/r1cs/prover.rs:123,358,550).  All coordinates are Fp<...> in Montgomery form.
//
//  * Affine  : (x, y); the identity is stored as (0, 0), which is on neither curve.
//  * XYZZ    : (X, Y, ZZ, ZZZ) with x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity: ZZ = 0.
//    Mixed add 8M+2S, full add 12M+2S, double 6M+3S (EFD "xyzz" formulas for a = 0); the two products of every Y3
//    share one Montgomery reduction (Fp::mul2), which takes half a multiplication off each.
// The add routines are COMPLETE: equal inputs fall through to doubling and opposite inputs
// to the identity, because bucket sums do meet P+P and P+(-P) (repeated generators, 0/1 witnesses).
#pragma once
#include "fp.cuh"

namespace bp {

template <class F>
struct Affine {
  F x, y;
  BP_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  BP_HD static Affine inf() { Affine a; a.x = F::zero(); a.y = F::zero(); return a; }
  BP_HD Affine neg() const { Affine a; a.x = x; a.y = is_inf() ? y : y.neg(); return a; }
};

template <class F>
struct XYZZ {
  F x, y, zz, zzz;
  BP_HD bool is_inf() const { return zz.is_zero(); }
  BP_HD static XYZZ inf() { XYZZ p; p.x = F::zero(); p.y = F::zero(); p.zz = F::zero(); p.zzz = F::zero(); return p; }
  BP_HD static XYZZ from_affine(const Affine<F>& a) {
    XYZZ p;
    if (a.is_inf()) return inf();
    p.x = a.x; p.y = a.y; p.zz = F::one(); p.zzz = F::one();
    return p;
  }
  BP_HD XYZZ neg() const { XYZZ p = *this; p.y = y.neg(); return p; }

  // dbl-2008-s-1 with a = 0
  BP_HD_COLD void dbl() {
    if (is_inf()) return;
    if (y.is_zero()) { *this = inf(); return; }
    F U = y.dbl();
    F V = F::mulc(U, U);
    F W = F::mulc(U, V);
    F S = F::mulc(x, V);
    F X2 = F::mulc(x, x);
    F M = X2.dbl() + X2;
    F X3 = F::mulc(M, M) - S.dbl();
    F Y3 = F::mul2c(M, S - X3, W, y.neg());             // M*(S - X3) - W*y, one reduction
    zz = F::mulc(V, zz);
    zzz = F::mulc(W, zzz);
    x = X3; y = Y3;
  }

  // mdbl-2008-s-1: 2 * affine
  BP_HD_COLD static XYZZ dbl_affine(const Affine<F>& a) {
    if (a.is_inf() || a.y.is_zero()) return inf();
    XYZZ p;
    F U = a.y.dbl();
    F V = F::mulc(U, U);
    F W = F::mulc(U, V);
    F S = F::mulc(a.x, V);
    F X2 = F::mulc(a.x, a.x);
    F M = X2.dbl() + X2;
    p.x = F::mulc(M, M) - S.dbl();
    p.y = F::mul2c(M, S - p.x, W, a.y.neg());
    p.zz = V; p.zzz = W;
    return p;
  }

  // madd-2008-s: this += affine
  BP_HD void madd(const Affine<F>& a) {
    if (a.is_inf()) return;
    if (is_inf()) { *this = from_affine(a); return; }
    F U2 = a.x * zz;
    F S2 = a.y * zzz;
    F Pd = U2 - x;
    F R = S2 - y;
    if (Pd.is_zero()) {
      if (R.is_zero()) *this = dbl_affine(a); else *this = inf();
      return;
    }
    F PP = Pd.sqr();
    F PPP = Pd * PP;
    F Q = x * PP;
    F X3 = R.sqr() - PPP - Q.dbl();
    F Y3 = F::mul2(R, Q - X3, y.neg(), PPP);            // R*(Q - X3) - Y1*PPP, one reduction
    zz = zz * PP;
    zzz = zzz * PPP;
    x = X3; y = Y3;
  }

  // the same mixed addition with its products as real calls (Fp::mulc / sqrc / mul2c): ~1 KB of code at the call site
  // instead of ~60 KB, for kernels whose loop would otherwise not stay in the instruction cache
  BP_HD_COLD void madd_c(const Affine<F>& a) {
    if (a.is_inf()) return;
    if (is_inf()) { *this = from_affine(a); return; }
    F U2 = F::mulc(a.x, zz);
    F S2 = F::mulc(a.y, zzz);
    F Pd = U2 - x;
    F R = S2 - y;
    if (Pd.is_zero()) {
      if (R.is_zero()) *this = dbl_affine(a); else *this = inf();
      return;
    }
    F PP = F::sqrc(Pd);
    F PPP = F::mulc(Pd, PP);
    F Q = F::mulc(x, PP);
    F X3 = F::sqrc(R) - PPP - Q.dbl();
    F Y3 = F::mul2c(R, Q - X3, y.neg(), PPP);
    zz = F::mulc(zz, PP);
    zzz = F::mulc(zzz, PPP);
    x = X3; y = Y3;
  }

  // add-2008-s: this += q
  // compact form: a real call whose products are real calls too (see Fp::mulc) -- for latency-bound code
  BP_HD_COLD void add(const XYZZ& q) { add_t<true>(q); }
  // fully inlined into the caller (throughput loops whose operands live in registers)
  BP_HD void add_inl(const XYZZ& q) { add_t<false>(q); }
  template <bool COMPACT>
  BP_HD void add_t(const XYZZ& q) {
    if (q.is_inf()) return;
    if (is_inf()) { *this = q; return; }
    F U1 = mul_sel<COMPACT>(x, q.zz);
    F U2 = mul_sel<COMPACT>(q.x, zz);
    F S1 = mul_sel<COMPACT>(y, q.zzz);
    F S2 = mul_sel<COMPACT>(q.y, zzz);
    F Pd = U2 - U1;
    F R = S2 - S1;
    if (Pd.is_zero()) {
      if (R.is_zero()) dbl(); else *this = inf();
      return;
    }
    F PP = mul_sel<COMPACT>(Pd, Pd);
    F PPP = mul_sel<COMPACT>(Pd, PP);
    F Q = mul_sel<COMPACT>(U1, PP);
    F X3 = mul_sel<COMPACT>(R, R) - PPP - Q.dbl();
    F Y3 = COMPACT ? F::mul2c(R, Q - X3, S1.neg(), PPP) : F::mul2(R, Q - X3, S1.neg(), PPP);
    zz = mul_sel<COMPACT>(mul_sel<COMPACT>(zz, q.zz), PP);
    zzz = mul_sel<COMPACT>(mul_sel<COMPACT>(zzz, q.zzz), PPP);
    x = X3; y = Y3;
  }
  template <bool COMPACT>
  BP_HD static F mul_sel(const F& a, const F& b) { return COMPACT ? F::mulc(a, b) : a * b; }

  // normalise; one field inversion
  BP_HD_COLD Affine<F> to_affine() const {
    if (is_inf()) return Affine<F>::inf();
    F i3 = zzz.inv();          // 1/Z^3
    F i1 = i3 * zz;            // 1/Z
    Affine<F> a;
    a.x = x * i1.sqr();
    a.y = y * i3;
    return a;
  }
};

// lane-wise select: c ? a : b  (keeps both operands in registers, no dynamic indexing)
template <class F>
BP_HD XYZZ<F> select(bool c, const XYZZ<F>& a, const XYZZ<F>& b) {
  XYZZ<F> r;
  for (int i = 0; i < F::N; i++) {
    r.x.v[i] = c ? a.x.v[i] : b.x.v[i];
    r.y.v[i] = c ? a.y.v[i] : b.y.v[i];
    r.zz.v[i] = c ? a.zz.v[i] : b.zz.v[i];
    r.zzz.v[i] = c ? a.zzz.v[i] : b.zzz.v[i];
  }
  return r;
}

// k * P for a small unsigned integer k (bucket-chunk offsets), double-and-add MSB first
template <class F>
BP_HD XYZZ<F> mul_small(const XYZZ<F>& p, uint32_t k) {
  XYZZ<F> r = XYZZ<F>::inf();
  if (k == 0 || p.is_inf()) return r;
  int top = 31;
  while (!((k >> top) & 1)) top--;
  r = p;
  for (int b = top - 1; b >= 0; b--) {
    r.dbl();
    if ((k >> b) & 1) r.add(p);
  }
  return r;
}

// k * P for a multi-limb little-endian integer scalar (NOT Montgomery form)
template <class F>
BP_HD XYZZ<F> mul_limbs(const XYZZ<F>& p, const uint32_t* k, int nlimbs) {
  XYZZ<F> r = XYZZ<F>::inf();
  for (int i = nlimbs - 1; i >= 0; i--) {
    for (int b = 31; b >= 0; b--) {
      r.dbl();
      if ((k[i] >> b) & 1) r.add(p);
    }
  }
  return r;
}

}  // namespace bp
