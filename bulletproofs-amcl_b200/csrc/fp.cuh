// Montgomery prime-field arithmetic on 32-bit limbs for sm_100a (and a bit-identical host
// emulation used by the CPU unit tests of the limb schedule).
//
// Replaces, on the device, the field arithmetic the reference gets from amcl_wrapper's
// FieldElement / AMCL FP (call sites /root/reference/src/ipp.rs:115-130,181-188;
// src/r1cs/prover.rs:469-486).  Values are kept in Montgomery form (x*2^(32N) mod p),
// fully reduced to [0, p) after every operation, so equality is limb equality.
//
// Multiplier schedule: two accumulators, one holding the products that start on even limb
// positions and one those on odd positions.  Every 32x32->64 product is issued as a
// mad.lo.cc / madc.hi.cc pair on adjacent limbs of ONE accumulator, so ptxas fuses each pair
// into a single IMAD.WIDE.U32(.X) with the carry riding the chain: 2*N*N wide multiply-adds
// per Montgomery product (288 for the 12-limb BLS12-381 Fq, 128 for 8 limbs) plus N IMADs for
// the quotient digits.  After each reduction row the accumulators swap roles, which performs
// the one-limb right shift by register renaming.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BP_HD __host__ __device__ __forceinline__
#define BP_D __device__ __forceinline__
// cold-path group operations are real calls on the device: keeps code size and ptxas time bounded
#define BP_HD_COLD __host__ __device__ __noinline__
#else
#define BP_HD inline
#define BP_D inline
#define BP_HD_COLD inline
#endif

#include "field_params.h"

#ifndef BP_FQ_SPLIT
#define BP_FQ_SPLIT 0
#endif
#ifndef BP_SQR_DEDICATED
#define BP_SQR_DEDICATED 1
#endif

namespace bp {

// ---------------------------------------------------------------------------------------------
// Carry-flag primitives.  Device: PTX with the .cc flag.  Host: emulation with an explicit flag
// (lets tests/host_fp_check.cpp run the exact same limb schedule on the CPU).
// ---------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
struct CarryChain {
  BP_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  BP_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  BP_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  BP_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  BP_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  BP_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  BP_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  BP_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  BP_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  BP_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  BP_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  // un-fused halves of a product: plain IMAD / IMAD.HI that may issue on either FMA sub-pipe
  BP_D static uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  BP_D static uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
};
#else
struct CarryChain {
  uint32_t cf = 0;
  uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; cf = (uint32_t)(t >> 32); return (uint32_t)t; }
  uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + cf; cf = (uint32_t)(t >> 32); return (uint32_t)t; }
  uint32_t addc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + cf; return (uint32_t)t; }
  uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; cf = (uint32_t)(t >> 63); return (uint32_t)t; }
  uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - cf; cf = (uint32_t)(t >> 63); return (uint32_t)t; }
  uint32_t subc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - cf; return (uint32_t)t; }
  uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc((uint32_t)((uint64_t)a * b), c); }
  uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc((uint32_t)((uint64_t)a * b), c); }
  uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc((uint32_t)(((uint64_t)a * b) >> 32), c); }
  uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc((uint32_t)(((uint64_t)a * b) >> 32), c); }
  uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc((uint32_t)(((uint64_t)a * b) >> 32), c); }
  static uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
  static uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
};
#endif

template <class P>
struct Fp {
  using Params = P;
  static constexpr int N = P::N;
  uint32_t v[N];

  BP_HD static Fp zero() { Fp r; for (int i = 0; i < N; i++) r.v[i] = 0; return r; }
  BP_HD static Fp one() { Fp r; for (int i = 0; i < N; i++) r.v[i] = P::R1(i); return r; }   // Montgomery 1
  BP_HD static Fp r2() { Fp r; for (int i = 0; i < N; i++) r.v[i] = P::R2(i); return r; }
  BP_HD static Fp modulus() { Fp r; for (int i = 0; i < N; i++) r.v[i] = P::P(i); return r; }

  BP_HD bool is_zero() const { uint32_t o = 0; for (int i = 0; i < N; i++) o |= v[i]; return o == 0; }
  BP_HD bool operator==(const Fp& b) const { uint32_t o = 0; for (int i = 0; i < N; i++) o |= v[i] ^ b.v[i]; return o == 0; }
  BP_HD bool operator!=(const Fp& b) const { return !(*this == b); }

  // r = a - p if a >= p (a < 2p, possibly with an extra top carry bit `hi`)
  BP_HD static void reduce_once(uint32_t* a, uint32_t hi = 0) {
    CarryChain c;
    uint32_t t[N];
    t[0] = c.sub_cc(a[0], P::P(0));
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = c.subc_cc(a[i], P::P(i));
    uint32_t borrow = c.subc(hi, 0);          // 0 if a >= p, 0xffffffff otherwise
#pragma unroll
    for (int i = 0; i < N; i++) a[i] = borrow ? a[i] : t[i];
  }

  BP_HD friend Fp operator+(const Fp& a, const Fp& b) {
    Fp r;
    CarryChain c;
    r.v[0] = c.add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.v[i] = c.addc_cc(a.v[i], b.v[i]);
    uint32_t hi = c.addc(0, 0);
    reduce_once(r.v, hi);
    return r;
  }

  BP_HD friend Fp operator-(const Fp& a, const Fp& b) {
    Fp r;
    CarryChain c;
    r.v[0] = c.sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.v[i] = c.subc_cc(a.v[i], b.v[i]);
    uint32_t borrow = c.subc(0, 0);           // 0xffffffff when a < b
    CarryChain d;
    r.v[0] = d.add_cc(r.v[0], P::P(0) & borrow);
#pragma unroll
    for (int i = 1; i < N; i++) r.v[i] = d.addc_cc(r.v[i], P::P(i) & borrow);
    return r;
  }

  BP_HD Fp neg() const { return zero() - *this; }
  BP_HD Fp dbl() const { return *this + *this; }

  // ---- Montgomery product, even/odd accumulator schedule (see file header)
  BP_HD friend Fp operator*(const Fp& a, const Fp& b) { return mul_t<BP_FQ_SPLIT>(a, b); }

  // SP = number of products per accumulate chain issued as separate IMAD + IMAD.HI (+2 IADD3.X)
  // instead of one fused IMAD.WIDE.U32.X.  IMAD.WIDE only issues on the "fmaheavy" half of the FMA
  // pipe (measured: half the IMAD rate); plain IMADs can use the other half, so a mix balances them.
  template <int SP>
  BP_HD static Fp mul_t(const Fp& a, const Fp& b) {
    static_assert(N % 2 == 0, "even limb count");
    uint32_t ev[N], od[N];
    // row 0: plain products, no accumulation
    {
      const uint32_t bi = b.v[0];
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        uint64_t pe = (uint64_t)a.v[j] * bi;       // IMAD.WIDE.U32
        ev[j] = (uint32_t)pe; ev[j + 1] = (uint32_t)(pe >> 32);
        uint64_t po = (uint64_t)a.v[j + 1] * bi;
        od[j] = (uint32_t)po; od[j + 1] = (uint32_t)(po >> 32);
      }
      reduce_row<SP>(ev, od);
    }
#pragma unroll
    for (int i = 1; i < N; i++) {
      // roles alternate: after a reduction row the accumulator whose limb 0 was cleared is,
      // shifted right by one limb, the odd accumulator of the next row.
      if (i & 1) { mul_row<SP>(od, ev, a.v, b.v[i]); reduce_row<SP>(od, ev); }
      else       { mul_row<SP>(ev, od, a.v, b.v[i]); reduce_row<SP>(ev, od); }
    }
    // after N rows (N even) the cleared accumulator is `od`... the last reduce_row call had
    // (even=od, odd=ev) when N-1 is odd.
    Fp r;
    {
      uint32_t* e = ((N - 1) & 1) ? od : ev;   // accumulator with limb 0 == 0
      uint32_t* o = ((N - 1) & 1) ? ev : od;
      CarryChain c;
      r.v[0] = c.add_cc(o[0], e[1]);
#pragma unroll
      for (int k = 1; k < N - 1; k++) r.v[k] = c.addc_cc(o[k], e[k + 1]);
      r.v[N - 1] = c.addc(o[N - 1], 0);
    }
    reduce_once(r.v);
    return r;
  }

#if BP_SQR_DEDICATED
  BP_HD Fp sqr() const { return sqr_t(*this); }
#else
  BP_HD Fp sqr() const { return (*this) * (*this); }
#endif

  // Dedicated squaring: the N(N-1)/2 products above the diagonal once, doubled, plus the N squares (N(N+1)/2 wide
  // multiply-adds instead of N^2), as a 2N-limb integer T; then N reduction rows on T's low half (the row of mul_t without
  // the a*b_i part: the one-limb shift rides the accumulate of m*p) and ONE addition of T's high half:
  // REDC(T_lo + 2^(32N) T_hi) = REDC(T_lo) + T_hi < 2p.  222 + 12 multiply-adds instead of 288 + 12 for the 12-limb field.
  BP_HD static Fp sqr_t(const Fp& a) {
    static_assert(N % 2 == 0, "even limb count");
    uint32_t EV[2 * N], OD[2 * N];          // EV[k]: limb position k ; OD[k]: position k + 1
#pragma unroll
    for (int k = 0; k < 2 * N; k++) { EV[k] = 0; OD[k] = 0; }
#pragma unroll
    for (int i = 0; i < N - 1; i++) {
      {                                      // j = i+1, i+3, ...: i + j odd -> OD[i+j-1], OD[i+j]
        CarryChain c;
        int p = i + (i + 1) - 1;
        OD[p] = c.mad_lo_cc(a.v[i], a.v[i + 1], OD[p]);
        OD[p + 1] = c.madc_hi_cc(a.v[i], a.v[i + 1], OD[p + 1]);
#pragma unroll
        for (int j = i + 3; j < N; j += 2) {
          p = i + j - 1;
          OD[p] = c.madc_lo_cc(a.v[i], a.v[j], OD[p]);
          OD[p + 1] = c.madc_hi_cc(a.v[i], a.v[j], OD[p + 1]);
        }
        OD[p + 2] = c.addc(OD[p + 2], 0);    // lands on a limb that holds only earlier rows' carries so far
      }
      if (i + 2 < N) {                       // j = i+2, i+4, ...: i + j even -> EV[i+j], EV[i+j+1]
        CarryChain c;
        int p = i + (i + 2);
        EV[p] = c.mad_lo_cc(a.v[i], a.v[i + 2], EV[p]);
        EV[p + 1] = c.madc_hi_cc(a.v[i], a.v[i + 2], EV[p + 1]);
#pragma unroll
        for (int j = i + 4; j < N; j += 2) {
          p = i + j;
          EV[p] = c.madc_lo_cc(a.v[i], a.v[j], EV[p]);
          EV[p + 1] = c.madc_hi_cc(a.v[i], a.v[j], EV[p + 1]);
        }
        EV[p + 2] = c.addc(EV[p + 2], 0);
      }
    }
    {                                        // double both halves of the triangle (the total is < 2^(64N-1))
      CarryChain c, d;
      EV[0] = c.add_cc(EV[0], EV[0]);
#pragma unroll
      for (int k = 1; k < 2 * N - 1; k++) EV[k] = c.addc_cc(EV[k], EV[k]);
      EV[2 * N - 1] = c.addc(EV[2 * N - 1], EV[2 * N - 1]);
      OD[0] = d.add_cc(OD[0], OD[0]);
#pragma unroll
      for (int k = 1; k < 2 * N - 1; k++) OD[k] = d.addc_cc(OD[k], OD[k]);
      OD[2 * N - 1] = d.addc(OD[2 * N - 1], OD[2 * N - 1]);
    }
    {                                        // the diagonal: a_i^2 at position 2i
      CarryChain c;
      EV[0] = c.mad_lo_cc(a.v[0], a.v[0], EV[0]);
      EV[1] = c.madc_hi_cc(a.v[0], a.v[0], EV[1]);
#pragma unroll
      for (int i = 1; i < N - 1; i++) {
        EV[2 * i] = c.madc_lo_cc(a.v[i], a.v[i], EV[2 * i]);
        EV[2 * i + 1] = c.madc_hi_cc(a.v[i], a.v[i], EV[2 * i + 1]);
      }
      EV[2 * N - 2] = c.madc_lo_cc(a.v[N - 1], a.v[N - 1], EV[2 * N - 2]);
      EV[2 * N - 1] = c.madc_hi(a.v[N - 1], a.v[N - 1], EV[2 * N - 1]);      // a^2 < 2^(64N): no carry out
    }
    uint32_t hi[N], ev[N], od[N];
    {                                        // T = EV + 2^32 OD as ONE integer: low half -> the window, high half kept
      CarryChain c;
      ev[0] = EV[0];
      ev[1] = c.add_cc(EV[1], OD[0]);
#pragma unroll
      for (int k = 2; k < N; k++) ev[k] = c.addc_cc(EV[k], OD[k - 1]);
#pragma unroll
      for (int k = 0; k < N - 1; k++) hi[k] = c.addc_cc(EV[N + k], OD[N + k - 1]);
      hi[N - 1] = c.addc(EV[2 * N - 1], OD[2 * N - 2]);
#pragma unroll
      for (int k = 0; k < N; k++) od[k] = 0;
    }
    reduce_row<0>(ev, od);
#pragma unroll
    for (int i = 1; i < N; i++) {
      if (i & 1) redc_row(od, ev); else redc_row(ev, od);
    }
    Fp r;
    {
      uint32_t* e = ((N - 1) & 1) ? od : ev;   // accumulator with limb 0 == 0
      uint32_t* o = ((N - 1) & 1) ? ev : od;
      CarryChain c;
      r.v[0] = c.add_cc(o[0], e[1]);
#pragma unroll
      for (int k = 1; k < N - 1; k++) r.v[k] = c.addc_cc(o[k], e[k + 1]);
      r.v[N - 1] = c.addc(o[N - 1], 0);
      CarryChain d;
      r.v[0] = d.add_cc(r.v[0], hi[0]);
#pragma unroll
      for (int k = 1; k < N - 1; k++) r.v[k] = d.addc_cc(r.v[k], hi[k]);
      r.v[N - 1] = d.addc(r.v[N - 1], hi[N - 1]);
    }
    reduce_once(r.v);
    return r;
  }

  // The same products as REAL CALLS on the device.  An inlined product is ~400 straight-line instructions (6 KB); a
  // group addition built from inlined products is ~90 KB of code that a latency-bound kernel (tree levels, the bucket
  // reduction's tail, single additions) runs through once per warp, so it waits on instruction fetch rather than on
  // the multiplier.  The cold-path group operations (XYZZ::add / dbl) are built from these instead: one 6 KB body
  // that stays in the instruction cache.  Throughput kernels keep the inlined forms.
  BP_HD_COLD static Fp mulc(Fp a, Fp b) { return mul_t<0>(a, b); }
  BP_HD_COLD static Fp sqrc(Fp a) { return a.sqr(); }
  BP_HD_COLD static Fp mul2c(Fp a, Fp b, Fp c, Fp d) { return mul2(a, b, c, d); }

  // (a*b + c*d) / 2^(32N) mod p with ONE Montgomery reduction: every row adds a*b_i AND c*d_i to the accumulators
  // before its reduction row, so the sum of two products costs 2*N*N + N*N + N wide multiply-adds instead of
  // 2 * (2*N*N + N) -- 444 instead of 600 for the 12-limb field.  The running value stays below 3p
  // ((3p + (2^32-1)*3p) / 2^32 = 3p), which must fit N limbs: true for both curve fields Fq (381 / 254 bits),
  // not for the 255-bit BLS12-381 Fr.  Used for Y3 = R*(Q - X3) + (-Y1)*PPP of the group law (ec.cuh).
  BP_HD static Fp mul2(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
    static_assert(N % 2 == 0, "even limb count");
    static_assert(P::BITS <= 32 * N - 2, "3p must fit N limbs");
    uint32_t ev[N], od[N];
    {
      const uint32_t bi = b.v[0];
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        uint64_t pe = (uint64_t)a.v[j] * bi;
        ev[j] = (uint32_t)pe; ev[j + 1] = (uint32_t)(pe >> 32);
        uint64_t po = (uint64_t)a.v[j + 1] * bi;
        od[j] = (uint32_t)po; od[j + 1] = (uint32_t)(po >> 32);
      }
      mac_row(ev, od, c.v, d.v[0]);
      reduce_row<0>(ev, od);
    }
#pragma unroll
    for (int i = 1; i < N; i++) {
      if (i & 1) { mul_row<0>(od, ev, a.v, b.v[i]); mac_row(od, ev, c.v, d.v[i]); reduce_row<0>(od, ev); }
      else       { mul_row<0>(ev, od, a.v, b.v[i]); mac_row(ev, od, c.v, d.v[i]); reduce_row<0>(ev, od); }
    }
    Fp r;
    {
      uint32_t* e = ((N - 1) & 1) ? od : ev;
      uint32_t* o = ((N - 1) & 1) ? ev : od;
      CarryChain k;
      r.v[0] = k.add_cc(o[0], e[1]);
#pragma unroll
      for (int t = 1; t < N - 1; t++) r.v[t] = k.addc_cc(o[t], e[t + 1]);
      r.v[N - 1] = k.addc(o[N - 1], 0);
    }
    reduce_once(r.v);                          // < 3p -> < 2p -> < p
    reduce_once(r.v);
    return r;
  }

  // out-of-Montgomery: multiply by 1
  BP_HD Fp from_mont() const { Fp o; for (int i = 0; i < N; i++) o.v[i] = (i == 0); return (*this) * o; }
  BP_HD Fp to_mont() const { return (*this) * r2(); }

  // x^e for a public N-limb exponent given by a getter (MSB first square-and-multiply; variable time)
  template <class E>
  BP_HD Fp pow_limbs(E exp_limb, int nlimbs) const {
    Fp acc = one();
    bool started = false;
    for (int i = nlimbs - 1; i >= 0; i--) {
      uint32_t w = exp_limb(i);
      for (int b = 31; b >= 0; b--) {
        if (started) acc = acc.sqr();
        if ((w >> b) & 1) { acc = started ? acc * (*this) : *this; started = true; }
      }
    }
    return acc;
  }
  // Fermat inverse; 0 -> 0 (FieldElement::inverse of 0 is 0 in AMCL)
  BP_HD_COLD Fp inv() const { return pow_limbs([](int i) { return P::PM2(i); }, N); }

 private:
  // even := od_prev (already at even positions) ; odd := ev_prev >> 64 (register renaming);
  // then even += a_even*bi, odd += a_odd*bi.
  //   `even` enters holding the previous odd accumulator, `odd` enters holding the previous even
  //   accumulator whose limb 0 is zero and whose limb 1 still has to be added at position 0.
  template <int SP>
  BP_HD static void mul_row(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi) {
    CarryChain c;
    even[0] = c.add_cc(even[0], odd[1]);                 // carry -> position 1 = first limb of the odd chain
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {                 // odd[k] <- odd[k+2] + a[j+1]*bi  (shifted accumulate)
      if (j >= N - 2 * SP) {
        odd[j] = c.addc_cc(CarryChain::mul_lo(a[j + 1], bi), odd[j + 2]);
        odd[j + 1] = c.addc_cc(CarryChain::mul_hi(a[j + 1], bi), odd[j + 3]);
      } else {
        odd[j] = c.madc_lo_cc(a[j + 1], bi, odd[j + 2]);
        odd[j + 1] = c.madc_hi_cc(a[j + 1], bi, odd[j + 3]);
      }
    }
    odd[N - 2] = c.madc_lo_cc(a[N - 1], bi, 0);
    odd[N - 1] = c.madc_hi(a[N - 1], bi, 0);            // no carry out: value bound < 2^(32(N+1))
    CarryChain d;
    even[0] = d.mad_lo_cc(a[0], bi, even[0]);
    even[1] = d.madc_hi_cc(a[0], bi, even[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      if (j >= N - 2 * SP) {
        even[j] = d.addc_cc(CarryChain::mul_lo(a[j], bi), even[j]);
        even[j + 1] = d.addc_cc(CarryChain::mul_hi(a[j], bi), even[j + 1]);
      } else {
        even[j] = d.madc_lo_cc(a[j], bi, even[j]);
        even[j + 1] = d.madc_hi_cc(a[j], bi, even[j + 1]);
      }
    }
    odd[N - 1] = d.addc(odd[N - 1], 0);                  // position N lives in odd[N-1]
  }

  // pure accumulate on the current window (no shift): even += c_even*di, odd += c_odd*di
  //   even[k] is limb position k, odd[k] is position k+1; the top limb (position N) lives in odd[N-1]
  BP_HD static void mac_row(uint32_t* even, uint32_t* odd, const uint32_t* c, uint32_t di) {
    CarryChain o;
    odd[0] = o.mad_lo_cc(c[1], di, odd[0]);
    odd[1] = o.madc_hi_cc(c[1], di, odd[1]);
#pragma unroll
    for (int j = 2; j < N - 2; j += 2) {
      odd[j] = o.madc_lo_cc(c[j + 1], di, odd[j]);
      odd[j + 1] = o.madc_hi_cc(c[j + 1], di, odd[j + 1]);
    }
    odd[N - 2] = o.madc_lo_cc(c[N - 1], di, odd[N - 2]);
    odd[N - 1] = o.madc_hi(c[N - 1], di, odd[N - 1]);   // no carry out: the window value stays below 2^(32(N+1))
    CarryChain e;
    even[0] = e.mad_lo_cc(c[0], di, even[0]);
    even[1] = e.madc_hi_cc(c[0], di, even[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      even[j] = e.madc_lo_cc(c[j], di, even[j]);
      even[j + 1] = e.madc_hi_cc(c[j], di, even[j + 1]);
    }
    odd[N - 1] = e.addc(odd[N - 1], 0);
  }

  // a reduction row WITH the one-limb shift (roles as in mul_row): even := previous odd accumulator, odd := previous even
  // accumulator (limb 0 cleared) two limbs down; m from the new limb 0; even += p_even*m, odd = shifted odd + p_odd*m
  BP_HD static void redc_row(uint32_t* even, uint32_t* odd) {
    CarryChain c;
    even[0] = c.add_cc(even[0], odd[1]);                 // carry -> position 1 = first limb of the odd chain
    const uint32_t m = even[0] * P::INV;
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
      odd[j] = c.madc_lo_cc(P::PC(j + 1), m, odd[j + 2]);
      odd[j + 1] = c.madc_hi_cc(P::PC(j + 1), m, odd[j + 3]);
    }
    odd[N - 2] = c.madc_lo_cc(P::PC(N - 1), m, 0);
    odd[N - 1] = c.madc_hi(P::PC(N - 1), m, 0);
    CarryChain d;
    even[0] = d.mad_lo_cc(P::PC(0), m, even[0]);
    even[1] = d.madc_hi_cc(P::PC(0), m, even[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      even[j] = d.madc_lo_cc(P::PC(j), m, even[j]);
      even[j + 1] = d.madc_hi_cc(P::PC(j), m, even[j + 1]);
    }
    odd[N - 1] = d.addc(odd[N - 1], 0);                  // position N lives in odd[N-1]
  }

  template <int SP>
  BP_HD static void reduce_row(uint32_t* even, uint32_t* odd) {
    const uint32_t m = even[0] * P::INV;
    CarryChain c;
    odd[0] = c.mad_lo_cc(P::PC(1), m, odd[0]);
    odd[1] = c.madc_hi_cc(P::PC(1), m, odd[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      if (j >= N - 2 * SP) {
        odd[j] = c.addc_cc(CarryChain::mul_lo(P::PC(j + 1), m), odd[j]);
        odd[j + 1] = c.addc_cc(CarryChain::mul_hi(P::PC(j + 1), m), odd[j + 1]);
      } else {
        odd[j] = c.madc_lo_cc(P::PC(j + 1), m, odd[j]);
        odd[j + 1] = c.madc_hi_cc(P::PC(j + 1), m, odd[j + 1]);
      }
    }
    CarryChain d;
    even[0] = d.mad_lo_cc(P::PC(0), m, even[0]);
    even[1] = d.madc_hi_cc(P::PC(0), m, even[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      if (j >= N - 2 * SP) {
        even[j] = d.addc_cc(CarryChain::mul_lo(P::PC(j), m), even[j]);
        even[j + 1] = d.addc_cc(CarryChain::mul_hi(P::PC(j), m), even[j + 1]);
      } else {
        even[j] = d.madc_lo_cc(P::PC(j), m, even[j]);
        even[j + 1] = d.madc_hi_cc(P::PC(j), m, even[j + 1]);
      }
    }
    odd[N - 1] = d.addc(odd[N - 1], 0);
  }
};

}  // namespace bp
