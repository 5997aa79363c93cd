// Host build of the device headers (fp.cuh / ec.cuh) so the exact limb schedule and group
// formulas the kernels use can be unit-tested on a machine without a GPU.  Compiled by
// tests/test_host_arith.py with g++; NOT part of the product library.
#include <string.h>
#include "../bulletproofs-amcl_b200/csrc/ec.cuh"
#include "../bulletproofs-amcl_b200/csrc/host_fp.h"

using namespace bp;

template <class F>
static void fp_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  F x, y, r;
  memcpy(x.v, a, sizeof(x.v));
  memcpy(y.v, b, sizeof(y.v));
  switch (op) {
    case 0: r = x * y; break;
    case 1: r = x + y; break;
    case 2: r = x - y; break;
    case 3: r = x.to_mont(); break;
    case 4: r = x.from_mont(); break;
    case 5: r = x.inv(); break;
    case 6: r = x.sqr(); break;
    default: r = F::zero();
  }
  memcpy(out, r.v, sizeof(r.v));
}

// points are passed as Montgomery-form limbs: affine = x|y, xyzz = x|y|zz|zzz
template <class F>
static void ec_op(int op, const uint32_t* p, const uint32_t* q, uint32_t k, uint32_t* out) {
  XYZZ<F> P;
  memcpy(&P, p, sizeof(P));
  if (op == 0) { Affine<F> A; memcpy(&A, q, sizeof(A)); P.madd(A); }
  else if (op == 1) { XYZZ<F> Q; memcpy(&Q, q, sizeof(Q)); P.add(Q); }
  else if (op == 2) { P.dbl(); }
  else if (op == 3) { P = mul_small(P, k); }
  else if (op == 4) { Affine<F> A = P.to_affine(); P = XYZZ<F>::from_affine(A); }
  memcpy(out, &P, sizeof(P));
}

template <class F>
static void fp_mul2(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* out) {
  F x, y, z, w;
  memcpy(x.v, a, sizeof(x.v)); memcpy(y.v, b, sizeof(y.v)); memcpy(z.v, c, sizeof(z.v)); memcpy(w.v, d, sizeof(w.v));
  F r = F::mul2(x, y, z, w);
  memcpy(out, r.v, sizeof(r.v));
}

// host-side field code (host_fp.h, 64-bit limbs): inv() by binary extended Euclid against the Fermat ladder; in/out are
// the 32-bit little-endian limbs of the Montgomery form
template <class P>
static int hostfp_inv(const uint32_t* a, uint32_t* out) {
  using H = host::HFp<P>;
  H x;
  memcpy(x.v, a, sizeof x.v);
  H i1 = x.inv(), i2 = x.inv_fermat();
  memcpy(out, i1.v, sizeof i1.v);
  return i1 == i2 ? 1 : 0;
}

template <class P>
static void hostfp_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  using H = host::HFp<P>;
  H x, y, r;
  memcpy(x.v, a, sizeof x.v);
  memcpy(y.v, b, sizeof y.v);
  switch (op) {
    case 0: r = x * y; break;
    case 1: r = x + y; break;
    case 2: r = x - y; break;
    case 3: r = x.to_mont(); break;
    case 4: r = x.from_mont(); break;
    default: r = x.sqr();
  }
  memcpy(out, r.v, sizeof r.v);
}

extern "C" {
void hc_hostfp_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  switch (field) {
    case 0: hostfp_op<BlsFq>(op, a, b, out); break;
    case 1: hostfp_op<BlsFr>(op, a, b, out); break;
    case 2: hostfp_op<BnFq>(op, a, b, out); break;
    default: hostfp_op<BnFr>(op, a, b, out);
  }
}
int hc_hostfp_inv(int field, const uint32_t* a, uint32_t* out) {
  switch (field) {
    case 0: return hostfp_inv<BlsFq>(a, out);
    case 1: return hostfp_inv<BlsFr>(a, out);
    case 2: return hostfp_inv<BnFq>(a, out);
    default: return hostfp_inv<BnFr>(a, out);
  }
}
// (a*b + c*d) / R mod p, curve fields only (field 0 = BLS12-381 Fq, 2 = BN254 Fq)
void hc_fp_mul2(int field, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* out) {
  if (field == 0) fp_mul2<Fp<BlsFq>>(a, b, c, d, out);
  else fp_mul2<Fp<BnFq>>(a, b, c, d, out);
}
void hc_fp_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  switch (field) {
    case 0: fp_op<Fp<BlsFq>>(op, a, b, out); break;
    case 1: fp_op<Fp<BlsFr>>(op, a, b, out); break;
    case 2: fp_op<Fp<BnFq>>(op, a, b, out); break;
    case 3: fp_op<Fp<BnFr>>(op, a, b, out); break;
  }
}
void hc_ec_op(int curve, int op, const uint32_t* p, const uint32_t* q, uint32_t k, uint32_t* out) {
  if (curve == 0) ec_op<Fp<BlsFq>>(op, p, q, k, out);
  else ec_op<Fp<BnFq>>(op, p, q, k, out);
}
}
