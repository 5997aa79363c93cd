/* Oracle (test infrastructure): Montgomery prime field on NL 64-bit limbs, instantiated by
 * including this file with NL, FN(name) and the constant table defined.  Plain CIOS with
 * unsigned __int128 -- an independent implementation from the device's 32-bit even/odd schedule
 * (bulletproofs-amcl_b200/csrc/fp.cuh), which is the point of having it.
 * Restates AMCL's FP arithmetic as used via amcl_wrapper; results are canonical
 * residues so the algorithm differences are unobservable. */

typedef struct { uint64_t v[NL]; } FE;

static const FE FN(P) = {FP_P};
static const FE FN(R1) = {FP_R1};
static const FE FN(R2) = {FP_R2};

static inline int FN(is_zero)(const FE* a) { uint64_t o = 0; for (int i = 0; i < NL; i++) o |= a->v[i]; return o == 0; }
static inline int FN(eq)(const FE* a, const FE* b) { uint64_t o = 0; for (int i = 0; i < NL; i++) o |= a->v[i] ^ b->v[i]; return o == 0; }
static inline int FN(geq_p)(const uint64_t* a) {
  for (int i = NL - 1; i >= 0; i--) { if (a[i] > FN(P).v[i]) return 1; if (a[i] < FN(P).v[i]) return 0; }
  return 1;
}
static inline void FN(sub_p)(uint64_t* a) {
  unsigned __int128 br = 0;
  for (int i = 0; i < NL; i++) { unsigned __int128 t = (unsigned __int128)a[i] - FN(P).v[i] - (uint64_t)br; a[i] = (uint64_t)t; br = (t >> 64) & 1; }
}
static inline void FN(add)(FE* r, const FE* a, const FE* b) {
  unsigned __int128 c = 0;
  for (int i = 0; i < NL; i++) { c += (unsigned __int128)a->v[i] + b->v[i]; r->v[i] = (uint64_t)c; c >>= 64; }
  if (c || FN(geq_p)(r->v)) FN(sub_p)(r->v);
}
static inline void FN(sub)(FE* r, const FE* a, const FE* b) {
  unsigned __int128 br = 0;
  for (int i = 0; i < NL; i++) { unsigned __int128 t = (unsigned __int128)a->v[i] - b->v[i] - (uint64_t)br; r->v[i] = (uint64_t)t; br = (t >> 64) & 1; }
  if (br) { unsigned __int128 c = 0; for (int i = 0; i < NL; i++) { c += (unsigned __int128)r->v[i] + FN(P).v[i]; r->v[i] = (uint64_t)c; c >>= 64; } }
}
static inline void FN(neg)(FE* r, const FE* a) { FE z; memset(&z, 0, sizeof z); FN(sub)(r, &z, a); }
static inline void FN(mul)(FE* r, const FE* a, const FE* b) {
  uint64_t t[NL + 2];
  memset(t, 0, sizeof t);
  for (int i = 0; i < NL; i++) {
    unsigned __int128 c = 0;
    for (int j = 0; j < NL; j++) { c += (unsigned __int128)a->v[j] * b->v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[NL]; t[NL] = (uint64_t)c; t[NL + 1] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * FP_INV;
    c = (unsigned __int128)m * FN(P).v[0] + t[0]; c >>= 64;
    for (int j = 1; j < NL; j++) { c += (unsigned __int128)m * FN(P).v[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[NL]; t[NL - 1] = (uint64_t)c; t[NL] = t[NL + 1] + (uint64_t)(c >> 64);
  }
  if (t[NL] || FN(geq_p)(t)) FN(sub_p)(t);
  memcpy(r->v, t, sizeof r->v);
}
static inline void FN(sqr)(FE* r, const FE* a) { FN(mul)(r, a, a); }
static inline void FN(to_mont)(FE* r, const FE* a) { FN(mul)(r, a, &FN(R2)); }
static inline void FN(from_mont)(FE* r, const FE* a) { FE one; memset(&one, 0, sizeof one); one.v[0] = 1; FN(mul)(r, a, &one); }
/* x^e, e = little-endian limbs */
static void FN(pow)(FE* r, const FE* x, const uint64_t* e, int nl) {
  FE acc = FN(R1);
  for (int i = nl - 1; i >= 0; i--)
    for (int b = 63; b >= 0; b--) { FN(sqr)(&acc, &acc); if ((e[i] >> b) & 1) FN(mul)(&acc, &acc, x); }
  *r = acc;
}
/* Fermat inverse, 0 -> 0 (AMCL's FieldElement::inverse of zero is zero) */
static void FN(inv)(FE* r, const FE* x) {
  uint64_t e[NL];
  memcpy(e, FN(P).v, sizeof e);
  e[0] -= 2;  /* p is odd and > 2: no borrow */
  FN(pow)(r, x, e, NL);
}
/* big-endian bytes (nbytes >= 8*NL allowed: leading bytes ignored) -> canonical -> Montgomery */
static void FN(from_be)(FE* r, const uint8_t* be, int nbytes) {
  FE t;
  for (int i = 0; i < NL; i++) {
    uint64_t w = 0;
    const uint8_t* p = be + nbytes - 8 * (i + 1);
    for (int k = 0; k < 8; k++) w = (w << 8) | p[k];
    t.v[i] = w;
  }
  while (FN(geq_p)(t.v)) FN(sub_p)(t.v);
  FN(to_mont)(r, &t);
}
static void FN(to_be)(const FE* a, uint8_t* be, int nbytes) {
  FE t;
  FN(from_mont)(&t, a);
  memset(be, 0, nbytes);
  for (int i = 0; i < NL; i++) {
    uint8_t* p = be + nbytes - 8 * (i + 1);
    for (int k = 0; k < 8; k++) p[k] = (uint8_t)(t.v[i] >> (56 - 8 * k));
  }
}
