/* Oracle (test infrastructure): G1 arithmetic (Jacobian, a = 0) + the reference's CPU MSM
 * algorithm, instantiated per curve by defining CN(name), Q(name) (base-field ops), QFE,
 * MODBYTES_, and CURVE_B3 is unused (a = 0 formulas do not need b).
 *
 * Restates what amcl_wrapper's G1 / G1Vector do for the reference's call sites
 * (/root/reference/src/ipp.rs:91-104,119-129,158-170,185-187,251-253; src/r1cs/verifier.rs:451):
 *   - `multi_scalar_mul_var_time`: Straus interleaving with width-5 wNAF digits and an
 *     8-entry table of odd multiples per point, shared doublings [amcl_wrapper, RECALLED];
 *   - `binary_scalar_mul`: two-scalar multiplication (same interleaving with n = 2).
 */

typedef struct { QFE X, Y, Z; } CN(jac);     /* Z == 0: identity */

static void CN(set_inf)(CN(jac)* p) { memset(p, 0, sizeof *p); p->Y = Q(R1); }
static int CN(is_inf)(const CN(jac)* p) { return Q(is_zero)(&p->Z); }

static void CN(dbl)(CN(jac)* r, const CN(jac)* p) {
  if (CN(is_inf)(p) || Q(is_zero)(&p->Y)) { CN(set_inf)(r); return; }
  QFE A, B, C, D, E, F, t, X3, Y3, Z3;
  Q(sqr)(&A, &p->X); Q(sqr)(&B, &p->Y); Q(sqr)(&C, &B);
  Q(add)(&t, &p->X, &B); Q(sqr)(&t, &t); Q(sub)(&t, &t, &A); Q(sub)(&t, &t, &C); Q(add)(&D, &t, &t);
  Q(add)(&E, &A, &A); Q(add)(&E, &E, &A);
  Q(sqr)(&F, &E);
  Q(sub)(&X3, &F, &D); Q(sub)(&X3, &X3, &D);
  Q(sub)(&t, &D, &X3); Q(mul)(&Y3, &E, &t);
  Q(add)(&C, &C, &C); Q(add)(&C, &C, &C); Q(add)(&C, &C, &C); Q(sub)(&Y3, &Y3, &C);
  Q(mul)(&Z3, &p->Y, &p->Z); Q(add)(&Z3, &Z3, &Z3);
  r->X = X3; r->Y = Y3; r->Z = Z3;
}

static void CN(add)(CN(jac)* r, const CN(jac)* p, const CN(jac)* q) {
  if (CN(is_inf)(p)) { *r = *q; return; }
  if (CN(is_inf)(q)) { *r = *p; return; }
  QFE Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t, X3, Y3, Z3;
  Q(sqr)(&Z1Z1, &p->Z); Q(sqr)(&Z2Z2, &q->Z);
  Q(mul)(&U1, &p->X, &Z2Z2); Q(mul)(&U2, &q->X, &Z1Z1);
  Q(mul)(&S1, &p->Y, &q->Z); Q(mul)(&S1, &S1, &Z2Z2);
  Q(mul)(&S2, &q->Y, &p->Z); Q(mul)(&S2, &S2, &Z1Z1);
  if (Q(eq)(&U1, &U2)) { if (Q(eq)(&S1, &S2)) CN(dbl)(r, p); else CN(set_inf)(r); return; }
  Q(sub)(&H, &U2, &U1);
  Q(add)(&I, &H, &H); Q(sqr)(&I, &I);
  Q(mul)(&J, &H, &I);
  Q(sub)(&rr, &S2, &S1); Q(add)(&rr, &rr, &rr);
  Q(mul)(&V, &U1, &I);
  Q(sqr)(&X3, &rr); Q(sub)(&X3, &X3, &J); Q(sub)(&X3, &X3, &V); Q(sub)(&X3, &X3, &V);
  Q(sub)(&t, &V, &X3); Q(mul)(&Y3, &rr, &t);
  Q(mul)(&t, &S1, &J); Q(add)(&t, &t, &t); Q(sub)(&Y3, &Y3, &t);
  Q(add)(&Z3, &p->Z, &q->Z); Q(sqr)(&Z3, &Z3); Q(sub)(&Z3, &Z3, &Z1Z1); Q(sub)(&Z3, &Z3, &Z2Z2); Q(mul)(&Z3, &Z3, &H);
  r->X = X3; r->Y = Y3; r->Z = Z3;
}

static void CN(neg)(CN(jac)* r, const CN(jac)* p) { *r = *p; Q(neg)(&r->Y, &p->Y); }

static void CN(to_affine)(const CN(jac)* p, QFE* x, QFE* y, int* inf) {
  if (CN(is_inf)(p)) { *inf = 1; return; }
  *inf = 0;
  QFE zi, zi2, zi3;
  Q(inv)(&zi, &p->Z); Q(sqr)(&zi2, &zi); Q(mul)(&zi3, &zi2, &zi);
  Q(mul)(x, &p->X, &zi2); Q(mul)(y, &p->Y, &zi3);
}

/* ABI encoding: X||Y big endian, identity = (0,1) (include/bpgpu.h) */
static void CN(from_xy)(CN(jac)* p, const uint8_t* xy) {
  int zero_x = 1, one_y = 1;
  for (int i = 0; i < MODBYTES_; i++) if (xy[i]) zero_x = 0;
  for (int i = 0; i < MODBYTES_ - 1; i++) if (xy[MODBYTES_ + i]) one_y = 0;
  if (xy[2 * MODBYTES_ - 1] != 1) one_y = 0;
  if (zero_x && one_y) { CN(set_inf)(p); return; }
  Q(from_be)(&p->X, xy, MODBYTES_); Q(from_be)(&p->Y, xy + MODBYTES_, MODBYTES_); p->Z = Q(R1);
}
static void CN(to_xy)(const CN(jac)* p, uint8_t* xy) {
  QFE x, y; int inf;
  CN(to_affine)(p, &x, &y, &inf);
  if (inf) { memset(xy, 0, 2 * MODBYTES_); xy[2 * MODBYTES_ - 1] = 1; return; }
  Q(to_be)(&x, xy, MODBYTES_); Q(to_be)(&y, xy + MODBYTES_, MODBYTES_);
}

/* width-5 wNAF of a 256-bit integer (4 LE limbs): digits in {0, +-1, +-3, ..., +-15}; returns length */
static int CN(wnaf5)(const uint64_t k_in[4], int8_t* naf /* >= 258 */) {
  uint64_t k[5] = {k_in[0], k_in[1], k_in[2], k_in[3], 0};
  int len = 0;
  for (;;) {
    if (!(k[0] | k[1] | k[2] | k[3] | k[4])) break;
    int d = 0;
    if (k[0] & 1) {
      d = (int)(k[0] & 31);
      if (d >= 16) d -= 32;
      /* k -= d */
      if (d > 0) { uint64_t b = (uint64_t)d; for (int i = 0; i < 5 && b; i++) { uint64_t o = k[i]; k[i] -= b; b = o < b; } }
      else { uint64_t c = (uint64_t)(-d); for (int i = 0; i < 5 && c; i++) { k[i] += c; c = k[i] < c; } }
    }
    naf[len++] = (int8_t)d;
    for (int i = 0; i < 4; i++) k[i] = (k[i] >> 1) | (k[i + 1] << 63);
    k[4] >>= 1;
  }
  return len;
}

/* Straus / interleaved wNAF-5 over points[lo..hi) ; scalars are canonical integers (4 LE limbs) */
static void CN(straus)(const CN(jac)* pts, const uint64_t (*sc)[4], size_t n, CN(jac)* out) {
  CN(jac)* tab = (CN(jac)*)malloc(n * 8 * sizeof(CN(jac)));
  int8_t* naf = (int8_t*)calloc(n, 260);
  int maxlen = 0;
  for (size_t i = 0; i < n; i++) {
    CN(jac) d2;
    tab[8 * i] = pts[i];
    CN(dbl)(&d2, &pts[i]);
    for (int j = 1; j < 8; j++) CN(add)(&tab[8 * i + j], &tab[8 * i + j - 1], &d2);
    int l = CN(wnaf5)(sc[i], naf + 260 * i);
    if (l > maxlen) maxlen = l;
  }
  CN(jac) acc; CN(set_inf)(&acc);
  for (int b = maxlen - 1; b >= 0; b--) {
    CN(dbl)(&acc, &acc);
    for (size_t i = 0; i < n; i++) {
      int d = naf[260 * i + b];
      if (d > 0) CN(add)(&acc, &acc, &tab[8 * i + (d >> 1)]);
      else if (d < 0) { CN(jac) t; CN(neg)(&t, &tab[8 * i + ((-d) >> 1)]); CN(add)(&acc, &acc, &t); }
    }
  }
  *out = acc;
  free(tab); free(naf);
}
