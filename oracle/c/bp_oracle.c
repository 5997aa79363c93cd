/* bp_oracle.c -- CPU oracle in plain C (TEST INFRASTRUCTURE; see oracle/__init__.py).
 *
 * A restatement of the reference's CPU algorithms for its hot path, used (a) to check the CUDA
 * path at sizes the Python oracle cannot reach and (b) as the timed CPU baseline of bench.py:
 *   - orc_msm            G1Vector::multi_scalar_mul_var_time / inner_product_var_time_with_ref_vecs
 *                        = Straus interleaving, wNAF width 5 (amcl_wrapper [RECALLED]); call sites
 *                        /root/reference/src/ipp.rs:91,104,158,170,251-253, src/r1cs/verifier.rs:451
 *   - orc_scalar_mul     `&G1 * &FieldElement` (prover.rs:358,550)
 *   - orc_binary_scalar_mul  G1::binary_scalar_mul (ipp.rs:119-129,185-187; prover.rs:123)
 *   - orc_multiples      synthetic structured inputs P_i = (i+1)*B (SURVEY.md 8d "family S")
 * The reference is single threaded; `threads` > 1 chunks the points over pthreads (each chunk
 * runs its own Straus, partial sums are added) so the baseline can also be quoted on all cores.
 * PARITY: unpinned by the reference (no golden vectors upstream); cross-checked against
 * oracle/curves.py in tests/test_oracle_c.py.
 * Byte formats = include/bpgpu.h: scalars big-endian MODBYTES, points X||Y big-endian.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "consts.h"

/* ---- BLS12-381 Fq (6 limbs) */
#define NL 6
#define FE fe_blsq
#define FN(x) blsq_##x
#define FP_P BLSFQ_P
#define FP_R1 BLSFQ_R1
#define FP_R2 BLSFQ_R2
#define FP_INV BLSFQ_INV
#include "field_impl.h"
#undef NL
#undef FE
#undef FN
#undef FP_P
#undef FP_R1
#undef FP_R2
#undef FP_INV
/* ---- BLS12-381 Fr (4 limbs) */
#define NL 4
#define FE fe_blsr
#define FN(x) blsr_##x
#define FP_P BLSFR_P
#define FP_R1 BLSFR_R1
#define FP_R2 BLSFR_R2
#define FP_INV BLSFR_INV
#include "field_impl.h"
#undef NL
#undef FE
#undef FN
#undef FP_P
#undef FP_R1
#undef FP_R2
#undef FP_INV
/* ---- BN254 Fq */
#define NL 4
#define FE fe_bnq
#define FN(x) bnq_##x
#define FP_P BNFQ_P
#define FP_R1 BNFQ_R1
#define FP_R2 BNFQ_R2
#define FP_INV BNFQ_INV
#include "field_impl.h"
#undef NL
#undef FE
#undef FN
#undef FP_P
#undef FP_R1
#undef FP_R2
#undef FP_INV
/* ---- BN254 Fr */
#define NL 4
#define FE fe_bnr
#define FN(x) bnr_##x
#define FP_P BNFR_P
#define FP_R1 BNFR_R1
#define FP_R2 BNFR_R2
#define FP_INV BNFR_INV
#include "field_impl.h"
#undef NL
#undef FE
#undef FN
#undef FP_P
#undef FP_R1
#undef FP_R2
#undef FP_INV

/* ---- curves */
#define CN(x) bls_##x
#define Q(x) blsq_##x
#define QFE fe_blsq
#define MODBYTES_ 48
#include "curve_impl.h"
#undef CN
#undef Q
#undef QFE
#undef MODBYTES_
#define CN(x) bn_##x
#define Q(x) bnq_##x
#define QFE fe_bnq
#define MODBYTES_ 32
#include "curve_impl.h"
#undef CN
#undef Q
#undef QFE
#undef MODBYTES_

static void scalar_from_be(const uint8_t* be, int modbytes, uint64_t out[4]) {
  for (int i = 0; i < 4; i++) {
    uint64_t w = 0;
    const uint8_t* p = be + modbytes - 8 * (i + 1);
    for (int k = 0; k < 8; k++) w = (w << 8) | p[k];
    out[i] = w;
  }
}

#define DEFINE_CURVE_API(C, MB)                                                                       \
  typedef struct { const C##_jac* pts; const uint64_t (*sc)[4]; size_t n; C##_jac out; } C##_job;      \
  static void* C##_worker(void* arg) { C##_job* j = (C##_job*)arg; C##_straus(j->pts, j->sc, j->n, &j->out); return 0; } \
  static int C##_msm(const uint8_t* pts_xy, const uint8_t* scalars_be, size_t n, int threads, uint8_t* out_xy) { \
    C##_jac* pts = (C##_jac*)malloc((n ? n : 1) * sizeof(C##_jac));                                   \
    uint64_t (*sc)[4] = (uint64_t (*)[4])malloc((n ? n : 1) * 32);                                    \
    for (size_t i = 0; i < n; i++) { C##_from_xy(&pts[i], pts_xy + i * 2 * MB); scalar_from_be(scalars_be + i * MB, MB, sc[i]); } \
    if (threads < 1) threads = 1;                                                                     \
    if ((size_t)threads > n) threads = n ? (int)n : 1;                                                \
    C##_job* jobs = (C##_job*)calloc(threads, sizeof(C##_job));                                       \
    pthread_t* th = (pthread_t*)calloc(threads, sizeof(pthread_t));                                   \
    size_t per = (n + threads - 1) / threads;                                                         \
    for (int t = 0; t < threads; t++) {                                                               \
      size_t lo = t * per, hi = lo + per > n ? n : lo + per;                                          \
      if (lo > n) lo = n;                                                                             \
      jobs[t].pts = pts + lo; jobs[t].sc = sc + lo; jobs[t].n = hi - lo;                              \
      if (threads > 1) pthread_create(&th[t], 0, C##_worker, &jobs[t]); else C##_worker(&jobs[t]);    \
    }                                                                                                 \
    C##_jac acc; C##_set_inf(&acc);                                                                   \
    for (int t = 0; t < threads; t++) { if (threads > 1) pthread_join(th[t], 0); C##_add(&acc, &acc, &jobs[t].out); } \
    C##_to_xy(&acc, out_xy);                                                                          \
    free(pts); free(sc); free(jobs); free(th);                                                        \
    return 0;                                                                                         \
  }                                                                                                   \
  static void C##_smul(const C##_jac* p, const uint64_t k[4], C##_jac* out) {                         \
    C##_jac r; C##_set_inf(&r);                                                                       \
    for (int i = 3; i >= 0; i--) for (int b = 63; b >= 0; b--) { C##_dbl(&r, &r); if ((k[i] >> b) & 1) C##_add(&r, &r, p); } \
    *out = r;                                                                                         \
  }                                                                                                   \
  static int C##_multiples(const uint8_t* base_xy, size_t n, uint8_t* out_xy) {                       \
    enum { CH = 512 };                                                                                \
    C##_jac b, cur; C##_from_xy(&b, base_xy); cur = b;                                                \
    C##_jac buf[CH]; C##_Q pre[CH];                                                                   \
    for (size_t i0 = 0; i0 < n; i0 += CH) {                                                           \
      size_t m = n - i0 < CH ? n - i0 : CH;                                                           \
      for (size_t i = 0; i < m; i++) { buf[i] = cur; C##_add(&cur, &cur, &b); }                       \
      /* Montgomery batch inversion of the Z coordinates (identity never occurs: B has prime order) */ \
      C##_Q acc = C##_QR1;                                                                            \
      for (size_t i = 0; i < m; i++) { pre[i] = acc; C##_qmul(&acc, &acc, &buf[i].Z); }               \
      C##_Q inv; C##_qinv(&inv, &acc);                                                                \
      for (size_t i = m; i-- > 0;) {                                                                  \
        C##_Q zi, zi2, zi3, x, y;                                                                     \
        C##_qmul(&zi, &inv, &pre[i]); C##_qmul(&inv, &inv, &buf[i].Z);                                \
        C##_qmul(&zi2, &zi, &zi); C##_qmul(&zi3, &zi2, &zi);                                          \
        C##_qmul(&x, &buf[i].X, &zi2); C##_qmul(&y, &buf[i].Y, &zi3);                                 \
        C##_qto_be(&x, out_xy + (i0 + i) * 2 * MB, MB); C##_qto_be(&y, out_xy + (i0 + i) * 2 * MB + MB, MB); \
      }                                                                                               \
    }                                                                                                 \
    return 0;                                                                                         \
  }

#define bls_Q fe_blsq
#define bls_QR1 blsq_R1
#define bls_qmul blsq_mul
#define bls_qinv blsq_inv
#define bls_qto_be blsq_to_be
#define bn_Q fe_bnq
#define bn_QR1 bnq_R1
#define bn_qmul bnq_mul
#define bn_qinv bnq_inv
#define bn_qto_be bnq_to_be
DEFINE_CURVE_API(bls, 48)
DEFINE_CURVE_API(bn, 32)

/* ------------------------------------------------------------------ exported API (curve: 0 BLS12-381, 1 BN254) */
int orc_msm(int curve, const uint8_t* pts_xy, const uint8_t* scalars_be, size_t n, int threads, uint8_t* out_xy) {
  return curve == 0 ? bls_msm(pts_xy, scalars_be, n, threads, out_xy) : bn_msm(pts_xy, scalars_be, n, threads, out_xy);
}

int orc_scalar_mul(int curve, const uint8_t* pt_xy, const uint8_t* scalar_be, uint8_t* out_xy) {
  uint64_t k[4];
  if (curve == 0) { bls_jac p, r; bls_from_xy(&p, pt_xy); scalar_from_be(scalar_be, 48, k); bls_smul(&p, k, &r); bls_to_xy(&r, out_xy); }
  else { bn_jac p, r; bn_from_xy(&p, pt_xy); scalar_from_be(scalar_be, 32, k); bn_smul(&p, k, &r); bn_to_xy(&r, out_xy); }
  return 0;
}

/* r1*g + r2*h by the same interleaved wNAF the MSM uses (G1::binary_scalar_mul) */
int orc_binary_scalar_mul(int curve, const uint8_t* g_xy, const uint8_t* h_xy, const uint8_t* r1_be, const uint8_t* r2_be,
                          uint8_t* out_xy) {
  int mb = curve == 0 ? 48 : 32;
  uint8_t pts[2 * 96], sc[2 * 48];
  memcpy(pts, g_xy, 2 * mb); memcpy(pts + 2 * mb, h_xy, 2 * mb);
  memcpy(sc, r1_be, mb); memcpy(sc + mb, r2_be, mb);
  return orc_msm(curve, pts, sc, 2, 1, out_xy);
}

/* out[i] = (i+1) * base, i < n : chained additions, each normalised (slow inversion per point is
 * avoided by batching: points are kept Jacobian and normalised with Montgomery's trick) */
int orc_multiples(int curve, const uint8_t* base_xy, size_t n, uint8_t* out_xy) {
  return curve == 0 ? bls_multiples(base_xy, n, out_xy) : bn_multiples(base_xy, n, out_xy);
}

/* Fr helpers for cross-checking the 64-bit field code against Python: op 0 mul 1 add 2 sub 3 inv */
int orc_fr_op(int curve, int op, const uint8_t* a_be, const uint8_t* b_be, uint8_t* out_be) {
  if (curve == 0) {
    fe_blsr a, b, r; blsr_from_be(&a, a_be, 48); blsr_from_be(&b, b_be, 48);
    if (op == 0) blsr_mul(&r, &a, &b); else if (op == 1) blsr_add(&r, &a, &b); else if (op == 2) blsr_sub(&r, &a, &b); else blsr_inv(&r, &a);
    blsr_to_be(&r, out_be, 48);
  } else {
    fe_bnr a, b, r; bnr_from_be(&a, a_be, 32); bnr_from_be(&b, b_be, 32);
    if (op == 0) bnr_mul(&r, &a, &b); else if (op == 1) bnr_add(&r, &a, &b); else if (op == 2) bnr_sub(&r, &a, &b); else bnr_inv(&r, &a);
    bnr_to_be(&r, out_be, 32);
  }
  return 0;
}

/* Keccak-f[1600] in place on a 200-byte little-endian state (FIPS 202): lets the Python Merlin transcript of the
 * CPU-baseline runs spend its time in the group operations, as the reference does, not in a Python permutation */
void orc_keccak_f1600(uint8_t* state) {
  static const uint64_t RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
                                  0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
                                  0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
                                  0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
                                  0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
                                  0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  uint64_t a[25], b[25], c[5], d[5];
  memcpy(a, state, 200);                       /* x86-64: little endian, as the state is defined */
  for (int r = 0; r < 24; r++) {
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ ((c[(x + 1) % 5] << 1) | (c[(x + 1) % 5] >> 63));
    for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) {
        const int i = x + 5 * y, rot = ROT[i];
        const uint64_t v = rot ? ((a[i] << rot) | (a[i] >> (64 - rot))) : a[i];
        b[y + 5 * ((2 * x + 3 * y) % 5)] = v;
      }
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= RC[r];
  }
  memcpy(state, a, 200);
}
