// Per-proof logic of the BATCHED R1CS verifier as host+device functions (kernels in verifybatch.cu are loops around
// these; tests/host_check2.cpp runs the same functions on the CPU against the oracle).
//
// What Verifier::verify computes between reading a proof and its single MSM (/root/reference/src/r1cs/verifier.rs:267-449):
//   * the transcript replay that yields y, z, u, x, w (verifier.rs:279-323) and the IPP challenges u_k (ipp.rs:278-288);
//   * flattened_constraints (verifier.rs:149-193) -- here a sparse matrix in CSR form shared by every proof of the batch
//     (one circuit, many proofs), multiplied by the powers of z on the device;
//   * verification_scalars (ipp.rs:295-312): one inversion, u_k^2, u_k^-2, s;
//   * y^-i, y_inv_wR, delta, g_scalars, h_scalars (verifier.rs:341-390) and the 13 + m head scalars (verifier.rs:392-429).
#pragma once
#include "curves.cuh"
#include "merlin.cuh"

namespace bp {

// One-phase constraint system in CSR form, rows = variables: [wL 0..n) | wR 0..n) | wO 0..n) | wV 0..m) | wc].
// Entry e of a row: constraint index q (the term is coeff * z^(q+1)) and its coefficient with the sign conventions of
// verifier.rs:166-190 already applied (committed and constant terms negated).  Coefficients +1 / -1 (nine in ten in the
// range gadgets) are flagged in ent_q and skip the product.
struct CircuitDev {
  uint32_t n, m, q, nnz;
  const uint32_t* row_start;   // 3n + m + 2 offsets
  const uint32_t* ent_q;       // bit 31: coeff = +1, bit 30: coeff = -1, low 30 bits: q
  const void* ent_c;           // Fr[nnz], Montgomery form
};
static const uint32_t CSR_PLUS = 0x80000000u, CSR_MINUS = 0x40000000u, CSR_QMASK = 0x3fffffffu;

// pw[k] = x^(2^k): x^e with popcount(e) products
template <class Fr>
BP_HD Fr hd_pow_tab(const Fr* pw, uint32_t e) {
  Fr acc = Fr::one();
  bool first = true;
  for (int k = 0; e; k++, e >>= 1)
    if (e & 1) { acc = first ? pw[k] : acc * pw[k]; first = false; }
  return acc;
}

// flat wire form of an R1CSProof (host/r1cs.hpp R1CSProof::to_bytes; proof.rs:26-58 field order)
template <class Curve>
struct ProofLayout {
  static constexpr uint32_t MB = Curve::MODBYTES, PB = 1 + 2 * MB;
  BP_HD static uint32_t point(uint32_t k) { return k * PB; }                       // A_I1 A_O1 S1 A_I2 A_O2 S2 T_1 T_3 T_4 T_5 T_6
  BP_HD static uint32_t scalar(uint32_t k) { return 11 * PB + k * MB; }            // t_x t_x_blinding e_blinding
  BP_HD static uint32_t L(uint32_t k) { return 11 * PB + 3 * MB + k * PB; }
  BP_HD static uint32_t R(uint32_t lg, uint32_t k) { return 11 * PB + 3 * MB + (lg + k) * PB; }
  BP_HD static uint32_t a(uint32_t lg) { return 11 * PB + 3 * MB + 2 * lg * PB; }
  BP_HD static uint32_t len(uint32_t lg) { return a(lg) + 2 * MB; }
};

// per-proof derived scalars (Montgomery form): VB_HDR header entries, then u_k^2 [lg], then u_k^-2 [lg]
enum { VB_YINV = 0, VB_Z, VB_U, VB_X, VB_W, VB_R, VB_A, VB_B, VB_TX, VB_TXB, VB_EB, VB_S0, VB_Y, VB_HDR };
BP_HD uint32_t vb_hdr_len(uint32_t lg) { return VB_HDR + 2 * lg; }

// challenge order of the host-transcript mode: y, z, u, x, w, then the lg IPP challenges
enum { VB_CH_Y = 0, VB_CH_Z, VB_CH_U, VB_CH_X, VB_CH_W, VB_CH_FIXED };

// FieldElement::random() of the verifier (verifier.rs:392) as draw `ctr` of the host layer's counter-mode stream
// (host/curve.hpp Rng): be_int(SHAKE256(key || le64(ctr))[..MODBYTES]) mod r
template <class Curve>
BP_HD typename Curve::Fr fr_stream_draw(const uint8_t* key, uint32_t klen, uint64_t ctr) {
  uint64_t st[25];
  for (int l = 0; l < 25; l++) st[l] = 0;
  uint32_t pos = 0;
  for (uint32_t k = 0; k < klen; k++, pos++) st[pos >> 3] ^= (uint64_t)key[k] << (8 * (pos & 7));
  for (int k = 0; k < 8; k++, pos++) st[pos >> 3] ^= (uint64_t)(uint8_t)(ctr >> (8 * k)) << (8 * (pos & 7));
  st[pos >> 3] ^= (uint64_t)0x1f << (8 * (pos & 7));      // SHAKE padding, rate 136
  st[16] ^= (uint64_t)0x80 << 56;
  keccak_f1600_hd(st);
  uint8_t ob[Curve::MODBYTES];
  for (int k = 0; k < Curve::MODBYTES; k++) ob[k] = (uint8_t)(st[k >> 3] >> (8 * (k & 7)));
  return fr_from_be_wide<Curve>(ob);
}

// Finish the header from the challenges: proof scalars (canonical or BPGPU_E_FORMAT), the shared inversion of
// FieldElement::batch_invert (ipp.rs:295; y^-1 rides along), u_k^2, u_k^-2, s[0] = prod u_k^-1.
template <class Curve>
BP_HD int vb_header_finish(const uint8_t* proof, uint32_t lg, const typename Curve::Fr& y, const typename Curve::Fr& z,
                           const typename Curve::Fr& u, const typename Curve::Fr& x, const typename Curve::Fr& w,
                           const typename Curve::Fr& r, const typename Curve::Fr* uk, typename Curve::Fr* hdr) {
  using Fr = typename Curve::Fr;
  using PL = ProofLayout<Curve>;
  const uint8_t* sc[5] = {proof + PL::a(lg), proof + PL::a(lg) + PL::MB, proof + PL::scalar(0), proof + PL::scalar(1), proof + PL::scalar(2)};
  for (int k = 0; k < 5; k++) {
    if (!fr_be_is_canonical<Curve>(sc[k])) return BPGPU_E_FORMAT;
    hdr[VB_A + k] = fr_from_be_wide<Curve>(sc[k]);
  }
  hdr[VB_Y] = y; hdr[VB_Z] = z; hdr[VB_U] = u; hdr[VB_X] = x; hdr[VB_W] = w; hdr[VB_R] = r;
  // prefix products y, y*u_0, ... ; one inversion; unwind (a zero factor zeroes every inverse, as inverting the zero
  // product does in the host layer)
  Fr pre[33];
  Fr run = y;
  for (uint32_t k = 0; k < lg; k++) { pre[k] = run; run = run * uk[k]; }
  Fr inv = run.inv();
  Fr s0 = Fr::one();
  for (uint32_t k = lg; k-- > 0;) {
    const Fr ui = inv * pre[k];
    inv = inv * uk[k];
    hdr[VB_HDR + k] = uk[k].sqr();
    hdr[VB_HDR + lg + k] = ui.sqr();
    s0 = s0 * ui;
  }
  hdr[VB_YINV] = inv;
  hdr[VB_S0] = s0;
  return BPGPU_OK;
}

// The transcript of one proof of a one-phase circuit in the four stages at which a PROVER can advance it (the verifier
// replays them back to back): identical on both sides (prover.rs:119-129,327,364-366,432-439,502-509,543-549; verifier.rs:
// 124-132,279-323; ipp.rs:62,106-113,172-179,278-288).  Points are read from the flat proof record / the commitment list.
template <class Curve>
BP_HD typename Curve::Fr tr_challenge(StrobeHD& t, const char* label, uint32_t label_len) {
  uint8_t buf[Curve::MODBYTES];
  t.challenge_bytes(label, label_len, buf, Curve::MODBYTES);
  return fr_from_be_wide<Curve>(buf);
}
// V_1..V_m, "m", A_I1 A_O1 S1, 1-phase separator, A_I2 A_O2 S2 -> y, z
template <class Curve>
BP_HD void tr_stage_commitments(StrobeHD& t, const uint8_t* proof, const uint8_t* comms_xy, uint32_t m, typename Curve::Fr* y,
                                typename Curve::Fr* z) {
  using PL = ProofLayout<Curve>;
  constexpr uint32_t MB = Curve::MODBYTES;
  for (uint32_t j = 0; j < m; j++) t.append_message(MRL_LIT("V"), comms_xy + (size_t)j * 2 * MB, 2 * MB, 4);
  t.append_u64(MRL_LIT("m"), m);
  t.append_message(MRL_LIT("A_I1"), proof + PL::point(0) + 1, 2 * MB, 4);
  t.append_message(MRL_LIT("A_O1"), proof + PL::point(1) + 1, 2 * MB, 4);
  t.append_message(MRL_LIT("S1"), proof + PL::point(2) + 1, 2 * MB, 4);
  t.append_message(MRL_LIT("dom-sep"), (const uint8_t*)"r1cs-1phase", 11);
  t.append_message(MRL_LIT("A_I2"), proof + PL::point(3) + 1, 2 * MB, 4);
  t.append_message(MRL_LIT("A_O2"), proof + PL::point(4) + 1, 2 * MB, 4);
  t.append_message(MRL_LIT("S2"), proof + PL::point(5) + 1, 2 * MB, 4);
  *y = tr_challenge<Curve>(t, MRL_LIT("y"));
  *z = tr_challenge<Curve>(t, MRL_LIT("z"));
}
// T_1 T_3 T_4 T_5 T_6 -> u, x
template <class Curve>
BP_HD void tr_stage_t(StrobeHD& t, const uint8_t* proof, typename Curve::Fr* u, typename Curve::Fr* x) {
  using PL = ProofLayout<Curve>;
  constexpr uint32_t MB = Curve::MODBYTES;
  t.append_message(MRL_LIT("T_1"), proof + PL::point(6) + 1, 2 * MB, 4);
  t.append_message(MRL_LIT("T_3"), proof + PL::point(7) + 1, 2 * MB, 4);
  t.append_message(MRL_LIT("T_4"), proof + PL::point(8) + 1, 2 * MB, 4);
  t.append_message(MRL_LIT("T_5"), proof + PL::point(9) + 1, 2 * MB, 4);
  t.append_message(MRL_LIT("T_6"), proof + PL::point(10) + 1, 2 * MB, 4);
  *u = tr_challenge<Curve>(t, MRL_LIT("u"));
  *x = tr_challenge<Curve>(t, MRL_LIT("x"));
}
// t_x, t_x_blinding, e_blinding -> w; then the inner-product argument's separator
template <class Curve>
BP_HD void tr_stage_scalars(StrobeHD& t, const uint8_t* proof, uint64_t N, typename Curve::Fr* w) {
  using PL = ProofLayout<Curve>;
  constexpr uint32_t MB = Curve::MODBYTES;
  t.append_message(MRL_LIT("t_x"), proof + PL::scalar(0), MB);
  t.append_message(MRL_LIT("t_x_blinding"), proof + PL::scalar(1), MB);
  t.append_message(MRL_LIT("e_blinding"), proof + PL::scalar(2), MB);
  *w = tr_challenge<Curve>(t, MRL_LIT("w"));
  t.append_message(MRL_LIT("dom-sep"), (const uint8_t*)"ipp v1", 6);
  t.append_u64(MRL_LIT("n"), N);
}
// L_k, R_k -> u_k
template <class Curve>
BP_HD typename Curve::Fr tr_round(StrobeHD& t, const uint8_t* proof, uint32_t lg, uint32_t k) {
  using PL = ProofLayout<Curve>;
  constexpr uint32_t MB = Curve::MODBYTES;
  t.append_message(MRL_LIT("L"), proof + PL::L(k) + 1, 2 * MB, 4);
  t.append_message(MRL_LIT("R"), proof + PL::R(lg, k) + 1, 2 * MB, 4);
  return tr_challenge<Curve>(t, MRL_LIT("u"));
}

// The verifier's replay for one proof, starting from the exported state after Transcript::new(label) and
// r1cs_domain_sep().  Writes the header; returns BPGPU_OK or BPGPU_E_FORMAT (a point without the 0x04 tag, a non-canonical
// scalar).
template <class Curve>
BP_HD int vb_replay(const uint8_t* state0, const uint8_t* proof, const uint8_t* comms_xy, uint32_t m, uint32_t lg, uint64_t N,
                    const typename Curve::Fr& r, typename Curve::Fr* hdr) {
  using Fr = typename Curve::Fr;
  using PL = ProofLayout<Curve>;
  for (uint32_t k = 0; k < 11 + 2 * lg; k++) {
    const uint32_t off = k < 11 ? PL::point(k) : PL::L(k - 11);
    if (proof[off] != 4) return BPGPU_E_FORMAT;
  }
  StrobeHD t;
  t.load(state0);
  Fr y, z, u, x, w;
  tr_stage_commitments<Curve>(t, proof, comms_xy, m, &y, &z);
  tr_stage_t<Curve>(t, proof, &u, &x);
  tr_stage_scalars<Curve>(t, proof, N, &w);
  Fr uk[32];
  for (uint32_t k = 0; k < lg; k++) uk[k] = tr_round<Curve>(t, proof, lg, k);
  return vb_header_finish<Curve>(proof, lg, y, z, u, x, w, r, uk, hdr);
}

// the same header from challenges the caller's own transcripts produced (count x (5 + lg) big-endian scalars)
template <class Curve>
BP_HD int vb_from_challenges(const uint8_t* chal_be, const uint8_t* proof, uint32_t lg, const typename Curve::Fr& r,
                             typename Curve::Fr* hdr) {
  using Fr = typename Curve::Fr;
  using PL = ProofLayout<Curve>;
  constexpr uint32_t MB = Curve::MODBYTES;
  for (uint32_t k = 0; k < 11 + 2 * lg; k++) {
    const uint32_t off = k < 11 ? PL::point(k) : PL::L(k - 11);
    if (proof[off] != 4) return BPGPU_E_FORMAT;
  }
  const Fr y = fr_from_be_wide<Curve>(chal_be + (size_t)VB_CH_Y * MB);
  const Fr z = fr_from_be_wide<Curve>(chal_be + (size_t)VB_CH_Z * MB);
  const Fr u = fr_from_be_wide<Curve>(chal_be + (size_t)VB_CH_U * MB);
  const Fr x = fr_from_be_wide<Curve>(chal_be + (size_t)VB_CH_X * MB);
  const Fr w = fr_from_be_wide<Curve>(chal_be + (size_t)VB_CH_W * MB);
  Fr uk[32];
  for (uint32_t k = 0; k < lg; k++) uk[k] = fr_from_be_wide<Curve>(chal_be + (size_t)(VB_CH_FIXED + k) * MB);
  return vb_header_finish<Curve>(proof, lg, y, z, u, x, w, r, uk, hdr);
}

// entries [lo, hi) of the flattened constraint matrices times the powers of z, every `step`-th one: sum_e coeff_e * z^(q_e + 1)
template <class Fr>
BP_HD Fr csr_span_eval(const CircuitDev& c, uint32_t lo, uint32_t hi, uint32_t step, const Fr* ztab) {
  Fr acc = Fr::zero();
  const Fr* ec = (const Fr*)c.ent_c;
  for (uint32_t e = lo; e < hi; e += step) {
    const uint32_t q = c.ent_q[e];
    const Fr zp = hd_pow_tab(ztab, (q & CSR_QMASK) + 1);
    if (q & CSR_PLUS) acc = acc + zp;
    else if (q & CSR_MINUS) acc = acc - zp;
    else acc = acc + ec[e] * zp;
  }
  return acc;
}
// a CONTIGUOUS run of entries [lo, hi): within a row the constraint indices ascend, so z^(q+1) is carried from entry to
// entry with a short product z^(gap) instead of a fresh power (the constants row of a range statement: gaps of 2 or 3)
template <class Fr>
BP_HD Fr csr_chunk_eval(const CircuitDev& c, uint32_t lo, uint32_t hi, const Fr* ztab) {
  Fr acc = Fr::zero();
  if (lo >= hi) return acc;
  const Fr* ec = (const Fr*)c.ent_c;
  uint32_t qp = c.ent_q[lo] & CSR_QMASK;
  Fr zp = hd_pow_tab(ztab, qp + 1);
  for (uint32_t e = lo; e < hi; e++) {
    const uint32_t q = c.ent_q[e], qi = q & CSR_QMASK;
    if (qi > qp) zp = zp * hd_pow_tab(ztab, qi - qp);
    else if (qi < qp) zp = hd_pow_tab(ztab, qi + 1);
    qp = qi;
    if (q & CSR_PLUS) acc = acc + zp;
    else if (q & CSR_MINUS) acc = acc - zp;
    else acc = acc + ec[e] * zp;
  }
  return acc;
}
// one row (a variable's weight; the LAST row, the constants, has one entry per constraint with a constant: kernels split it
// over a block with csr_span_eval)
template <class Fr>
BP_HD Fr csr_row_eval(const CircuitDev& c, uint32_t row, const Fr* ztab) {
  return csr_span_eval(c, c.row_start[row], c.row_start[row + 1], 1u, ztab);
}

// x^(2^k), k < 32
template <class Fr>
BP_HD void hd_square_table(const Fr& x, Fr* tab, int count = 32) {
  Fr cur = x;
  for (int k = 0; k < count; k++) { tab[k] = cur; cur = cur.sqr(); }
}

// verifier.rs:341-390 for index i of one proof (one-phase circuit: n1 = n): g_scalars[i], h_scalars[i] and this index's
// term of delta = <y^-n o wR, wL>.  yitab = (y^-1)^(2^k), ztab = z^(2^k).
template <class Fr>
BP_HD void vb_gh_element(const CircuitDev& c, uint32_t i, uint32_t N, uint32_t lg, const Fr* hdr, const Fr* yitab, const Fr* ztab, Fr* gs,
                         Fr* hs, Fr* delta_term) {
  const uint32_t n = c.n;
  const Fr* usq = hdr + VB_HDR;
  // s[i] = s[0] * prod_{bit j of i set} u^2_{lg-1-j}   (ipp.rs:303-312);  s[N-1-i] takes the complementary bits
  Fr si = hdr[VB_S0], sj = hdr[VB_S0];
  for (uint32_t j = 0; j < lg; j++) {
    if ((i >> j) & 1) si = si * usq[lg - 1 - j];
    else sj = sj * usq[lg - 1 - j];
  }
  const Fr yinv_i = hd_pow_tab(yitab, i);
  Fr wl = Fr::zero(), wr = Fr::zero(), wo = Fr::zero();
  if (i < n) { wl = csr_row_eval(c, i, ztab); wr = csr_row_eval(c, n + i, ztab); wo = csr_row_eval(c, 2 * n + i, ztab); }
  const Fr yw = wr * yinv_i;
  *delta_term = i < n ? yw * wl : Fr::zero();
  Fr g = hdr[VB_X] * yw - hdr[VB_A] * si;
  Fr h = yinv_i * (hdr[VB_X] * wl + wo - hdr[VB_B] * sj) - Fr::one();
  if (i >= n) { g = hdr[VB_U] * g; h = hdr[VB_U] * h; }
  *gs = g; *hs = h;
  (void)N;
}

// verifier.rs:392-429: the scalars of g and h and of the proof's own points, given delta and wc.  var = 6 + m + 5 + 2 lg
// scalars in the order [A_I1 A_O1 S1 A_I2 A_O2 S2 | V_j | T_1 T_3 T_4 T_5 T_6 | L_k | R_k]; the m entries of V are written
// by vb_var_wv.  All outputs are CANONICAL integers (the MSM kernels take digits).
template <class Fr>
BP_HD void vb_head(uint32_t m, uint32_t lg, const Fr* hdr, const Fr& delta, const Fr& wc, Fr* fixed_g, Fr* fixed_h, Fr* var) {
  const Fr x = hdr[VB_X], u = hdr[VB_U], rnd = hdr[VB_R];
  const Fr xx = x.sqr(), xxx = x * xx;
  const Fr rx = rnd * x, rx3 = rnd * xxx, rx4 = rx3 * x, rx5 = rx4 * x, rx6 = rx5 * x;
  *fixed_g = (hdr[VB_W] * (hdr[VB_TX] - hdr[VB_A] * hdr[VB_B]) + rnd * (xx * (wc + delta) - hdr[VB_TX])).from_mont();   // :421
  *fixed_h = (hdr[VB_EB] + rnd * hdr[VB_TXB]).neg().from_mont();                                                        // :424
  var[0] = x.from_mont(); var[1] = xx.from_mont(); var[2] = xxx.from_mont();
  var[3] = (u * x).from_mont(); var[4] = (u * xx).from_mont(); var[5] = (u * xxx).from_mont();
  Fr* t = var + 6 + m;
  t[0] = rx.from_mont(); t[1] = rx3.from_mont(); t[2] = rx4.from_mont(); t[3] = rx5.from_mont(); t[4] = rx6.from_mont();
  for (uint32_t k = 0; k < 2 * lg; k++) t[5 + k] = hdr[VB_HDR + k].from_mont();
}
// scalar of V_j: wV[j] * r * x^2 (verifier.rs:416-418)
template <class Fr>
BP_HD Fr vb_var_wv(const CircuitDev& c, uint32_t j, const Fr* hdr, const Fr* ztab) {
  const Fr r_xx = hdr[VB_R] * hdr[VB_X].sqr();
  return (csr_row_eval(c, 3 * c.n + j, ztab) * r_xx).from_mont();
}

// which proof byte range holds variable point k (k < 6 + m + 5 + 2 lg), or the commitment index for the V block
template <class Curve>
BP_HD const uint8_t* vb_var_point_bytes(const uint8_t* proof, const uint8_t* comms_xy, uint32_t m, uint32_t lg, uint32_t k) {
  using PL = ProofLayout<Curve>;
  constexpr uint32_t MB = Curve::MODBYTES;
  if (k < 6) return proof + PL::point(k) + 1;
  if (k < 6 + m) return comms_xy + (size_t)(k - 6) * 2 * MB;
  k -= 6 + m;
  if (k < 5) return proof + PL::point(6 + k) + 1;
  k -= 5;
  return proof + PL::L(k) + 1;            // L_0.. then R_0.. are contiguous
}

}  // namespace bp

// bpgpu_circuit (include/bpgpu.h): a one-phase constraint system resident on one device
struct bpgpu_ctx;
struct bpgpu_circuit {
  bpgpu_ctx* ctx;
  bp::CircuitDev dev;
  void* mem;
};
