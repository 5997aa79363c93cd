// Host build of the per-proof device logic of the batched verifier / prover (csrc/merlin.cuh, csrc/verify_core.cuh,
// csrc/prove_core.cuh): the exact functions the kernels loop over, compiled with g++ so they can be checked against the
// oracle on a machine without a GPU.  Compiled by tests/test_host_core.py; NOT part of the product library.
#include <string.h>

#include <string>
#include <vector>

#include "../bulletproofs-amcl_b200/csrc/verify_core.cuh"
#include "../bulletproofs-amcl_b200/host/merlin.hpp"

using namespace bp;

// Merlin KAT through the device transcript code: Transcript::new(label) on the host implementation, exported, then
// append_message + challenge_bytes by StrobeHD
extern "C" int hc2_merlin_kat(const char* label, const char* msg_label, const uint8_t* msg, uint32_t msg_len, const char* ch_label,
                              uint8_t* out, uint32_t out_len) {
  bph::Transcript t{std::string(label)};
  uint8_t st[MERLIN_STATE_BYTES];
  t.export_state(st);
  StrobeHD s;
  s.load(st);
  s.append_message(msg_label, (uint32_t)strlen(msg_label), msg, msg_len);
  s.challenge_bytes(ch_label, (uint32_t)strlen(ch_label), out, out_len);
  return 0;
}

// exported state after Transcript::new(label) + r1cs_domain_sep(): where the verifier / prover replay starts
extern "C" void hc2_r1cs_state0(const char* label, uint8_t* out) {
  bph::Transcript t{std::string(label)};
  t.r1cs_domain_sep();
  t.export_state(out);
}

template <class Curve>
static int verify_terms_t(const uint8_t* state0, const uint8_t* chal_be, const uint8_t* proof, const uint8_t* comms, uint32_t m, uint32_t lg,
                          uint32_t n, uint32_t q, const uint32_t* row_start, const uint32_t* ent_q, const uint8_t* ent_c_be,
                          const uint8_t* key, uint32_t klen, uint64_t ctr, uint8_t* fixed_be, uint8_t* var_be, uint8_t* pts_ok) {
  using Fr = typename Curve::Fr;
  using Fq = typename Curve::Fq;
  constexpr int MB = Curve::MODBYTES;
  const uint32_t N = 1u << lg, nnz = row_start[3 * n + m + 1];
  std::vector<Fr> ec(nnz ? nnz : 1);
  for (uint32_t e = 0; e < nnz; e++) ec[e] = fr_from_be_wide<Curve>(ent_c_be + (size_t)e * MB);
  CircuitDev c{n, m, q, nnz, row_start, ent_q, ec.data()};
  std::vector<Fr> hdr(vb_hdr_len(lg));
  const Fr r = fr_stream_draw<Curve>(key, klen, ctr);
  int st = state0 ? vb_replay<Curve>(state0, proof, comms, m, lg, N, r, hdr.data())
                  : vb_from_challenges<Curve>(chal_be, proof, lg, r, hdr.data());
  if (st) return st;
  Fr yitab[32], ztab[32];
  hd_square_table(hdr[VB_YINV], yitab);
  hd_square_table(hdr[VB_Z], ztab);
  Fr delta = Fr::zero();
  std::vector<Fr> fixed(2 * N + 2), var(6 + m + 5 + 2 * lg);
  for (uint32_t i = 0; i < N; i++) {
    Fr g, h, d;
    vb_gh_element(c, i, N, lg, hdr.data(), yitab, ztab, &g, &h, &d);
    fixed[i] = g.from_mont();
    fixed[N + i] = h.from_mont();
    delta = delta + d;
  }
  const Fr wc = csr_row_eval(c, 3 * n + m, ztab);
  {                                                    // the kernels' split evaluations of the constants row agree with it
    const uint32_t lo = row_start[3 * n + m], hi = row_start[3 * n + m + 1];
    Fr strided = Fr::zero(), chunked = Fr::zero();
    for (uint32_t t = 0; t < 7; t++) strided = strided + csr_span_eval(c, lo + t, hi, 7u, ztab);
    const uint32_t per = (hi - lo + 4) / 5;
    for (uint32_t t = 0; t < 5; t++) {
      const uint32_t a = lo + t * per;
      if (a < hi) chunked = chunked + csr_chunk_eval(c, a, a + per < hi ? a + per : hi, ztab);
    }
    if (!(strided == wc) || !(chunked == wc)) return -100;
  }
  vb_head(m, lg, hdr.data(), delta, wc, &fixed[2 * N], &fixed[2 * N + 1], var.data());
  for (uint32_t j = 0; j < m; j++) var[6 + j] = vb_var_wv(c, j, hdr.data(), ztab);
  for (size_t i = 0; i < fixed.size(); i++) hd_limbs_to_be<8>(fixed[i].v, MB, fixed_be + i * MB);
  for (size_t i = 0; i < var.size(); i++) hd_limbs_to_be<8>(var[i].v, MB, var_be + i * MB);
  for (uint32_t k = 0; k < var.size(); k++) {
    Affine<Fq> a;
    pts_ok[k] = g1_from_be_checked<Curve>(vb_var_point_bytes<Curve>(proof, comms, m, lg, k), &a) ? 1 : 0;
  }
  return 0;
}

// the scalars of the verification MSM of ONE proof, as the batched device verifier builds them:
// fixed_be = (2N + 2) scalars of [G | H | g | h], var_be = 6 + m + 5 + 2 lg scalars of the proof's own points
extern "C" int hc2_verify_terms(int curve, const uint8_t* state0, const uint8_t* chal_be, const uint8_t* proof, const uint8_t* comms, uint32_t m,
                                uint32_t lg, uint32_t n, uint32_t q, const uint32_t* row_start, const uint32_t* ent_q, const uint8_t* ent_c_be,
                                const uint8_t* key, uint32_t klen, uint64_t ctr, uint8_t* fixed_be, uint8_t* var_be, uint8_t* pts_ok) {
  return curve == 0 ? verify_terms_t<Bls>(state0, chal_be, proof, comms, m, lg, n, q, row_start, ent_q, ent_c_be, key, klen, ctr, fixed_be, var_be, pts_ok)
                    : verify_terms_t<Bn>(state0, chal_be, proof, comms, m, lg, n, q, row_start, ent_q, ent_c_be, key, klen, ctr, fixed_be, var_be, pts_ok);
}

// point decoding as the ABI boundary does it: 1 valid (incl. the identity), 0 rejected
extern "C" int hc2_point_valid(int curve, const uint8_t* xy) {
  if (curve == 0) { Affine<Bls::Fq> a; return g1_from_be_checked<Bls>(xy, &a) ? 1 : 0; }
  Affine<Bn::Fq> a;
  return g1_from_be_checked<Bn>(xy, &a) ? 1 : 0;
}

// FieldElement::from(&[u8; MODBYTES]) as the device does it: out = canonical big-endian value mod r
extern "C" void hc2_fr_from_be_wide(int curve, const uint8_t* be, uint8_t* out) {
  if (curve == 0) { Bls::Fr v = fr_from_be_wide<Bls>(be).from_mont(); hd_limbs_to_be<8>(v.v, 48, out); }
  else { Bn::Fr v = fr_from_be_wide<Bn>(be).from_mont(); hd_limbs_to_be<8>(v.v, 32, out); }
}

// GLV split on BLS12-381: k (32-byte big endian, < r) -> k1 | k2 (16 bytes big endian each); and phi(P) = (beta * x, y)
extern "C" void hc2_glv_split(const uint8_t* k_be, uint8_t* k1_be, uint8_t* k2_be) {
  uint32_t k[8], out[8];
  hd_be_to_limbs<8>(k_be, 32, k);
  glv_bls_split(k, out);
  hd_limbs_to_be<4>(out, 16, k1_be);
  hd_limbs_to_be<4>(out + 4, 16, k2_be);
}
extern "C" void hc2_glv_phi_x(const uint8_t* x_be, uint8_t* out_be) {
  Bls::Fq x;
  hd_be_to_limbs<12>(x_be, 48, x.v);
  Bls::Fq y = (x.to_mont() * glv_bls_beta()).from_mont();
  hd_limbs_to_be<12>(y.v, 48, out_be);
}
