// Inner-product argument, host side: the mirror of `IPP` in /root/reference/src/ipp.rs with the same
// entry points (create_ipp ipp.rs:35, verify_ipp ipp.rs:204, verification_scalars ipp.rs:262).
//
// The Merlin transcript and the challenges stay here; the vectors live on the device for all lg n
// rounds (bpgpu_ipp_*).  Per round the host receives L, R (2 points) and sends u, u^-1.
#pragma once
#include <vector>

#include "device.hpp"

namespace bph {

template <class C>
struct InnerProductArgumentProof {            // ipp.rs:13-20
  std::vector<G1<C>> L, R;
  FieldElement<C> a, b;
};

template <class C>
struct IPP {
  using FE = FieldElement<C>;
  using TP = TranscriptProtocol<C>;

  // ipp.rs:35-202.  G_vec[goff..goff+n), H_vec[hoff..hoff+n) are the generator slices (prover.rs:565-574 passes
  // G[..padded_n]); every FieldElementVector has length n.  n not a power of two -> E_NOT_POW2 (assert ipp.rs:48),
  // length mismatch -> E_LEN (asserts ipp.rs:51-55).
  static int create_ipp(bpgpu_ctx* ctx, Transcript& transcript, const G1<C>& Q, const FieldElementVector<C>& G_factors,
                        const FieldElementVector<C>& H_factors, const G1Vector<C>& G_vec, size_t goff, const G1Vector<C>& H_vec,
                        size_t hoff, const FieldElementVector<C>& a_vec, const FieldElementVector<C>& b_vec, size_t n,
                        InnerProductArgumentProof<C>* proof, const G1<C>* q_base = nullptr, const FE* q_scalar = nullptr) {
    // q_base / q_scalar: optional statement that Q == q_scalar * q_base for a fixed base (the R1CS prover's Q = g * w);
    // it changes no output, it lets the device keep every round on its window tables
    bpgpu_ipp* st = nullptr;
    int rc;
    if (q_base && q_scalar) {
      uint8_t qs[C::MODBYTES];
      q_scalar->to_bytes(qs);
      rc = bpgpu_ipp_begin_fixed_q(ctx, G_vec.handle(), goff, H_vec.handle(), hoff, q_base->xy, qs, G_factors.handle(), H_factors.handle(),
                                   a_vec.handle(), b_vec.handle(), n, &st);
    } else if (bpgpu_points_has_tables(G_vec.handle()) && bpgpu_points_has_tables(H_vec.handle())) {
      // generators with window tables: Q = 1 * Q takes a table of its own (cached in the ctx; the reference's callers reuse one
      // Q for many arguments, ipp.rs:342), so that every round stays on the table path
      uint8_t one[C::MODBYTES];
      FE::one().to_bytes(one);
      rc = bpgpu_ipp_begin_fixed_q(ctx, G_vec.handle(), goff, H_vec.handle(), hoff, Q.xy, one, G_factors.handle(), H_factors.handle(),
                                   a_vec.handle(), b_vec.handle(), n, &st);
    } else {
      rc = bpgpu_ipp_begin(ctx, G_vec.handle(), goff, H_vec.handle(), hoff, Q.xy, G_factors.handle(), H_factors.handle(), a_vec.handle(),
                           b_vec.handle(), n, &st);
    }
    if (rc) return rc;
    transcript.innerproduct_domain_sep(n);                                   // ipp.rs:62
    proof->L.clear();
    proof->R.clear();
    while (bpgpu_ipp_len(st) != 1) {                                         // ipp.rs:68,138
      G1<C> L, R;
      if ((rc = bpgpu_ipp_round_LR(st, L.xy, R.xy))) break;                  // ipp.rs:77-104 / 145-170
      TP::commit_point(transcript, "L", L);                                  // ipp.rs:106-107 / 172-173
      TP::commit_point(transcript, "R", R);
      proof->L.push_back(L);
      proof->R.push_back(R);
      FE u = TP::challenge_scalar(transcript, "u");                          // ipp.rs:112 / 178
      FE u_inv = u.inverse();
      uint8_t ub[C::MODBYTES], uib[C::MODBYTES];
      u.to_bytes(ub);
      u_inv.to_bytes(uib);
      if ((rc = bpgpu_ipp_fold(st, ub, uib))) break;                         // ipp.rs:115-130 / 181-188
    }
    if (!rc) {
      uint8_t ab[C::MODBYTES], bb[C::MODBYTES];
      rc = bpgpu_ipp_finish(st, ab, bb);                                     // ipp.rs:196-201
      if (!rc) { proof->a = FE::from_bytes(ab); proof->b = FE::from_bytes(bb); }
    }
    bpgpu_ipp_free(st);
    return rc;
  }

  // The transcript replay of ipp.rs:262-298: challenges u_k (creation order), u_k^2, u_k^-2.
  // The s vector (ipp.rs:303-312) is built on the device from the challenges (bpgpu_ipp_verification_scalars).
  static int verification_challenges(const std::vector<G1<C>>& L_vec, const std::vector<G1<C>>& R_vec, size_t n, Transcript& transcript,
                                     std::vector<FE>* challenges, std::vector<FE>* u_sq, std::vector<FE>* u_inv_sq) {
    size_t lg_n = L_vec.size();
    if (lg_n >= 32) return E_VERIFICATION;                                   // ipp.rs:269-273
    if (R_vec.size() != lg_n) return E_VERIFICATION;
    if (n != ((size_t)1 << lg_n)) return E_VERIFICATION;                     // ipp.rs:274-276
    transcript.innerproduct_domain_sep(n);                                   // ipp.rs:278
    challenges->clear(); u_sq->clear(); u_inv_sq->clear();
    for (size_t k = 0; k < lg_n; k++) {                                      // ipp.rs:283-288
      TP::commit_point(transcript, "L", L_vec[k]);
      TP::commit_point(transcript, "R", R_vec[k]);
      challenges->push_back(TP::challenge_scalar(transcript, "u"));
    }
    for (const FE& u : *challenges) {                                        // ipp.rs:295-301
      u_sq->push_back(u.square());
      u_inv_sq->push_back(u.inverse().square());
    }
    return OK;
  }

  // ipp.rs:262-315 with s left on the device
  static int verification_scalars(bpgpu_ctx* ctx, const std::vector<G1<C>>& L_vec, const std::vector<G1<C>>& R_vec, size_t n,
                                  Transcript& transcript, std::vector<FE>* u_sq, std::vector<FE>* u_inv_sq, FieldElementVector<C>* s,
                                  std::vector<FE>* challenges_out = nullptr) {
    std::vector<FE> ch;
    int rc = verification_challenges(L_vec, R_vec, n, transcript, &ch, u_sq, u_inv_sq);
    if (rc) return rc;
    std::vector<uint8_t> ub(ch.size() * C::MODBYTES + 1);
    for (size_t k = 0; k < ch.size(); k++) ch[k].to_bytes(ub.data() + k * C::MODBYTES);
    bpgpu_scalars* sh = nullptr;
    if ((rc = bpgpu_ipp_verification_scalars(ctx, ub.data(), ch.size(), &sh))) return rc;
    *s = FieldElementVector<C>::adopt(ctx, sh);
    if (challenges_out) *challenges_out = ch;
    return OK;
  }

  // ipp.rs:204-260: Ok iff  a*b*Q + <a*s*Gf, G> + <b*s_rev*Hf, H> - sum u_k^2 L_k - sum u_k^-2 R_k == P
  static int verify_ipp(bpgpu_ctx* ctx, size_t n, Transcript& transcript, const FieldElementVector<C>& G_factors,
                        const FieldElementVector<C>& H_factors, const G1<C>& P, const G1<C>& Q, const G1Vector<C>& G, size_t goff,
                        const G1Vector<C>& H, size_t hoff, const FE& a, const FE& b, const std::vector<G1<C>>& L_vec,
                        const std::vector<G1<C>>& R_vec) {
    std::vector<FE> ch, u_sq, u_inv_sq;
    int rc = verification_challenges(L_vec, R_vec, n, transcript, &ch, &u_sq, &u_inv_sq);
    if (rc) return rc;
    const size_t lg = ch.size();
    std::vector<uint8_t> ub(lg * C::MODBYTES + 1), Lb(lg * 2 * C::MODBYTES + 1), Rb(lg * 2 * C::MODBYTES + 1);
    for (size_t k = 0; k < lg; k++) {
      ch[k].to_bytes(ub.data() + k * C::MODBYTES);
      memcpy(Lb.data() + k * 2 * C::MODBYTES, L_vec[k].xy, 2 * C::MODBYTES);
      memcpy(Rb.data() + k * 2 * C::MODBYTES, R_vec[k].xy, 2 * C::MODBYTES);
    }
    uint8_t ab[C::MODBYTES], bb[C::MODBYTES];
    a.to_bytes(ab);
    b.to_bytes(bb);
    G1<C> expected;
    rc = bpgpu_ipp_verify_msm(ctx, G.handle(), goff, H.handle(), hoff, Q.xy, G_factors.handle(), H_factors.handle(), ab, bb, ub.data(),
                              Lb.data(), Rb.data(), lg, expected.xy);                   // ipp.rs:220-253
    if (rc) return rc;
    return expected == P ? OK : E_VERIFICATION;                                          // ipp.rs:255-259
  }
};

}  // namespace bph
