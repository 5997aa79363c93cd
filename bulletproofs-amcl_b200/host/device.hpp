// Device-resident G1Vector / FieldElementVector (amcl_wrapper's vector newtypes as the reference uses them:
// /root/reference/src/ipp.rs:35-60, src/r1cs/prover.rs:322) as thin RAII handles over the C ABI of
// include/bpgpu.h, plus the handful of group operations the host layer needs, all executed on the device.
#pragma once
#include <initializer_list>
#include <utility>
#include <vector>

#include "curve.hpp"

namespace bph {

template <class C>
class G1Vector {
 public:
  G1Vector() = default;
  G1Vector(const G1Vector&) = delete;
  G1Vector& operator=(const G1Vector&) = delete;
  G1Vector(G1Vector&& o) noexcept : ctx_(o.ctx_), h_(o.h_), borrowed_(o.borrowed_) { o.h_ = nullptr; }
  G1Vector& operator=(G1Vector&& o) noexcept { reset(); ctx_ = o.ctx_; h_ = o.h_; borrowed_ = o.borrowed_; o.h_ = nullptr; return *this; }
  ~G1Vector() { reset(); }
  void reset() { if (h_ && !borrowed_) bpgpu_points_free(h_); h_ = nullptr; borrowed_ = false; }

  static int from_host(bpgpu_ctx* ctx, const std::vector<G1<C>>& pts, G1Vector* out) {
    return from_bytes(ctx, pts.empty() ? nullptr : pts[0].xy, pts.size(), out);   // G1<C> is exactly 2*MODBYTES bytes
  }
  static int from_bytes(bpgpu_ctx* ctx, const uint8_t* xy, size_t n, G1Vector* out) {
    out->reset();
    out->ctx_ = ctx;
    static const uint8_t dummy = 0;
    return bpgpu_points_upload(ctx, n ? xy : &dummy, n, &out->h_);
  }
  // non-owning view of a table the caller keeps alive (C ABI entry points)
  static G1Vector borrow(bpgpu_ctx* ctx, const bpgpu_points* p) { G1Vector v; v.ctx_ = ctx; v.h_ = const_cast<bpgpu_points*>(p); v.borrowed_ = true; return v; }
  size_t len() const { return h_ ? bpgpu_points_len(h_) : 0; }
  const bpgpu_points* handle() const { return h_; }
  int to_host(std::vector<G1<C>>* out) const {
    out->resize(len());
    return len() ? bpgpu_points_download(ctx_, h_, 0, len(), (*out)[0].xy) : OK;
  }

 private:
  bpgpu_ctx* ctx_ = nullptr;
  bpgpu_points* h_ = nullptr;
  bool borrowed_ = false;
};

template <class C>
class FieldElementVector {
 public:
  FieldElementVector() = default;
  FieldElementVector(const FieldElementVector&) = delete;
  FieldElementVector& operator=(const FieldElementVector&) = delete;
  FieldElementVector(FieldElementVector&& o) noexcept : ctx_(o.ctx_), h_(o.h_), borrowed_(o.borrowed_) { o.h_ = nullptr; }
  FieldElementVector& operator=(FieldElementVector&& o) noexcept { reset(); ctx_ = o.ctx_; h_ = o.h_; borrowed_ = o.borrowed_; o.h_ = nullptr; return *this; }
  ~FieldElementVector() { reset(); }
  void reset() { if (h_ && !borrowed_) bpgpu_scalars_free(h_); h_ = nullptr; borrowed_ = false; }

  static int from_host(bpgpu_ctx* ctx, const std::vector<FieldElement<C>>& v, FieldElementVector* out) {
    std::vector<uint8_t> be(v.size() * C::MODBYTES + 1);
    for (size_t i = 0; i < v.size(); i++) v[i].to_bytes(be.data() + i * C::MODBYTES);
    return from_bytes(ctx, be.data(), v.size(), out);
  }
  static int from_bytes(bpgpu_ctx* ctx, const uint8_t* be, size_t n, FieldElementVector* out) {
    out->reset();
    out->ctx_ = ctx;
    static const uint8_t dummy = 0;
    return bpgpu_scalars_upload(ctx, n ? be : &dummy, n, &out->h_);
  }
  // several host vectors in ONE upload (one copy, one conversion launch, one synchronisation), handed out as views
  static int from_host_many(bpgpu_ctx* ctx, std::initializer_list<const std::vector<FieldElement<C>>*> vs,
                            std::initializer_list<FieldElementVector*> outs) {
    size_t total = 0;
    for (auto v : vs) total += v->size();
    std::vector<uint8_t> be(total * C::MODBYTES + 1);
    size_t o = 0;
    for (auto v : vs) for (const auto& x : *v) { x.to_bytes(be.data() + o * C::MODBYTES); o++; }
    FieldElementVector all;
    int rc = from_bytes(ctx, be.data(), total, &all);
    if (rc) return rc;
    o = 0;
    auto out = outs.begin();
    for (auto v : vs) {
      if ((rc = all.view(o, v->size(), *out))) return rc;
      o += v->size();
      ++out;
    }
    return OK;
  }
  int view(size_t off, size_t n, FieldElementVector* out) {
    out->reset();
    out->ctx_ = ctx_;
    return bpgpu_scalars_view(h_, off, n, &out->h_);
  }
  static FieldElementVector adopt(bpgpu_ctx* ctx, bpgpu_scalars* h) { FieldElementVector v; v.ctx_ = ctx; v.h_ = h; return v; }
  static FieldElementVector borrow(bpgpu_ctx* ctx, const bpgpu_scalars* h) { FieldElementVector v; v.ctx_ = ctx; v.h_ = const_cast<bpgpu_scalars*>(h); v.borrowed_ = true; return v; }
  size_t len() const { return h_ ? bpgpu_scalars_len(h_) : 0; }
  const bpgpu_scalars* handle() const { return h_; }
  bpgpu_scalars* handle() { return h_; }
  int to_host(std::vector<FieldElement<C>>* out) const {
    size_t n = len();
    std::vector<uint8_t> be(n * C::MODBYTES + 1);
    int rc = n ? bpgpu_scalars_download(ctx_, h_, 0, n, be.data()) : OK;
    if (rc) return rc;
    out->resize(n);
    for (size_t i = 0; i < n; i++) (*out)[i] = FieldElement<C>::from_bytes(be.data() + i * C::MODBYTES);
    return OK;
  }

 private:
  bpgpu_ctx* ctx_ = nullptr;
  bpgpu_scalars* h_ = nullptr;
  bool borrowed_ = false;
};

// ---- group operations on host-side values, executed by the device ------------------------------------
// commitment::commit_to_field_element(g, h, v, r) = g*v + h*r = g.binary_scalar_mul(h, v, r)
// (prover.rs:123,496-500; ipp.rs:119-129)
template <class C>
inline int binary_scalar_mul(bpgpu_ctx* ctx, const G1<C>& g, const G1<C>& h, const FieldElement<C>& r1, const FieldElement<C>& r2,
                             G1<C>* out) {
  uint8_t pts[4 * C::MODBYTES], sc[2 * C::MODBYTES];
  memcpy(pts, g.xy, sizeof g.xy);
  memcpy(pts + sizeof g.xy, h.xy, sizeof h.xy);
  r1.to_bytes(sc);
  r2.to_bytes(sc + C::MODBYTES);
  return bpgpu_msm_refs(ctx, pts, sc, 2, out->xy);
}
// commitment::commit_to_field_element(g, h, v_i, r_i) for a batch of (v_i, r_i): the Pedersen pair is a fixed base
// (prover.rs:123 V, :496-500 T_1..T_6), so the ctx keeps its window table and a commitment needs no doublings.
template <class C>
inline int commit_batch(bpgpu_ctx* ctx, const G1<C>& g, const G1<C>& h, const std::vector<FieldElement<C>>& v,
                        const std::vector<FieldElement<C>>& r, std::vector<G1<C>>* out) {
  if (v.size() != r.size()) return E_LEN;
  uint8_t bases[4 * C::MODBYTES];
  memcpy(bases, g.xy, sizeof g.xy);
  memcpy(bases + sizeof g.xy, h.xy, sizeof h.xy);
  bpgpu_fixed_bases* fb = nullptr;
  int rc = bpgpu_fixed_bases_get(ctx, bases, 2, &fb);
  if (rc) return rc;
  std::vector<uint8_t> sc(2 * v.size() * C::MODBYTES + 1);
  for (size_t i = 0; i < v.size(); i++) { v[i].to_bytes(sc.data() + 2 * i * C::MODBYTES); r[i].to_bytes(sc.data() + (2 * i + 1) * C::MODBYTES); }
  out->resize(v.size());
  return bpgpu_fixed_bases_commit(ctx, fb, sc.data(), v.size(), v.empty() ? nullptr : (*out)[0].xy);
}
template <class C>
inline int commit_to_field_element(bpgpu_ctx* ctx, const G1<C>& g, const G1<C>& h, const FieldElement<C>& v, const FieldElement<C>& r,
                                   G1<C>* out) {
  std::vector<G1<C>> o;
  int rc = commit_batch<C>(ctx, g, h, {v}, {r}, &o);
  if (!rc) *out = o[0];
  return rc;
}
// &G1 * &FieldElement for a base that is part of a cached pair (prover.rs:550: Q = g * w)
template <class C>
inline int scalar_mul_fixed(bpgpu_ctx* ctx, const G1<C>& g, const G1<C>& h, const FieldElement<C>& s, G1<C>* out) {
  return commit_to_field_element<C>(ctx, g, h, s, FieldElement<C>::zero(), out);
}
// &G1 * &FieldElement, arbitrary base
template <class C>
inline int scalar_mul(bpgpu_ctx* ctx, const G1<C>& g, const FieldElement<C>& s, G1<C>* out) {
  uint8_t sc[C::MODBYTES];
  s.to_bytes(sc);
  return bpgpu_msm_refs(ctx, g.xy, sc, 1, out->xy);
}
// G1Vector::inner_product_var_time_with_ref_vecs / multi_scalar_mul_var_time on ad-hoc host lists
template <class C>
inline int msm_refs(bpgpu_ctx* ctx, const std::vector<G1<C>>& pts, const std::vector<FieldElement<C>>& sc, G1<C>* out) {
  if (pts.size() != sc.size()) return E_LEN;                       // ValueError::UnequalSizeVectors
  std::vector<uint8_t> be(sc.size() * C::MODBYTES + 1);
  for (size_t i = 0; i < sc.size(); i++) sc[i].to_bytes(be.data() + i * C::MODBYTES);
  static const uint8_t dummy = 0;
  return bpgpu_msm_refs(ctx, pts.empty() ? &dummy : pts[0].xy, be.data(), pts.size(), out->xy);
}

}  // namespace bph
