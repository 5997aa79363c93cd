// R1CS proof system, host side: the mirror of /root/reference/src/r1cs/{linear_combination,constraint_system,
// prover,verifier,proof}.rs with the same names and call order.  Constraint bookkeeping, the Merlin transcript
// and challenge derivation run here; every vector operation and every group operation runs on the device through
// include/bpgpu.h (commitment MSMs, l/r polynomial kernels, the nine inner products, IPP rounds, the single
// verification MSM).
#pragma once
#include <atomic>
#include <functional>
#include <memory>
#include <utility>
#include <vector>

#include "ipp.hpp"

namespace bph {

// ---- linear_combination.rs:11-37 ---------------------------------------------------------------------
struct Variable {
  enum Kind : uint8_t { Committed, MultiplierLeft, MultiplierRight, MultiplierOutput, One };
  Kind kind;
  size_t index;
  static Variable committed(size_t i) { return {Committed, i}; }
  static Variable left(size_t i) { return {MultiplierLeft, i}; }
  static Variable right(size_t i) { return {MultiplierRight, i}; }
  static Variable output(size_t i) { return {MultiplierOutput, i}; }
  static Variable one() { return {One, 0}; }
  bool operator==(const Variable& o) const { return kind == o.kind && (kind == One || index == o.index); }
};

// z^(q+1) * coeff for the flattening loops (prover.rs:142-184, verifier.rs:149-193): nine coefficients in ten of the range
// gadgets are +1 or -1, where the product is a copy or a negation
template <class C>
inline FieldElement<C> scaled_coeff(const FieldElement<C>& exp_z, const FieldElement<C>& coeff) {
  static const FieldElement<C> one = FieldElement<C>::one(), minus_one = FieldElement<C>::minus_one();
  if (coeff == one) return exp_z;
  if (coeff == minus_one) return exp_z.negation();
  return exp_z * coeff;
}

template <class C>
struct LinearCombination {
  using FE = FieldElement<C>;
  std::vector<std::pair<Variable, FE>> terms;

  LinearCombination() = default;
  LinearCombination(Variable v) { terms.emplace_back(v, FE::one()); }                 // From<Variable>       :74-81
  LinearCombination(const FE& s) { terms.emplace_back(Variable::one(), s); }          // From<FieldElement>   :83-89
  LinearCombination(std::vector<std::pair<Variable, FE>> t) : terms(std::move(t)) {}  // FromIterator         :91-113
  size_t len() const { return terms.size(); }

  // simplify(): merge the coefficients of equal variables, first-occurrence order (:53-67)
  LinearCombination simplify() const {
    LinearCombination out;
    for (const auto& t : terms) {
      bool found = false;
      for (auto& o : out.terms)
        if (o.first == t.first) { o.second = o.second + t.second; found = true; break; }
      if (!found) out.terms.push_back(t);
    }
    return out;
  }
  LinearCombination operator-() const {                                               // Neg :141-150
    LinearCombination r = *this;
    for (auto& t : r.terms) t.second = t.second.negation();
    return r;
  }
  LinearCombination operator+(const LinearCombination& rhs) const {                   // Add :152-159
    LinearCombination r = *this;
    r.terms.insert(r.terms.end(), rhs.terms.begin(), rhs.terms.end());
    return r;
  }
  LinearCombination operator-(const LinearCombination& rhs) const {                   // Sub :161-173
    LinearCombination r = *this;
    for (const auto& t : rhs.terms) r.terms.emplace_back(t.first, t.second.negation());
    return r;
  }
  LinearCombination& operator+=(const LinearCombination& rhs) { terms.insert(terms.end(), rhs.terms.begin(), rhs.terms.end()); return *this; }
  friend LinearCombination operator*(const FE& s, const LinearCombination& lc) {      // Mul :115-139
    LinearCombination r = lc;
    for (auto& t : r.terms) t.second = t.second * s;
    return r;
  }
};

template <class C>
struct AllocatedQuantity {                     // linear_combination.rs:26-29
  Variable variable;
  bool has_assignment;
  FieldElement<C> assignment;
};

// ---- constraint_system.rs:24-136 -----------------------------------------------------------------------
// One abstract interface serves both traits: `challenge_scalar` (RandomizedConstraintSystem) is only valid
// inside a callback given to specify_randomized_constraints.
template <class C>
class ConstraintSystem {
 public:
  using FE = FieldElement<C>;
  using LC = LinearCombination<C>;
  using Callback = std::function<int(ConstraintSystem<C>&)>;
  virtual ~ConstraintSystem() {}
  virtual void multiply(LC left, LC right, Variable* l, Variable* r, Variable* o) = 0;
  virtual int allocate(const FE* assignment, Variable* out) = 0;
  virtual int allocate_multiplier(const FE* left, const FE* right, Variable* l, Variable* r, Variable* o) = 0;
  virtual void constrain(LC lc) = 0;
  virtual int specify_randomized_constraints(Callback cb) = 0;
  virtual bool evaluate_lc(const LC& lc, FE* out) const = 0;
  virtual int allocate_single(const FE* assignment, Variable* var, bool* has_out, Variable* out) = 0;
  virtual int challenge_scalar(const char* label, FE* out) = 0;
};

// ---- proof.rs:26-58 ------------------------------------------------------------------------------------
template <class C>
struct R1CSProof {
  G1<C> A_I1, A_O1, S1, A_I2, A_O2, S2, T_1, T_3, T_4, T_5, T_6;
  FieldElement<C> t_x, t_x_blinding, e_blinding;
  InnerProductArgumentProof<C> ipp_proof;

  // Flat wire form used by this repo's fixtures and C ABI (the reference only derives serde): to_bytes() of every
  // element in struct order, then L[], R[], a, b.
  std::vector<uint8_t> to_bytes() const {
    std::vector<uint8_t> out;
    const G1<C>* pts[11] = {&A_I1, &A_O1, &S1, &A_I2, &A_O2, &S2, &T_1, &T_3, &T_4, &T_5, &T_6};
    for (auto p : pts) { auto b = p->to_bytes(); out.insert(out.end(), b.begin(), b.end()); }
    append_fe(out, t_x); append_fe(out, t_x_blinding); append_fe(out, e_blinding);
    for (const auto& p : ipp_proof.L) { auto b = p.to_bytes(); out.insert(out.end(), b.begin(), b.end()); }
    for (const auto& p : ipp_proof.R) { auto b = p.to_bytes(); out.insert(out.end(), b.begin(), b.end()); }
    append_fe(out, ipp_proof.a); append_fe(out, ipp_proof.b);
    return out;
  }
  // malformed length / tag / non-canonical scalar / coordinate >= p / point off the curve -> E_FORMAT (R1CSError::FormatError)
  static int from_bytes(const uint8_t* buf, size_t len, R1CSProof* out) {
    const size_t PB = 1 + 2 * C::MODBYTES, SB = C::MODBYTES;
    const size_t fixed = 11 * PB + 3 * SB + 2 * SB;
    if (len < fixed || (len - fixed) % (2 * PB)) return E_FORMAT;
    const size_t lg = (len - fixed) / (2 * PB);
    if (lg >= 32) return E_FORMAT;
    size_t off = 0;
    auto point = [&](G1<C>* p) {                  // 0x04 || X || Y with X, Y < p on the curve (or AMCL's identity): ECP::frombytes
      if (buf[off] != 4 || !G1<C>::xy_valid(buf + off + 1)) return false;
      *p = G1<C>::from_xy(buf + off + 1); off += PB; return true;
    };
    auto scalar = [&](FieldElement<C>* s) {
      *s = FieldElement<C>::from_bytes(buf + off);
      uint8_t chk[C::MODBYTES];
      s->to_bytes(chk);
      bool ok = memcmp(chk, buf + off, SB) == 0;   // canonical (< r)
      off += SB;
      return ok;
    };
    G1<C>* pts[11] = {&out->A_I1, &out->A_O1, &out->S1, &out->A_I2, &out->A_O2, &out->S2, &out->T_1, &out->T_3, &out->T_4, &out->T_5, &out->T_6};
    for (auto p : pts) if (!point(p)) return E_FORMAT;
    if (!scalar(&out->t_x) || !scalar(&out->t_x_blinding) || !scalar(&out->e_blinding)) return E_FORMAT;
    out->ipp_proof.L.resize(lg);
    out->ipp_proof.R.resize(lg);
    for (size_t k = 0; k < lg; k++) if (!point(&out->ipp_proof.L[k])) return E_FORMAT;
    for (size_t k = 0; k < lg; k++) if (!point(&out->ipp_proof.R[k])) return E_FORMAT;
    if (!scalar(&out->ipp_proof.a) || !scalar(&out->ipp_proof.b)) return E_FORMAT;
    return OK;
  }
};

// process-wide switch (bph_set_secret_fixed_schedule, or BPH_FIXED_SCHEDULE=1 in the environment): the prover's witness
// commitments A_I, A_O, S run with bpgpu_ctx_set_fixed_schedule on
inline std::atomic<int>& secret_fixed_schedule_flag() {
  static std::atomic<int> f{getenv("BPH_FIXED_SCHEDULE") && atoi(getenv("BPH_FIXED_SCHEDULE")) ? 1 : 0};
  return f;
}
inline bool secret_fixed_schedule() { return secret_fixed_schedule_flag().load(std::memory_order_relaxed) != 0; }

inline size_t next_power_of_two(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }   // usize::next_power_of_two (0 -> 1)

// ---- prover.rs ---------------------------------------------------------------------------------------
template <class C> struct BatchProverAccess;    // prove_batch.hpp: the lock-step batched prover drives these stages itself

template <class C>
class Prover : public ConstraintSystem<C> {
  friend struct BatchProverAccess<C>;
 public:
  using FE = FieldElement<C>;
  using LC = LinearCombination<C>;
  using TP = TranscriptProtocol<C>;
  using Callback = typename ConstraintSystem<C>::Callback;

  // prover.rs:84-101.  `rng` supplies every FieldElement::random() of prove() in the reference's order.
  Prover(bpgpu_ctx* ctx, const G1<C>& g, const G1<C>& h, Transcript& transcript, Rng<C>& rng)
      : ctx_(ctx), g_(g), h_(h), transcript_(transcript), rng_(rng) {
    transcript_.r1cs_domain_sep();
  }

  // prover.rs:119-129
  int commit(const FE& v, const FE& v_blinding, G1<C>* V, Variable* var) {
    size_t i = v_.size();
    int rc = commit_to_field_element(ctx_, g_, h_, v, v_blinding, V);
    if (rc) return rc;
    v_.push_back(v);
    v_blinding_.push_back(v_blinding);
    TP::commit_point(transcript_, "V", *V);
    *var = Variable::committed(i);
    return OK;
  }

  // m calls of commit() with ONE device launch: V_j = commit_to_field_element(g, h, v_j, blinding_j), appended to the
  // transcript in order -- byte-identical to calling commit() m times (prover.rs:119-129).
  int commit_vec(const std::vector<FE>& v, const std::vector<FE>& v_blinding, std::vector<G1<C>>* V, std::vector<Variable>* vars) {
    int rc = commit_batch<C>(ctx_, g_, h_, v, v_blinding, V);
    if (rc) return rc;
    vars->clear();
    for (size_t j = 0; j < v.size(); j++) {
      vars->push_back(Variable::committed(v_.size()));
      v_.push_back(v[j]);
      v_blinding_.push_back(v_blinding[j]);
      TP::commit_point(transcript_, "V", (*V)[j]);
    }
    return OK;
  }

  // commit (prover.rs:119-129) with the commitment point already evaluated (a batch evaluates them all in one call)
  Variable commit_precomputed(const FE& v, const FE& v_blinding, const G1<C>& V) {
    Variable var = Variable::committed(v_.size());
    v_.push_back(v);
    v_blinding_.push_back(v_blinding);
    TP::commit_point(transcript_, "V", V);
    return var;
  }

  size_t num_constraints() const { return constraints_.size(); }     // prover.rs:595-597
  size_t num_multipliers() const { return a_O_.size(); }             // prover.rs:599-601

  // ---- ConstraintSystem (prover.rs:604-691)
  void multiply(LC left, LC right, Variable* l, Variable* r, Variable* o) override {
    FE lv = eval(left), rv = eval(right);
    allocate_vars(lv, rv, lv * rv, l, r, o);
    left.terms.emplace_back(*l, FE::minus_one());
    right.terms.emplace_back(*r, FE::minus_one());
    constrain(std::move(left));
    constrain(std::move(right));
  }
  int allocate(const FE* assignment, Variable* out) override {
    if (!assignment) return E_MISSING_ASSIGNMENT;
    if (!has_pending_) {
      size_t i = a_L_.size();
      has_pending_ = true; pending_ = i;
      a_L_.push_back(*assignment); a_R_.push_back(FE::zero()); a_O_.push_back(FE::zero());
      *out = Variable::left(i);
    } else {
      size_t i = pending_;
      has_pending_ = false;
      a_R_[i] = *assignment;
      a_O_[i] = a_L_[i] * a_R_[i];
      *out = Variable::right(i);
    }
    return OK;
  }
  int allocate_multiplier(const FE* left, const FE* right, Variable* l, Variable* r, Variable* o) override {
    if (!left || !right) return E_MISSING_ASSIGNMENT;
    allocate_vars(*left, *right, (*left) * (*right), l, r, o);
    return OK;
  }
  void constrain(LC lc) override { constraints_.push_back(std::move(lc)); }
  int specify_randomized_constraints(Callback cb) override {
    if (randomizing_) return cb(*this);                               // RandomizingProver: prover.rs:740-745
    deferred_.push_back(std::move(cb));
    return OK;
  }
  bool evaluate_lc(const LC& lc, FE* out) const override { *out = eval(lc); return true; }
  int allocate_single(const FE* assignment, Variable* var, bool* has_out, Variable* out) override {
    int rc = allocate(assignment, var);
    if (rc) return rc;
    *has_out = var->kind == Variable::MultiplierRight;
    if (*has_out) *out = Variable::output(var->index);
    return OK;
  }
  int challenge_scalar(const char* label, FE* out) override {         // RandomizedConstraintSystem, prover.rs:759-763
    if (!randomizing_) return E_ARG;
    *out = TP::challenge_scalar(transcript_, label);
    return OK;
  }

  // A one-phase statement whose circuit is already on the device (a bpgpu_circuit recorded once per statement shape) and
  // whose multiplier assignments [a_L | a_R | a_O] were built there: prove() then never touches LinearCombinations --
  // no gadget run, no host flattening, no witness upload.  Same transcript, same draws, same proof bytes.
  struct DeviceCircuit {
    const bpgpu_circuit* circuit;
    const bpgpu_scalars* witness;    // 3 * n scalars
    size_t n;
  };

  // prover.rs:322-593.  G, H are device-resident generator tables of length >= padded_n.
  int prove(const G1Vector<C>& G, const G1Vector<C>& H, R1CSProof<C>* proof, const DeviceCircuit* dc = nullptr) {
    Trace tr("prove");
    if (dc && (!a_L_.empty() || !constraints_.empty() || !deferred_.empty() || bpgpu_circuit_multipliers(dc->circuit) != dc->n ||
               bpgpu_circuit_commitments(dc->circuit) != v_.size()))
      return E_ARG;
    transcript_.append_u64("m", v_.size());                            // :327
    const size_t n1 = dc ? dc->n : a_L_.size();
    if (G.len() < n1 || H.len() < n1) return E_INVALID_GENERATORS_LENGTH;   // :332-334
    const FE i_blinding1 = rng_.next(), o_blinding1 = rng_.next(), s_blinding1 = rng_.next();   // :336-338
    int rc;
    // The first-phase witness goes to the device once and stays: the same vectors feed the commitment MSMs here and the
    // polynomial kernels later.  s_L, s_R (:340-341) are DRAWN on the device (FieldElementVector::random as
    // bpgpu_fr_random: the same stream positions n1 + n1 host draws would use), so they never exist on the host.
    FieldElementVector<C> d_aL, d_aR, d_aO, d_sL, d_sR;
    if (dc) {
      FieldElementVector<C> w = FieldElementVector<C>::borrow(ctx_, dc->witness);
      if ((rc = w.view(0, n1, &d_aL)) || (rc = w.view(n1, n1, &d_aR)) || (rc = w.view(2 * n1, n1, &d_aO))) return rc;
    } else if ((rc = FieldElementVector<C>::from_host_many(ctx_, {&a_L_, &a_R_, &a_O_}, {&d_aL, &d_aR, &d_aO}))) {
      return rc;
    }
    {
      bpgpu_scalars* hs = nullptr;                                     // s_L (:340) then s_R (:341): 2*n1 consecutive draws
      if ((rc = rng_.fill_device(ctx_, 2 * n1, &hs))) return rc;
      FieldElementVector<C> both = FieldElementVector<C>::adopt(ctx_, hs);
      if ((rc = both.view(0, n1, &d_sL)) || (rc = both.view(n1, n1, &d_sR))) return rc;
    }
    tr.mark("witness upload + blindings");
    if ((rc = phase_commit_device(G, H, n1, d_aL, d_aR, d_aO, d_sL, d_sR, i_blinding1, o_blinding1, s_blinding1, &proof->A_I1, &proof->A_O1,
                                  &proof->S1)))
      return rc;
    tr.mark("A_I1 A_O1 S1");
    TP::commit_point(transcript_, "A_I1", proof->A_I1);                // :364-366
    TP::commit_point(transcript_, "A_O1", proof->A_O1);
    TP::commit_point(transcript_, "S1", proof->S1);

    if ((rc = create_randomized_constraints())) return rc;             // :369

    const size_t n = dc ? n1 : a_L_.size();
    const size_t n2 = n - n1;
    const size_t padded_n = next_power_of_two(n);
    if (G.len() < padded_n || H.len() < padded_n) return E_INVALID_GENERATORS_LENGTH;   // :379-381
    const bool has_2nd = n2 > 0;
    FE i_blinding2 = FE::zero(), o_blinding2 = FE::zero(), s_blinding2 = FE::zero();
    if (has_2nd) { i_blinding2 = rng_.next(); o_blinding2 = rng_.next(); s_blinding2 = rng_.next(); }   // :387-399
    if (has_2nd) {                                                     // :401-427
      // two-phase circuits (randomised constraints): the second-phase variables are few; bring s_L, s_R to the host,
      // extend them (:401-402), commit the second phase from host vectors and re-upload the full-length witness
      std::vector<FE> s_L, s_R;
      if ((rc = d_sL.to_host(&s_L)) || (rc = d_sR.to_host(&s_R))) return rc;
      s_L.resize(n); s_R.resize(n);
      for (size_t i = n1; i < n; i++) s_L[i] = rng_.next();            // :401
      for (size_t i = n1; i < n; i++) s_R[i] = rng_.next();            // :402
      if ((rc = phase_commit(G, H, n1, n, s_L, s_R, i_blinding2, o_blinding2, s_blinding2, &proof->A_I2, &proof->A_O2, &proof->S2))) return rc;
      if ((rc = FieldElementVector<C>::from_host(ctx_, a_L_, &d_aL)) || (rc = FieldElementVector<C>::from_host(ctx_, a_R_, &d_aR)) ||
          (rc = FieldElementVector<C>::from_host(ctx_, a_O_, &d_aO)) || (rc = FieldElementVector<C>::from_host(ctx_, s_L, &d_sL)) ||
          (rc = FieldElementVector<C>::from_host(ctx_, s_R, &d_sR)))
        return rc;
    } else {
      proof->A_I2 = proof->A_O2 = proof->S2 = G1<C>::identity();      // :429
    }
    TP::commit_point(transcript_, "A_I2", proof->A_I2);                // :432-434
    TP::commit_point(transcript_, "A_O2", proof->A_O2);
    TP::commit_point(transcript_, "S2", proof->S2);

    const FE y = TP::challenge_scalar(transcript_, "y");               // :438-439
    const FE z = TP::challenge_scalar(transcript_, "z");
    std::vector<FE> wL, wR, wO, wV;
    FieldElementVector<C> d_wL, d_wR, d_wO;                            // l(x), r(x) coefficient vectors on the device (:458-486)
    if (dc) {
      // the recorded matrix times the powers of z, on the device; only wV (m scalars, for t_2_blinding) comes back
      uint8_t zb[C::MODBYTES];
      z.to_bytes(zb);
      bpgpu_scalars* hw = nullptr;
      if ((rc = bpgpu_circuit_flatten(ctx_, dc->circuit, zb, &hw))) return rc;
      FieldElementVector<C> all = FieldElementVector<C>::adopt(ctx_, hw), d_wV;
      if ((rc = all.view(0, n, &d_wL)) || (rc = all.view(n, n, &d_wR)) || (rc = all.view(2 * n, n, &d_wO)) ||
          (rc = all.view(3 * n, v_.size(), &d_wV)) || (rc = d_wV.to_host(&wV)))
        return rc;
    } else {
      flattened_constraints(z, &wL, &wR, &wO, &wV);                    // :441
      if ((rc = FieldElementVector<C>::from_host_many(ctx_, {&wL, &wR, &wO}, {&d_wL, &d_wR, &d_wO}))) return rc;
    }
    tr.mark("flattened_constraints");
    uint8_t yb[C::MODBYTES];
    y.to_bytes(yb);
    bpgpu_scalars *h_l1 = nullptr, *h_r0 = nullptr, *h_r1 = nullptr, *h_r3 = nullptr;
    if ((rc = bpgpu_r1cs_prover_polys(ctx_, n, d_aL.handle(), d_aR.handle(), d_sR.handle(), d_wL.handle(), d_wR.handle(), d_wO.handle(), yb,
                                      &h_l1, &h_r0, &h_r1, &h_r3)))
      return rc;
    FieldElementVector<C> l1 = FieldElementVector<C>::adopt(ctx_, h_l1), r0 = FieldElementVector<C>::adopt(ctx_, h_r0),
                          r1 = FieldElementVector<C>::adopt(ctx_, h_r1), r3 = FieldElementVector<C>::adopt(ctx_, h_r3);
    // t_poly = <l(x), r(x)> (:488, vector_poly.rs:79-97)
    uint8_t tb[6 * C::MODBYTES];
    if ((rc = bpgpu_fr_poly3_special_inner_product(ctx_, l1.handle(), d_aO.handle(), d_sL.handle(), r0.handle(), r1.handle(), r3.handle(), n, tb)))
      return rc;
    tr.mark("polys + t_poly");
    FE t[7];
    for (int k = 1; k <= 6; k++) t[k] = FE::from_bytes(tb + (k - 1) * C::MODBYTES);
    FE tb_[7];
    tb_[1] = rng_.next(); tb_[3] = rng_.next(); tb_[4] = rng_.next(); tb_[5] = rng_.next(); tb_[6] = rng_.next();   // :490-494
    {                                                                  // T_i = commit_to_field_element(g, h, t_i, blinding_i)  :496-500
      std::vector<G1<C>> T;
      if ((rc = commit_batch<C>(ctx_, g_, h_, {t[1], t[3], t[4], t[5], t[6]}, {tb_[1], tb_[3], tb_[4], tb_[5], tb_[6]}, &T))) return rc;
      proof->T_1 = T[0]; proof->T_3 = T[1]; proof->T_4 = T[2]; proof->T_5 = T[3]; proof->T_6 = T[4];
    }
    TP::commit_point(transcript_, "T_1", proof->T_1);                  // :502-506
    TP::commit_point(transcript_, "T_3", proof->T_3);
    TP::commit_point(transcript_, "T_4", proof->T_4);
    TP::commit_point(transcript_, "T_5", proof->T_5);
    TP::commit_point(transcript_, "T_6", proof->T_6);
    const FE u = TP::challenge_scalar(transcript_, "u");               // :508-509
    const FE x = TP::challenge_scalar(transcript_, "x");
    tb_[2] = FE::zero();                                               // t_2_blinding = <wV, v_blinding> (:513)
    for (size_t j = 0; j < wV.size(); j++) tb_[2] = tb_[2] + wV[j] * v_blinding_[j];
    auto poly6 = [&](const FE* c) { return x * (c[1] + x * (c[2] + x * (c[3] + x * (c[4] + x * (c[5] + x * c[6]))))); };   // vector_poly.rs:115-119
    proof->t_x = poly6(t);                                             // :524
    proof->t_x_blinding = poly6(tb_);                                  // :525
    const FE i_blinding = i_blinding1 + u * i_blinding2;               // :537-539
    const FE o_blinding = o_blinding1 + u * o_blinding2;
    const FE s_blinding = s_blinding1 + u * s_blinding2;
    proof->e_blinding = x * (i_blinding + x * (o_blinding + x * s_blinding));   // :541
    TP::commit_scalar(transcript_, "t_x", proof->t_x);                 // :543-546
    TP::commit_scalar(transcript_, "t_x_blinding", proof->t_x_blinding);
    TP::commit_scalar(transcript_, "e_blinding", proof->e_blinding);
    const FE w = TP::challenge_scalar(transcript_, "w");               // :549
    // Q = g * w (:550) is handed to the IPP as the pair (g, w): with window tables the device never needs the point
    // itself (c_L * Q = (c_L * w) * g), without them bpgpu_ipp_begin_fixed_q evaluates it once
    const G1<C> Q = G1<C>::identity();
    // l_vec, r_vec (with padding), G_factors, H_factors on the device (:526-535, :552-563)
    uint8_t xb[C::MODBYTES], ub[C::MODBYTES];
    x.to_bytes(xb);
    u.to_bytes(ub);
    bpgpu_scalars *h_lv = nullptr, *h_rv = nullptr, *h_gf = nullptr, *h_hf = nullptr;
    if ((rc = bpgpu_r1cs_prover_eval(ctx_, n, n1, padded_n, l1.handle(), d_aO.handle(), d_sL.handle(), r0.handle(), r1.handle(), r3.handle(), xb,
                                     ub, yb, &h_lv, &h_rv, &h_gf, &h_hf)))
      return rc;
    tr.mark("T, Q, eval");
    FieldElementVector<C> l_vec = FieldElementVector<C>::adopt(ctx_, h_lv), r_vec = FieldElementVector<C>::adopt(ctx_, h_rv),
                          G_factors = FieldElementVector<C>::adopt(ctx_, h_gf), H_factors = FieldElementVector<C>::adopt(ctx_, h_hf);
    rc = IPP<C>::create_ipp(ctx_, transcript_, Q, G_factors, H_factors, G, 0, H, 0, l_vec, r_vec, padded_n, &proof->ipp_proof, &g_, &w);   // :565-574
    tr.mark("create_ipp");
    return rc;
  }

 private:
  bpgpu_ctx* ctx_;
  G1<C> g_, h_;
  Transcript& transcript_;
  Rng<C>& rng_;
  std::vector<LC> constraints_;
  std::vector<FE> a_L_, a_R_, a_O_, v_, v_blinding_;
  std::vector<Callback> deferred_;
  bool has_pending_ = false, randomizing_ = false;
  size_t pending_ = 0;

  void allocate_vars(const FE& l, const FE& r, const FE& o, Variable* lv, Variable* rv, Variable* ov) {   // prover.rs:693-711
    *lv = Variable::left(a_L_.size());
    *rv = Variable::right(a_R_.size());
    *ov = Variable::output(a_O_.size());
    a_L_.push_back(l); a_R_.push_back(r); a_O_.push_back(o);
  }
  FE eval(const LC& lc) const {                                        // prover.rs:282-296
    FE sum = FE::zero();
    for (const auto& t : lc.terms) {
      FE val;
      switch (t.first.kind) {
        case Variable::MultiplierLeft: val = a_L_[t.first.index]; break;
        case Variable::MultiplierRight: val = a_R_[t.first.index]; break;
        case Variable::MultiplierOutput: val = a_O_[t.first.index]; break;
        case Variable::Committed: val = v_[t.first.index]; break;
        default: val = FE::one();
      }
      sum = sum + t.second * val;
    }
    return sum;
  }
  // prover.rs:142-184: wL, wR, wO, wV = z * z^Q * W_{L,R,O,V}; a sparse scatter over the constraints, host side
  void flattened_constraints(const FE& z, std::vector<FE>* wL, std::vector<FE>* wR, std::vector<FE>* wO, std::vector<FE>* wV) const {
    const size_t n = a_L_.size(), m = v_.size();
    wL->assign(n, FE::zero()); wR->assign(n, FE::zero()); wO->assign(n, FE::zero()); wV->assign(m, FE::zero());
    FE exp_z = z;
    for (const LC& lc : constraints_) {
      for (const auto& t : lc.terms) {
        switch (t.first.kind) {
          case Variable::MultiplierLeft: (*wL)[t.first.index] = (*wL)[t.first.index] + scaled_coeff<C>(exp_z, t.second); break;
          case Variable::MultiplierRight: (*wR)[t.first.index] = (*wR)[t.first.index] + scaled_coeff<C>(exp_z, t.second); break;
          case Variable::MultiplierOutput: (*wO)[t.first.index] = (*wO)[t.first.index] + scaled_coeff<C>(exp_z, t.second); break;
          case Variable::Committed: (*wV)[t.first.index] = (*wV)[t.first.index] - scaled_coeff<C>(exp_z, t.second); break;
          default: break;                                               // the prover ignores constant terms
        }
      }
      exp_z = exp_z * z;
    }
  }
  // prover.rs:300-319
  int create_randomized_constraints() {
    has_pending_ = false;
    if (deferred_.empty()) { transcript_.r1cs_1phase_domain_sep(); return OK; }
    transcript_.r1cs_2phase_domain_sep();
    std::vector<Callback> cbs;
    cbs.swap(deferred_);
    randomizing_ = true;
    int rc = OK;
    for (auto& cb : cbs) if ((rc = cb(*this))) break;
    randomizing_ = false;
    return rc;
  }
  // A_I = <a_L, G> + <a_R, H> + i_b*h ; A_O = <a_O, G> + o_b*h ; S = <s_L, G> + <s_R, H> + s_b*h over the
  // multipliers [lo, hi)  (prover.rs:347-362 and :404-427): three composite MSMs on cached generator tables
  // the same three commitments from device-resident scalar vectors (first phase: generators and scalars at offset 0)
  int phase_commit_device(const G1Vector<C>& G, const G1Vector<C>& H, size_t k, const FieldElementVector<C>& aL, const FieldElementVector<C>& aR,
                          const FieldElementVector<C>& aO, const FieldElementVector<C>& sL, const FieldElementVector<C>& sR, const FE& i_b,
                          const FE& o_b, const FE& s_b, G1<C>* A_I, G1<C>* A_O, G1<C>* S) {
    uint8_t ib[C::MODBYTES], ob[C::MODBYTES], sb[C::MODBYTES];
    i_b.to_bytes(ib); o_b.to_bytes(ob); s_b.to_bytes(sb);
    auto part_dev = [&](const G1Vector<C>& T, const FieldElementVector<C>& sc) { bpgpu_msm_part p{T.handle(), 0, nullptr, sc.handle(), 0, nullptr, k}; return p; };
    auto part_h = [&](const uint8_t* sc) { bpgpu_msm_part p{nullptr, 0, h_.xy, nullptr, 0, sc, 1}; return p; };
    bpgpu_msm_part parts[8] = {part_dev(G, aL), part_dev(H, aR), part_h(ib),      // A_I  (prover.rs:347-355)
                               part_dev(G, aO), part_h(ob),                       // A_O  (:358)
                               part_dev(G, sL), part_dev(H, sR), part_h(sb)};    // S    (:361-362)
    const size_t counts[3] = {3, 2, 3};
    uint8_t out[3 * 2 * C::MODBYTES];
    // the witness and its blindings are the prover's secrets: fixed-schedule table sums when the caller asked for them
    // (secret_fixed_schedule(); the reference uses inner_product_const_time here, prover.rs:347-362)
    if (secret_fixed_schedule()) bpgpu_ctx_set_fixed_schedule(ctx_, 1);
    int rc = bpgpu_msm_parts_batch(ctx_, parts, counts, 3, out);
    if (secret_fixed_schedule()) bpgpu_ctx_set_fixed_schedule(ctx_, 0);
    if (rc) return rc;
    *A_I = G1<C>::from_xy(out);
    *A_O = G1<C>::from_xy(out + 2 * C::MODBYTES);
    *S = G1<C>::from_xy(out + 4 * C::MODBYTES);
    return OK;
  }
  int phase_commit(const G1Vector<C>& G, const G1Vector<C>& H, size_t lo, size_t hi, const std::vector<FE>& s_L, const std::vector<FE>& s_R,
                   const FE& i_b, const FE& o_b, const FE& s_b, G1<C>* A_I, G1<C>* A_O, G1<C>* S) {
    const size_t k = hi - lo, mb = C::MODBYTES;
    auto pack = [&](const std::vector<FE>& v) {
      std::vector<uint8_t> be(k * mb + 1);
      for (size_t i = 0; i < k; i++) v[lo + i].to_bytes(be.data() + i * mb);
      return be;
    };
    Trace tr("phase_commit");
    std::vector<uint8_t> aL = pack(a_L_), aR = pack(a_R_), aO = pack(a_O_), sL = pack(s_L), sR = pack(s_R);
    tr.mark("pack to bytes");
    uint8_t ib[C::MODBYTES], ob[C::MODBYTES], sb[C::MODBYTES];
    i_b.to_bytes(ib); o_b.to_bytes(ob); s_b.to_bytes(sb);
    auto part_dev = [&](const G1Vector<C>& T, const uint8_t* sc) { bpgpu_msm_part p{T.handle(), lo, nullptr, nullptr, 0, sc, k}; return p; };
    auto part_h = [&](const uint8_t* sc) { bpgpu_msm_part p{nullptr, 0, h_.xy, nullptr, 0, sc, 1}; return p; };
    bpgpu_msm_part parts[8] = {part_dev(G, aL.data()), part_dev(H, aR.data()), part_h(ib),      // A_I
                               part_dev(G, aO.data()), part_h(ob),                                // A_O
                               part_dev(G, sL.data()), part_dev(H, sR.data()), part_h(sb)};      // S
    const size_t counts[3] = {3, 2, 3};
    uint8_t out[3 * 2 * C::MODBYTES];
    int rc = bpgpu_msm_parts_batch(ctx_, parts, counts, 3, out);
    if (rc) return rc;
    *A_I = G1<C>::from_xy(out);
    *A_O = G1<C>::from_xy(out + 2 * C::MODBYTES);
    *S = G1<C>::from_xy(out + 4 * C::MODBYTES);
    return OK;
  }
};

// ---- verifier.rs -------------------------------------------------------------------------------------
template <class C>
class Verifier : public ConstraintSystem<C> {
 public:
  using FE = FieldElement<C>;
  using LC = LinearCombination<C>;
  using TP = TranscriptProtocol<C>;
  using Callback = typename ConstraintSystem<C>::Callback;

  Verifier(bpgpu_ctx* ctx, Transcript& transcript) : ctx_(ctx), transcript_(transcript) { transcript_.r1cs_domain_sep(); }   // verifier.rs:97-108

  Variable commit(const G1<C>& commitment) {                           // verifier.rs:124-132
    size_t i = V_.size();
    V_.push_back(commitment);
    TP::commit_point(transcript_, "V", commitment);
    return Variable::committed(i);
  }

  // ---- ConstraintSystem (verifier.rs:460-534)
  void multiply(LC left, LC right, Variable* l, Variable* r, Variable* o) override {
    allocate_vars(l, r, o);
    left.terms.emplace_back(*l, FE::minus_one());
    right.terms.emplace_back(*r, FE::minus_one());
    constrain(std::move(left));
    constrain(std::move(right));
  }
  int allocate(const FE*, Variable* out) override {
    if (!has_pending_) { size_t i = num_vars_++; has_pending_ = true; pending_ = i; *out = Variable::left(i); }
    else { has_pending_ = false; *out = Variable::right(pending_); }
    return OK;
  }
  int allocate_multiplier(const FE*, const FE*, Variable* l, Variable* r, Variable* o) override { allocate_vars(l, r, o); return OK; }
  void constrain(LC lc) override { constraints_.push_back(std::move(lc)); }
  int specify_randomized_constraints(Callback cb) override {
    if (randomizing_) return cb(*this);
    deferred_.push_back(std::move(cb));
    return OK;
  }
  bool evaluate_lc(const LC&, FE*) const override { return false; }
  int allocate_single(const FE*, Variable* var, bool* has_out, Variable* out) override {
    int rc = allocate(nullptr, var);
    if (rc) return rc;
    *has_out = var->kind == Variable::MultiplierRight;
    if (*has_out) *out = Variable::output(var->index);
    return OK;
  }
  int challenge_scalar(const char* label, FE* out) override {
    if (!randomizing_) return E_ARG;
    *out = TP::challenge_scalar(transcript_, label);
    return OK;
  }

  // verifier.rs:267-457.  `r` is the verifier's random batching scalar (FieldElement::random(), :392).
  // Returns OK, E_VERIFICATION or E_INVALID_GENERATORS_LENGTH.
  // `recorded`: the statement's circuit already on the device (a one-phase bpgpu_circuit with as many commitments as this
  // verifier holds); no gadget has been run against this verifier, the weights come from bpgpu_circuit_flatten.
  int verify(const R1CSProof<C>& proof, const G1<C>& g, const G1<C>& h, const G1Vector<C>& G, const G1Vector<C>& H, const FE& r,
             const bpgpu_circuit* recorded = nullptr) {
    bool is_identity = false;
    if (recorded) {
      if (num_vars_ || !constraints_.empty() || !deferred_.empty() || bpgpu_circuit_commitments(recorded) != V_.size()) return E_ARG;
      num_vars_ = bpgpu_circuit_multipliers(recorded);
    }
    int rc = verification_msm(proof, g, h, G, H, r, &is_identity, recorded);
    if (rc) return rc;
    return is_identity ? OK : E_VERIFICATION;                          // :453-456
  }

 private:
  bpgpu_ctx* ctx_;
  Transcript& transcript_;
  std::vector<LC> constraints_;
  size_t num_vars_ = 0;
  std::vector<G1<C>> V_;
  std::vector<Callback> deferred_;
  bool has_pending_ = false, randomizing_ = false;
  size_t pending_ = 0;

  void allocate_vars(Variable* l, Variable* r, Variable* o) {          // verifier.rs:536-548
    size_t i = num_vars_++;
    *l = Variable::left(i); *r = Variable::right(i); *o = Variable::output(i);
  }
  // verifier.rs:149-193
  void flattened_constraints(const FE& z, std::vector<FE>* wL, std::vector<FE>* wR, std::vector<FE>* wO, std::vector<FE>* wV, FE* wc) const {
    const size_t n = num_vars_, m = V_.size();
    wL->assign(n, FE::zero()); wR->assign(n, FE::zero()); wO->assign(n, FE::zero()); wV->assign(m, FE::zero());
    *wc = FE::zero();
    FE exp_z = z;
    for (const LC& lc : constraints_) {
      for (const auto& t : lc.terms) {
        switch (t.first.kind) {
          case Variable::MultiplierLeft: (*wL)[t.first.index] = (*wL)[t.first.index] + scaled_coeff<C>(exp_z, t.second); break;
          case Variable::MultiplierRight: (*wR)[t.first.index] = (*wR)[t.first.index] + scaled_coeff<C>(exp_z, t.second); break;
          case Variable::MultiplierOutput: (*wO)[t.first.index] = (*wO)[t.first.index] + scaled_coeff<C>(exp_z, t.second); break;
          case Variable::Committed: (*wV)[t.first.index] = (*wV)[t.first.index] - scaled_coeff<C>(exp_z, t.second); break;
          default: *wc = *wc - scaled_coeff<C>(exp_z, t.second); break;
        }
      }
      exp_z = exp_z * z;
    }
  }
  int create_randomized_constraints() {                                // verifier.rs:245-264
    has_pending_ = false;
    if (deferred_.empty()) { transcript_.r1cs_1phase_domain_sep(); return OK; }
    transcript_.r1cs_2phase_domain_sep();
    std::vector<Callback> cbs;
    cbs.swap(deferred_);
    randomizing_ = true;
    int rc = OK;
    for (auto& cb : cbs) if ((rc = cb(*this))) break;
    randomizing_ = false;
    return rc;
  }

  struct Challenges { FE y, z, u, x, w; size_t n1, n, padded_n; };

 public:
  // flattened_constraints (verifier.rs:149-193) as a sparse matrix, for circuits shared by a whole batch of proofs
  // (bpgpu_circuit_create): rows = variables [wL 0..n) | wR | wO | wV 0..m) | wc], an entry = (constraint index q, coefficient)
  // meaning coefficient * z^(q+1), with the signs of :166-190 applied (committed and constant terms negated).
  // ent_q flags coefficients +1 (bit 31) and -1 (bit 30).  One-phase circuits only (no deferred constraints): E_ARG otherwise.
  struct CircuitCSR { size_t n = 0, m = 0, q = 0; std::vector<uint32_t> row_start, ent_q; std::vector<uint8_t> ent_c_be; };
  int export_csr(CircuitCSR* out) const {
    if (!deferred_.empty() || constraints_.size() >= (1u << 30)) return E_ARG;
    const size_t n = num_vars_, m = V_.size(), rows = 3 * n + m + 1;
    auto row_of = [&](const Variable& v) -> size_t {
      switch (v.kind) {
        case Variable::MultiplierLeft: return v.index;
        case Variable::MultiplierRight: return n + v.index;
        case Variable::MultiplierOutput: return 2 * n + v.index;
        case Variable::Committed: return 3 * n + v.index;
        default: return 3 * n + m;
      }
    };
    out->n = n; out->m = m; out->q = constraints_.size();
    out->row_start.assign(rows + 1, 0);
    for (const LC& lc : constraints_)
      for (const auto& t : lc.terms) out->row_start[row_of(t.first) + 1]++;
    for (size_t r = 0; r < rows; r++) out->row_start[r + 1] += out->row_start[r];
    const size_t nnz = out->row_start[rows];
    out->ent_q.assign(nnz, 0);
    out->ent_c_be.assign(nnz * C::MODBYTES + 1, 0);
    std::vector<uint32_t> fill(out->row_start.begin(), out->row_start.end() - 1);
    const FE one = FE::one(), minus_one = FE::minus_one();
    uint32_t q = 0;
    for (const LC& lc : constraints_) {
      for (const auto& t : lc.terms) {
        const bool neg = t.first.kind == Variable::Committed || t.first.kind == Variable::One;
        const FE c = neg ? t.second.negation() : t.second;
        const uint32_t e = fill[row_of(t.first)]++;
        out->ent_q[e] = q | (c == one ? 0x80000000u : 0u) | (c == minus_one ? 0x40000000u : 0u);
        c.to_bytes(out->ent_c_be.data() + (size_t)e * C::MODBYTES);
      }
      q++;
    }
    return OK;
  }

 private:
  // verifier.rs:279-323: absorb the proof, derive y, z, u, x, w (and run the deferred constraints in between)
  int replay_transcript(const R1CSProof<C>& proof, size_t gens_len, Challenges* c) {
    transcript_.append_u64("m", V_.size());                            // :279
    c->n1 = num_vars_;
    TP::commit_point(transcript_, "A_I1", proof.A_I1);                 // :282-284
    TP::commit_point(transcript_, "A_O1", proof.A_O1);
    TP::commit_point(transcript_, "S1", proof.S1);
    int rc = create_randomized_constraints();                          // :287
    if (rc) return rc;
    c->n = num_vars_;
    c->padded_n = next_power_of_two(c->n);
    if (gens_len < c->padded_n) return E_INVALID_GENERATORS_LENGTH;    // :297-299
    TP::commit_point(transcript_, "A_I2", proof.A_I2);                 // :301-303
    TP::commit_point(transcript_, "A_O2", proof.A_O2);
    TP::commit_point(transcript_, "S2", proof.S2);
    c->y = TP::challenge_scalar(transcript_, "y");                     // :305-306
    c->z = TP::challenge_scalar(transcript_, "z");
    TP::commit_point(transcript_, "T_1", proof.T_1);                   // :308-312
    TP::commit_point(transcript_, "T_3", proof.T_3);
    TP::commit_point(transcript_, "T_4", proof.T_4);
    TP::commit_point(transcript_, "T_5", proof.T_5);
    TP::commit_point(transcript_, "T_6", proof.T_6);
    c->u = TP::challenge_scalar(transcript_, "u");                     // :314-315
    c->x = TP::challenge_scalar(transcript_, "x");
    TP::commit_scalar(transcript_, "t_x", proof.t_x);                  // :317-321
    TP::commit_scalar(transcript_, "t_x_blinding", proof.t_x_blinding);
    TP::commit_scalar(transcript_, "e_blinding", proof.e_blinding);
    c->w = TP::challenge_scalar(transcript_, "w");                     // :323
    return OK;
  }

 public:
  // The terms of the verification MSM (verifier.rs:394-449) computed entirely on the host, split by kind, for the batched
  // device check (bpgpu_msm_batch_is_identity): `fixed` are the scalars of [G[..N] | H[..N] | g | h], `var_*` the
  // proof's own points A_I1, A_O1, S1, A_I2, A_O2, S2, V.., T_1, T_3, T_4, T_5, T_6, L.., R.. with their scalars.
  // Same arithmetic as verify(); meant for small circuits verified in bulk, where a per-proof device round trip costs
  // more than these O(N) host loops.
  struct VerificationTerms { size_t padded_n; std::vector<FE> fixed; std::vector<G1<C>> var_points; std::vector<FE> var_scalars; };
  int verification_terms_host(const R1CSProof<C>& proof, size_t gens_len, const FE& rnd, VerificationTerms* out) {
    Challenges ch;
    int rc = replay_transcript(proof, gens_len, &ch);
    if (rc) return rc;
    const size_t n1 = ch.n1, n = ch.n, N = ch.padded_n;
    const FE y = ch.y, z = ch.z, u = ch.u, x = ch.x, w = ch.w;
    std::vector<FE> wL, wR, wO, wV;
    FE wc;
    flattened_constraints(z, &wL, &wR, &wO, &wV, &wc);                 // :325
    const FE a = proof.ipp_proof.a, b = proof.ipp_proof.b;
    // ipp.rs:262-312 on the host: challenges, one shared inversion, s[i] = s[i - 2^lg_i] * u^2_{(lg-1) - lg_i}
    const size_t lg = proof.ipp_proof.L.size();
    if (lg >= 32 || proof.ipp_proof.R.size() != lg || N != ((size_t)1 << lg)) return E_VERIFICATION;
    transcript_.innerproduct_domain_sep(N);
    std::vector<FE> uk(lg), uinv(lg);
    for (size_t k = 0; k < lg; k++) {
      TP::commit_point(transcript_, "L", proof.ipp_proof.L[k]);
      TP::commit_point(transcript_, "R", proof.ipp_proof.R[k]);
      uk[k] = TP::challenge_scalar(transcript_, "u");
    }
    FE prod_inv = FE::one();
    {                                                                  // FieldElement::batch_invert (ipp.rs:295); y^-1 rides along
      std::vector<FE> pre(lg + 1);
      FE run = y;
      for (size_t k = 0; k < lg; k++) { pre[k] = run; run = run * uk[k]; }
      FE inv = run.inverse();
      for (size_t k = lg; k-- > 0;) { uinv[k] = inv * pre[k]; inv = inv * uk[k]; }
      // inv is now y^-1 (0 if any factor was 0, as batch inversion of a zero would be)
      pre[lg] = inv;
      for (size_t k = 0; k < lg; k++) prod_inv = prod_inv * uinv[k];
      uinv.push_back(inv);
    }
    const FE y_inv = uinv[lg];
    std::vector<FE> s(N);
    s[0] = prod_inv;
    for (size_t i = 1; i < N; i++) {
      size_t lg_i = 0;
      while (((size_t)2 << lg_i) <= i) lg_i++;
      s[i] = s[i - ((size_t)1 << lg_i)] * uk[(lg - 1) - lg_i].square();
    }
    // verifier.rs:341-390
    out->padded_n = N;
    out->fixed.assign(2 * N + 2, FE::zero());
    FE yinv_i = FE::one(), delta = FE::zero();
    const FE one = FE::one();
    for (size_t i = 0; i < N; i++) {
      const FE wl = i < n ? wL[i] : FE::zero(), wr = i < n ? wR[i] : FE::zero(), wo = i < n ? wO[i] : FE::zero();
      const FE yw = wr * yinv_i;
      if (i < n) delta = delta + yw * wl;
      FE gs = x * yw - a * s[i];
      FE hs = yinv_i * (x * wl + wo - b * s[N - 1 - i]) - one;
      if (i >= n1) { gs = u * gs; hs = u * hs; }
      out->fixed[i] = gs;
      out->fixed[N + i] = hs;
      yinv_i = yinv_i * y_inv;
    }
    const FE xx = x.square(), xxx = x * xx;
    const FE r_xx = rnd * xx, rx = rnd * x, rx3 = rnd * xxx, rx4 = rx3 * x, rx5 = rx4 * x, rx6 = rx5 * x;
    out->fixed[2 * N] = w * (proof.t_x - a * b) + rnd * (xx * (wc + delta) - proof.t_x);      // scalar of g  (:421)
    out->fixed[2 * N + 1] = (proof.e_blinding + rnd * proof.t_x_blinding).negation();         // scalar of h  (:424)
    out->var_points = {proof.A_I1, proof.A_O1, proof.S1, proof.A_I2, proof.A_O2, proof.S2};
    out->var_scalars = {x, xx, xxx, u * x, u * xx, u * xxx};
    for (size_t j = 0; j < V_.size(); j++) { out->var_points.push_back(V_[j]); out->var_scalars.push_back(wV[j] * r_xx); }
    const G1<C>* T[5] = {&proof.T_1, &proof.T_3, &proof.T_4, &proof.T_5, &proof.T_6};
    const FE ts[5] = {rx, rx3, rx4, rx5, rx6};
    for (int k = 0; k < 5; k++) { out->var_points.push_back(*T[k]); out->var_scalars.push_back(ts[k]); }
    for (size_t k = 0; k < lg; k++) { out->var_points.push_back(proof.ipp_proof.L[k]); out->var_scalars.push_back(uk[k].square()); }
    for (size_t k = 0; k < lg; k++) { out->var_points.push_back(proof.ipp_proof.R[k]); out->var_scalars.push_back(uinv[k].square()); }
    return OK;
  }

 private:
  int verification_msm(const R1CSProof<C>& proof, const G1<C>& g, const G1<C>& h, const G1Vector<C>& G, const G1Vector<C>& H, const FE& rnd,
                       bool* is_identity, const bpgpu_circuit* recorded = nullptr) {
    const size_t mb = C::MODBYTES;
    Trace tr("verify");
    Challenges ch;
    int rc = replay_transcript(proof, G.len() < H.len() ? G.len() : H.len(), &ch);
    if (rc) return rc;
    const size_t n1 = ch.n1, n = ch.n, padded_n = ch.padded_n;
    const FE y = ch.y, z = ch.z, u = ch.u, x = ch.x, w = ch.w;
    std::vector<FE> wL, wR, wO, wV;
    FE wc;
    tr.mark("transcript");
    FieldElementVector<C> d_wL, d_wR, d_wO;
    if (recorded) {                                                    // :325 on the device; wV and wc (m + 1 scalars) come back
      uint8_t zb[C::MODBYTES];
      z.to_bytes(zb);
      bpgpu_scalars* hw = nullptr;
      if ((rc = bpgpu_circuit_flatten(ctx_, recorded, zb, &hw))) return rc;
      FieldElementVector<C> all = FieldElementVector<C>::adopt(ctx_, hw), d_tail;
      if ((rc = all.view(0, n, &d_wL)) || (rc = all.view(n, n, &d_wR)) || (rc = all.view(2 * n, n, &d_wO)) ||
          (rc = all.view(3 * n, V_.size() + 1, &d_tail)) || (rc = d_tail.to_host(&wV)))
        return rc;
      wc = wV.back();
      wV.pop_back();
    } else {
      flattened_constraints(z, &wL, &wR, &wO, &wV, &wc);               // :325
    }
    tr.mark("flattened_constraints");
    const FE a = proof.ipp_proof.a, b = proof.ipp_proof.b;

    // IPP challenges (transcript replay, host) and the s vector (device)  (:354-360)
    std::vector<FE> u_sq, u_inv_sq;
    FieldElementVector<C> s;
    if ((rc = IPP<C>::verification_scalars(ctx_, proof.ipp_proof.L, proof.ipp_proof.R, padded_n, transcript_, &u_sq, &u_inv_sq, &s))) return rc;

    tr.mark("ipp scalars");
    // g_scalars | h_scalars and delta on the device (:341-390)
    if (!recorded && (rc = FieldElementVector<C>::from_host_many(ctx_, {&wL, &wR, &wO}, {&d_wL, &d_wR, &d_wO}))) return rc;
    uint8_t yb[C::MODBYTES], xb[C::MODBYTES], ab[C::MODBYTES], bb[C::MODBYTES], ub[C::MODBYTES], deltab[C::MODBYTES];
    y.to_bytes(yb); x.to_bytes(xb); a.to_bytes(ab); b.to_bytes(bb); u.to_bytes(ub);
    bpgpu_scalars* h_gh = nullptr;
    if ((rc = bpgpu_r1cs_verifier_scalars(ctx_, n, n1, padded_n, d_wL.handle(), d_wR.handle(), d_wO.handle(), s.handle(), yb, xb, ab, bb, ub,
                                          &h_gh, deltab)))
      return rc;
    FieldElementVector<C> gh = FieldElementVector<C>::adopt(ctx_, h_gh);
    const FE delta = FE::from_bytes(deltab);

    tr.mark("verifier_scalars");
    // the fixed scalars of arg1 (:392-429)
    const FE xx = x.square(), xxx = x * xx;
    const FE r_xx = rnd * xx;
    const FE rx = rnd * x, rx3 = rnd * xxx, rx4 = rx3 * x, rx5 = rx4 * x, rx6 = rx5 * x;
    std::vector<FE> sc_head = {x, xx, xxx, u * x, u * xx, u * xxx};
    for (const FE& wv : wV) sc_head.push_back(wv * r_xx);                                     // :416-418
    sc_head.push_back(rx); sc_head.push_back(rx3); sc_head.push_back(rx4); sc_head.push_back(rx5); sc_head.push_back(rx6);
    sc_head.push_back(w * (proof.t_x - a * b) + rnd * (xx * (wc + delta) - proof.t_x));      // :421
    sc_head.push_back((proof.e_blinding + rnd * proof.t_x_blinding).negation());              // :424
    std::vector<G1<C>> pt_head = {proof.A_I1, proof.A_O1, proof.S1, proof.A_I2, proof.A_O2, proof.S2};   // :431-446
    pt_head.insert(pt_head.end(), V_.begin(), V_.end());
    pt_head.push_back(proof.T_1); pt_head.push_back(proof.T_3); pt_head.push_back(proof.T_4); pt_head.push_back(proof.T_5); pt_head.push_back(proof.T_6);
    pt_head.push_back(g); pt_head.push_back(h);
    std::vector<uint8_t> sc_head_b(sc_head.size() * mb);
    for (size_t i = 0; i < sc_head.size(); i++) sc_head[i].to_bytes(sc_head_b.data() + i * mb);
    const size_t lg = u_sq.size();
    std::vector<G1<C>> pt_tail(proof.ipp_proof.L);
    pt_tail.insert(pt_tail.end(), proof.ipp_proof.R.begin(), proof.ipp_proof.R.end());
    std::vector<uint8_t> sc_tail_b(2 * lg * mb + 1);
    for (size_t k = 0; k < lg; k++) { u_sq[k].to_bytes(sc_tail_b.data() + k * mb); u_inv_sq[k].to_bytes(sc_tail_b.data() + (lg + k) * mb); }

    // ONE MSM over [head | G | H | L | R]  (:451-452)
    bpgpu_msm_part parts[4] = {
        {nullptr, 0, pt_head[0].xy, nullptr, 0, sc_head_b.data(), pt_head.size()},
        {G.handle(), 0, nullptr, gh.handle(), 0, nullptr, padded_n},
        {H.handle(), 0, nullptr, gh.handle(), padded_n, nullptr, padded_n},
        {nullptr, 0, lg ? pt_tail[0].xy : nullptr, nullptr, 0, sc_tail_b.data(), 2 * lg},
    };
    G1<C> res;
    if ((rc = bpgpu_msm_parts(ctx_, parts, 4, res.xy))) return rc;
    tr.mark("msm");
    *is_identity = res.is_identity();
    return OK;
  }
};

}  // namespace bph
