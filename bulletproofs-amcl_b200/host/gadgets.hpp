// The range gadgets that drive the benchmark configurations, mirrored from
// /root/reference/src/r1cs/gadgets/helper_constraints/positive_no.rs:8-40, helper_constraints/mod.rs:16-22 and
// src/r1cs/gadgets/bound_check.rs:13-178.  Written against the abstract ConstraintSystem, so the same code
// builds the prover's and the verifier's circuit, as in the reference.
#pragma once
#include "r1cs.hpp"

namespace bph {

template <class C>
inline bool fe_bit(const FieldElement<C>& q, size_t i) {       // q.shift_right(i).is_odd()
  uint8_t be[C::MODBYTES];
  q.to_bytes(be);
  if (i >= 8 * (size_t)C::MODBYTES) return false;
  return (be[C::MODBYTES - 1 - i / 8] >> (i % 8)) & 1;
}

// positive_no.rs:8-40 -- v in [0, 2^n)
template <class C>
int positive_no_gadget(ConstraintSystem<C>& cs, const AllocatedQuantity<C>& v, size_t n) {
  using FE = FieldElement<C>;
  using LC = LinearCombination<C>;
  std::vector<std::pair<Variable, FE>> constraint_v = {{v.variable, FE::minus_one()}};
  FE exp_2 = FE::one();
  const FE zero = FE::zero(), one = FE::one(), minus_one = FE::minus_one();
  constraint_v.reserve(n + 1);
  for (size_t i = 0; i < n; i++) {
    Variable a, b, o;
    int rc;
    if (v.has_assignment) {
      bool odd = fe_bit(v.assignment, i);                        // :18-24: (left, right) = (1 - bit, bit)
      rc = cs.allocate_multiplier(odd ? &zero : &one, odd ? &one : &zero, &a, &b, &o);
    } else {
      rc = cs.allocate_multiplier(nullptr, nullptr, &a, &b, &o);
    }
    if (rc) return rc;
    cs.constrain(LC(o));                                         // :27   a * b = 0
    // :30   a = 1 - b, i.e. the terms of `a + (b - 1)` in the reference's order, built without the temporaries
    cs.constrain(LC(std::vector<std::pair<Variable, FE>>{{a, one}, {b, one}, {Variable::one(), minus_one}}));
    constraint_v.emplace_back(b, exp_2);
    exp_2 = exp_2 + exp_2;
  }
  cs.constrain(LC(constraint_v));                                // :37   sum b_i 2^i = v
  return OK;
}

template <class C>
void constrain_lc_with_scalar(ConstraintSystem<C>& cs, const LinearCombination<C>& lc, const FieldElement<C>& scalar) {   // helper_constraints/mod.rs:16-22
  cs.constrain(lc - LinearCombination<C>(scalar));
}

// bound_check.rs:13-39 -- v in [min, max] via a = v - min, b = max - v, both n-bit
template <class C>
int bound_check_gadget(ConstraintSystem<C>& cs, const AllocatedQuantity<C>& v, const AllocatedQuantity<C>& a, const AllocatedQuantity<C>& b,
                       uint64_t max, uint64_t min, size_t n) {
  using FE = FieldElement<C>;
  using LC = LinearCombination<C>;
  cs.constrain(LC(v.variable) - LC(FE::from_u64(min)) - LC(a.variable));       // :26
  cs.constrain(LC(FE::from_u64(max)) - LC(v.variable) - LC(b.variable));       // :28
  constrain_lc_with_scalar<C>(cs, LC(a.variable) + LC(b.variable), FE::from_u64(max - min));   // :31
  int rc = positive_no_gadget<C>(cs, a, n);                                     // :34
  if (rc) return rc;
  return positive_no_gadget<C>(cs, b, n);                                       // :36
}

// bound_check.rs:41-92.  `randomness` = blinding of the commitment to val (None -> drawn from rng);
// the blindings of a and b are FieldElement::random() in the reference, here the next two draws of `rng`.
template <class C>
int prove_bounded_num(uint64_t val, const FieldElement<C>* randomness, uint64_t lower, uint64_t upper, size_t max_bits_in_val, Rng<C>& rng,
                      Prover<C>& prover, std::vector<G1<C>>* comms) {
  using FE = FieldElement<C>;
  const uint64_t a = val - lower, b = upper - val;
  comms->clear();
  const FE vals[3] = {FE::from_u64(val), FE::from_u64(a), FE::from_u64(b)};
  AllocatedQuantity<C> q[3];
  for (int k = 0; k < 3; k++) {
    FE blind = (k == 0 && randomness) ? *randomness : rng.next();
    G1<C> com;
    Variable var;
    int rc = prover.commit(vals[k], blind, &com, &var);
    if (rc) return rc;
    q[k] = {var, true, vals[k]};
    comms->push_back(com);
  }
  return bound_check_gadget<C>(prover, q[0], q[1], q[2], upper, lower, max_bits_in_val);
}

// bound_check.rs:94-129
template <class C>
int verify_bounded_num(uint64_t lower, uint64_t upper, size_t max_bits_in_val, const std::vector<G1<C>>& commitments, Verifier<C>& verifier) {
  if (commitments.size() < 3) return E_FORMAT;
  AllocatedQuantity<C> q[3];
  for (int k = 0; k < 3; k++) q[k] = {verifier.commit(commitments[k]), false, FieldElement<C>::zero()};
  return bound_check_gadget<C>(verifier, q[0], q[1], q[2], upper, lower, max_bits_in_val);
}

// bound_check.rs:133-161
template <class C>
int gen_proof_of_bounded_num(bpgpu_ctx* ctx, uint64_t val, const FieldElement<C>* randomness, uint64_t lower, uint64_t upper,
                             size_t max_bits_in_val, Rng<C>& rng, const std::string& transcript_label, const G1<C>& g, const G1<C>& h,
                             const G1Vector<C>& G, const G1Vector<C>& H, R1CSProof<C>* proof, std::vector<G1<C>>* comms) {
  Transcript prover_transcript(transcript_label);
  Prover<C> prover(ctx, g, h, prover_transcript, rng);
  int rc = prove_bounded_num<C>(val, randomness, lower, upper, max_bits_in_val, rng, prover, comms);
  if (rc) return rc;
  return prover.prove(G, H, proof);
}

// bound_check.rs:163-178
template <class C>
int verify_proof_of_bounded_num(bpgpu_ctx* ctx, uint64_t lower, uint64_t upper, size_t max_bits_in_val, const R1CSProof<C>& proof,
                                const std::vector<G1<C>>& commitments, const std::string& transcript_label, const G1<C>& g, const G1<C>& h,
                                const G1Vector<C>& G, const G1Vector<C>& H, const FieldElement<C>& verifier_r) {
  Transcript verifier_transcript(transcript_label);
  Verifier<C> verifier(ctx, verifier_transcript);
  int rc = verify_bounded_num<C>(lower, upper, max_bits_in_val, commitments, verifier);
  if (rc) return rc;
  return verifier.verify(proof, g, h, G, H, verifier_r);
}

// Aggregated range statement used by the benchmark configurations (BASELINE.json configs 2, 3, 5): m committed
// values, each constrained to [0, 2^bits) by one positive_no_gadget -> n = m * bits multipliers.
template <class C>
int gen_proof_of_positive_nums(bpgpu_ctx* ctx, const std::vector<uint64_t>& vals, size_t bits, Rng<C>& rng, const std::string& transcript_label,
                               const G1<C>& g, const G1<C>& h, const G1Vector<C>& G, const G1Vector<C>& H, R1CSProof<C>* proof,
                               std::vector<G1<C>>* comms) {
  using FE = FieldElement<C>;
  Trace tr("range_prove");
  Transcript prover_transcript(transcript_label);
  Prover<C> prover(ctx, g, h, prover_transcript, rng);
  // the reference would call prover.commit(v_j, FieldElement::random()) then the gadget, per value; the commitments do
  // not depend on the gadget's allocations, so they are evaluated as one batch with the same transcript order
  std::vector<FE> fv, blind;
  for (uint64_t v : vals) { fv.push_back(FE::from_u64(v)); blind.push_back(rng.next()); }
  std::vector<Variable> vars;
  int rc = prover.commit_vec(fv, blind, comms, &vars);
  if (rc) return rc;
  for (size_t j = 0; j < vals.size(); j++)
    if ((rc = positive_no_gadget<C>(prover, AllocatedQuantity<C>{vars[j], true, fv[j]}, bits))) return rc;
  tr.mark("commit + gadget");
  return prover.prove(G, H, proof);
}

// The same proof from a circuit recorded earlier (bpgpu_circuit of m positive_no gadgets) and a witness built on the
// device: what a caller proving the same statement shape over and over does instead of re-running the gadget per proof.
// The draws (one blinding per value, then prove()'s) and the transcript are those of gen_proof_of_positive_nums.
template <class C>
int gen_proof_of_positive_nums_recorded(bpgpu_ctx* ctx, const bpgpu_circuit* circuit, const std::vector<uint64_t>& vals, size_t bits, Rng<C>& rng,
                                        const std::string& transcript_label, const G1<C>& g, const G1<C>& h, const G1Vector<C>& G,
                                        const G1Vector<C>& H, R1CSProof<C>* proof, std::vector<G1<C>>* comms) {
  using FE = FieldElement<C>;
  Trace tr("range_prove");
  Transcript prover_transcript(transcript_label);
  Prover<C> prover(ctx, g, h, prover_transcript, rng);
  std::vector<FE> fv, blind;
  for (uint64_t v : vals) { fv.push_back(FE::from_u64(v)); blind.push_back(rng.next()); }
  std::vector<Variable> vars;
  int rc = prover.commit_vec(fv, blind, comms, &vars);
  if (rc) return rc;
  bpgpu_scalars* wit = nullptr;
  if ((rc = bpgpu_range_witness(ctx, vals.data(), vals.size(), bits, &wit))) return rc;
  tr.mark("commit + device witness");
  typename Prover<C>::DeviceCircuit dc{circuit, wit, vals.size() * bits};
  rc = prover.prove(G, H, proof, &dc);
  bpgpu_scalars_free(wit);
  return rc;
}

template <class C>
int verify_proof_of_positive_nums(bpgpu_ctx* ctx, size_t bits, const R1CSProof<C>& proof, const std::vector<G1<C>>& commitments,
                                  const std::string& transcript_label, const G1<C>& g, const G1<C>& h, const G1Vector<C>& G,
                                  const G1Vector<C>& H, const FieldElement<C>& verifier_r) {
  Trace tr("range_verify");
  Transcript verifier_transcript(transcript_label);
  Verifier<C> verifier(ctx, verifier_transcript);
  for (const auto& com : commitments) {
    Variable var = verifier.commit(com);
    int rc = positive_no_gadget<C>(verifier, AllocatedQuantity<C>{var, false, FieldElement<C>::zero()}, bits);
    if (rc) return rc;
  }
  tr.mark("commit + gadget");
  return verifier.verify(proof, g, h, G, H, verifier_r);
}

// the verifier's side of gen_proof_of_positive_nums_recorded: commitments into the transcript, weights from the recorded circuit
template <class C>
int verify_proof_of_positive_nums_recorded(bpgpu_ctx* ctx, const bpgpu_circuit* circuit, const R1CSProof<C>& proof,
                                           const std::vector<G1<C>>& commitments, const std::string& transcript_label, const G1<C>& g,
                                           const G1<C>& h, const G1Vector<C>& G, const G1Vector<C>& H, const FieldElement<C>& verifier_r) {
  Trace tr("range_verify");
  Transcript verifier_transcript(transcript_label);
  Verifier<C> verifier(ctx, verifier_transcript);
  for (const auto& com : commitments) verifier.commit(com);
  tr.mark("commit");
  return verifier.verify(proof, g, h, G, H, verifier_r, circuit);
}

// k-shuffle: {y} is a permutation of {x}.  The TWO-PHASE gadget of the reference's ConstraintSystem documentation
// (constraint_system.rs:86-135): the challenge z exists only after the first-phase commitments A_I1, A_O1, S1, and
// prod (x_i - z) = prod (y_i - z) costs 2(k-1) second-phase multipliers -- the path through
// specify_randomized_constraints, RandomizedConstraintSystem::challenge_scalar and the A_I2 / A_O2 / S2 commitments
// (prover.rs:300-319,384-436; verifier.rs:245-264).
template <class C>
int shuffle_gadget(ConstraintSystem<C>& cs, const std::vector<Variable>& x, const std::vector<Variable>& y) {
  using FE = FieldElement<C>;
  using LC = LinearCombination<C>;
  if (x.size() != y.size() || x.empty()) return E_ARG;
  return cs.specify_randomized_constraints([x, y](ConstraintSystem<C>& cs) -> int {
    const size_t k = x.size();
    FE z;
    int rc = cs.challenge_scalar("shuffle challenge", &z);
    if (rc) return rc;
    if (k == 1) { cs.constrain(LC(y[0]) - LC(x[0])); return OK; }
    auto chain = [&](const std::vector<Variable>& v) {
      Variable l, r, o;
      cs.multiply(LC(v[k - 1]) - LC(z), LC(v[k - 2]) - LC(z), &l, &r, &o);
      for (size_t i = k - 2; i-- > 0;) cs.multiply(LC(o), LC(v[i]) - LC(z), &l, &r, &o);
      return o;
    };
    const Variable ox = chain(x), oy = chain(y);
    cs.constrain(LC(ox) - LC(oy));
    return OK;
  });
}

// shuffle statement over 2k committed values, with x[0] additionally range-checked to `bits` bits in the FIRST phase
// (bits = 0: no first-phase multipliers at all) -- a circuit with n1 = bits and n2 = 2(k-1)
template <class C>
int gen_proof_of_shuffle(bpgpu_ctx* ctx, const std::vector<uint64_t>& xs, const std::vector<uint64_t>& ys, size_t bits, Rng<C>& rng,
                         const std::string& transcript_label, const G1<C>& g, const G1<C>& h, const G1Vector<C>& G, const G1Vector<C>& H,
                         R1CSProof<C>* proof, std::vector<G1<C>>* comms) {
  using FE = FieldElement<C>;
  Transcript prover_transcript(transcript_label);
  Prover<C> prover(ctx, g, h, prover_transcript, rng);
  std::vector<FE> fv, blind;
  for (uint64_t v : xs) { fv.push_back(FE::from_u64(v)); blind.push_back(rng.next()); }
  for (uint64_t v : ys) { fv.push_back(FE::from_u64(v)); blind.push_back(rng.next()); }
  std::vector<Variable> vars;
  int rc = prover.commit_vec(fv, blind, comms, &vars);
  if (rc) return rc;
  const size_t k = xs.size();
  if (bits && (rc = positive_no_gadget<C>(prover, AllocatedQuantity<C>{vars[0], true, fv[0]}, bits))) return rc;
  std::vector<Variable> vx(vars.begin(), vars.begin() + k), vy(vars.begin() + k, vars.end());
  if ((rc = shuffle_gadget<C>(prover, vx, vy))) return rc;
  return prover.prove(G, H, proof);
}

template <class C>
int verify_proof_of_shuffle(bpgpu_ctx* ctx, size_t k, size_t bits, const R1CSProof<C>& proof, const std::vector<G1<C>>& commitments,
                            const std::string& transcript_label, const G1<C>& g, const G1<C>& h, const G1Vector<C>& G, const G1Vector<C>& H,
                            const FieldElement<C>& verifier_r) {
  if (commitments.size() != 2 * k || k == 0) return E_ARG;
  Transcript verifier_transcript(transcript_label);
  Verifier<C> verifier(ctx, verifier_transcript);
  std::vector<Variable> vars;
  for (const auto& com : commitments) vars.push_back(verifier.commit(com));
  int rc;
  if (bits && (rc = positive_no_gadget<C>(verifier, AllocatedQuantity<C>{vars[0], false, FieldElement<C>::zero()}, bits))) return rc;
  std::vector<Variable> vx(vars.begin(), vars.begin() + k), vy(vars.begin() + k, vars.end());
  if ((rc = shuffle_gadget<C>(verifier, vx, vy))) return rc;
  return verifier.verify(proof, g, h, G, H, verifier_r);
}

}  // namespace bph
