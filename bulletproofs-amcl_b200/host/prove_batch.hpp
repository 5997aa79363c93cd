// Lock-step proving of a batch of independent range proofs (m x positive_no_gadget each) over bpgpu_pbatch_*
// (csrc/provebatch.cu): the stages of Prover::prove (/root/reference/src/r1cs/prover.rs:322-593) and the rounds of
// IPP::create_ipp (src/ipp.rs:68-194) are executed for ALL proofs of a slab by one device call each, the transcripts
// of the proofs advance on host threads in between.  Proof i is byte-identical to the proof gen_proof_of_positive_nums
// produces with the same Rng: same draws in the same order, same transcript, same group elements.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

#include "gadgets.hpp"

namespace bph {

template <class C>
struct BatchProverAccess {
  using FE = FieldElement<C>;
  using TP = TranscriptProtocol<C>;
  struct State {
    Transcript tr;
    Rng<C> rng;
    Prover<C> prover;
    R1CSProof<C> proof;
    std::vector<FE> wV, uk, uinv;
    FE i_b, o_b, s_b, y, z, u, x, w;
    FE t[7], tb[7];
    State(bpgpu_ctx* ctx, const G1<C>& g, const G1<C>& h, const std::string& label, Rng<C> r)
        : tr(label), rng(std::move(r)), prover(ctx, g, h, tr, rng) {}
  };

  // worker threads that live for the whole call: a slab runs ~15 short host stages, and starting threads for each of them
  // costs as much as the stage
  class Pool {
   public:
    explicit Pool(size_t nthreads) {
      for (size_t k = 1; k < nthreads; k++) th_.emplace_back([this]() { loop(); });
    }
    ~Pool() {
      { std::lock_guard<std::mutex> l(m_); stop_ = true; gen_++; }
      cv_.notify_all();
      for (auto& t : th_) t.join();
    }
    template <class F>
    void run(size_t count, F fn) {
      std::function<void(size_t)> f = fn;
      {
        std::lock_guard<std::mutex> l(m_);
        fn_ = &f; count_ = count; next_.store(0); active_ = th_.size(); gen_++;
      }
      cv_.notify_all();
      work();
      std::unique_lock<std::mutex> l(m_);
      done_.wait(l, [this]() { return active_ == 0; });
      fn_ = nullptr;
    }

   private:
    void work() { for (;;) { size_t i = next_.fetch_add(1); if (i >= count_) break; (*fn_)(i); } }
    void loop() {
      uint64_t seen = 0;
      for (;;) {
        {
          std::unique_lock<std::mutex> l(m_);
          cv_.wait(l, [&]() { return gen_ != seen; });
          seen = gen_;
          if (stop_) return;
        }
        work();
        { std::lock_guard<std::mutex> l(m_); if (--active_ == 0) done_.notify_one(); }
      }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::function<void(size_t)>* fn_ = nullptr;
    size_t count_ = 0, active_ = 0;
    std::atomic<size_t> next_{0};
    uint64_t gen_ = 0;
    bool stop_ = false;
  };

  // one slab of B proofs
  static int prove_slab(bpgpu_ctx* ctx, bpgpu_pbatch* pb, const std::string& label, const G1<C>& g, const G1<C>& h, const uint64_t* values, size_t B,
                        size_t m, size_t bits, int rng_mode, uint64_t seed0, Pool& pool, uint8_t* proofs, size_t stride, uint8_t* comms_xy) {
    const size_t mb = C::MODBYTES, pbts = 2 * mb;
    const size_t n = m * bits, N = next_power_of_two(n);
    size_t lg = 0;
    while (((size_t)1 << lg) < N) lg++;
    std::vector<std::unique_ptr<State>> st(B);
    for (size_t i = 0; i < B; i++)
      st[i].reset(new State(ctx, g, h, label, rng_mode == 1 ? Rng<C>(seed0 + i, "blind") : Rng<C>()));
    for (size_t i = 0; i < B; i++) if (!st[i]->rng.ok()) return E_ENTROPY;
    int rc;
    Trace tr_("prove_slab");
    std::atomic<int> err{0};
    auto fail = [&](int e) { int z = 0; err.compare_exchange_strong(z, e); };
    bpgpu_fixed_bases* fb = nullptr;
    uint8_t pair[4 * 48];
    memcpy(pair, g.xy, pbts);
    memcpy(pair + pbts, h.xy, pbts);
    if ((rc = bpgpu_fixed_bases_get(ctx, pair, 2, &fb))) return rc;

    // ---- V_j = commit(v_j, blinding_j) for every value of every proof: one call (prover.rs:119-129)
    std::vector<FE> fv(B * m), blind(B * m);
    std::vector<uint8_t> sc(B * m * 2 * mb);
    for (size_t i = 0; i < B; i++)
      for (size_t j = 0; j < m; j++) {
        fv[i * m + j] = FE::from_u64(values[i * m + j]);
        blind[i * m + j] = st[i]->rng.next();
        fv[i * m + j].to_bytes(sc.data() + (i * m + j) * 2 * mb);
        blind[i * m + j].to_bytes(sc.data() + (i * m + j) * 2 * mb + mb);
      }
    if ((rc = bpgpu_fixed_bases_commit(ctx, fb, sc.data(), B * m, comms_xy))) return rc;
    tr_.mark("states + V commitments");

    // ---- circuits, first-phase blindings, witness
    std::vector<uint8_t> witness(B * 3 * n * mb), blind3(B * 3 * mb), keys(B * 64, 0);
    std::vector<uint64_t> ctr0(B);
    size_t key_len = st[0]->rng.key_len();
    pool.run(B, [&](size_t i) {
      State& s = *st[i];
      Prover<C>& p = s.prover;
      for (size_t j = 0; j < m; j++) {
        Variable var = p.commit_precomputed(fv[i * m + j], blind[i * m + j], G1<C>::from_xy(comms_xy + (i * m + j) * pbts));
        int e = positive_no_gadget<C>(p, AllocatedQuantity<C>{var, true, fv[i * m + j]}, bits);
        if (e) { fail(e); return; }
      }
      p.transcript_.append_u64("m", p.v_.size());                        // prover.rs:327
      if (p.a_L_.size() != n || !p.deferred_.empty()) { fail(E_ARG); return; }
      s.i_b = s.rng.next(); s.o_b = s.rng.next(); s.s_b = s.rng.next();  // :336-338
      memcpy(keys.data() + i * 64, s.rng.key(), s.rng.key_len());
      ctr0[i] = s.rng.counter();                                          // s_L (:340) then s_R (:341): 2n draws, made on the device
      s.rng.skip(2 * n);
      uint8_t* w = witness.data() + i * 3 * n * mb;
      for (size_t k = 0; k < n; k++) {
        p.a_L_[k].to_bytes(w + k * mb);
        p.a_R_[k].to_bytes(w + (n + k) * mb);
        p.a_O_[k].to_bytes(w + (2 * n + k) * mb);
      }
      s.i_b.to_bytes(blind3.data() + (i * 3) * mb);
      s.o_b.to_bytes(blind3.data() + (i * 3 + 1) * mb);
      s.s_b.to_bytes(blind3.data() + (i * 3 + 2) * mb);
    });
    tr_.mark("host: gadgets, witness");
    if (err.load()) return err.load();
    // keys were copied with their own stride of 64; the device call wants them packed by key_len
    std::vector<uint8_t> kp(B * key_len);
    for (size_t i = 0; i < B; i++) memcpy(kp.data() + i * key_len, keys.data() + i * 64, key_len);
    std::vector<uint8_t> pts3(B * 3 * pbts);
    if ((rc = bpgpu_pbatch_commit3(pb, witness.data(), kp.data(), key_len, ctr0.data(), blind3.data(), pts3.data()))) return rc;
    tr_.mark("device: A_I A_O S");

    // ---- commit A_I1, A_O1, S1; one-phase separator; y, z; flattened constraints (prover.rs:364-441)
    std::vector<uint8_t> weights(B * 3 * n * mb), yb(B * 2 * mb);
    pool.run(B, [&](size_t i) {
      State& s = *st[i];
      Prover<C>& p = s.prover;
      s.proof.A_I1 = G1<C>::from_xy(pts3.data() + (i * 3) * pbts);
      s.proof.A_O1 = G1<C>::from_xy(pts3.data() + (i * 3 + 1) * pbts);
      s.proof.S1 = G1<C>::from_xy(pts3.data() + (i * 3 + 2) * pbts);
      TP::commit_point(p.transcript_, "A_I1", s.proof.A_I1);
      TP::commit_point(p.transcript_, "A_O1", s.proof.A_O1);
      TP::commit_point(p.transcript_, "S1", s.proof.S1);
      int e = p.create_randomized_constraints();
      if (e) { fail(e); return; }
      s.proof.A_I2 = s.proof.A_O2 = s.proof.S2 = G1<C>::identity();       // :429
      TP::commit_point(p.transcript_, "A_I2", s.proof.A_I2);
      TP::commit_point(p.transcript_, "A_O2", s.proof.A_O2);
      TP::commit_point(p.transcript_, "S2", s.proof.S2);
      s.y = TP::challenge_scalar(p.transcript_, "y");
      s.z = TP::challenge_scalar(p.transcript_, "z");
      std::vector<FE> wL, wR, wO;
      p.flattened_constraints(s.z, &wL, &wR, &wO, &s.wV);
      uint8_t* w = weights.data() + i * 3 * n * mb;
      for (size_t k = 0; k < n; k++) {
        wL[k].to_bytes(w + k * mb);
        wR[k].to_bytes(w + (n + k) * mb);
        wO[k].to_bytes(w + (2 * n + k) * mb);
      }
    });
    {                                                                    // y^-1 of every proof with one inversion
      std::vector<FE> pre(B + 1);
      pre[0] = FE::one();
      for (size_t i = 0; i < B; i++) pre[i + 1] = st[i]->y.is_zero() ? pre[i] : pre[i] * st[i]->y;
      FE inv = pre[B].inverse();
      for (size_t i = B; i-- > 0;) {
        const FE& y = st[i]->y;
        FE yi = FE::zero();
        if (!y.is_zero()) { yi = inv * pre[i]; inv = inv * y; }
        y.to_bytes(yb.data() + (i * 2) * mb);
        yi.to_bytes(yb.data() + (i * 2 + 1) * mb);
      }
    }
    if (err.load()) return err.load();
    tr_.mark("host: y z, flatten");
    std::vector<uint8_t> tbe(B * 6 * mb);
    if ((rc = bpgpu_pbatch_polys(pb, weights.data(), yb.data(), tbe.data()))) return rc;
    tr_.mark("device: polys, t");

    // ---- T_1, T_3, T_4, T_5, T_6 for every proof: one call (prover.rs:490-500)
    std::vector<uint8_t> tsc(B * 5 * 2 * mb), tpts(B * 5 * pbts);
    pool.run(B, [&](size_t i) {
      State& s = *st[i];
      for (int k = 1; k <= 6; k++) s.t[k] = FE::from_bytes(tbe.data() + (i * 6 + (k - 1)) * mb);
      s.tb[1] = s.rng.next(); s.tb[3] = s.rng.next(); s.tb[4] = s.rng.next(); s.tb[5] = s.rng.next(); s.tb[6] = s.rng.next();
      const int idx[5] = {1, 3, 4, 5, 6};
      for (int k = 0; k < 5; k++) {
        s.t[idx[k]].to_bytes(tsc.data() + ((i * 5 + k) * 2) * mb);
        s.tb[idx[k]].to_bytes(tsc.data() + ((i * 5 + k) * 2 + 1) * mb);
      }
    });
    if ((rc = bpgpu_fixed_bases_commit(ctx, fb, tsc.data(), B * 5, tpts.data()))) return rc;
    tr_.mark("T commitments");

    // ---- u, x, t_x, blindings, w (prover.rs:502-549)
    std::vector<uint8_t> xuw(B * 3 * mb);
    pool.run(B, [&](size_t i) {
      State& s = *st[i];
      Prover<C>& p = s.prover;
      G1<C>* T[5] = {&s.proof.T_1, &s.proof.T_3, &s.proof.T_4, &s.proof.T_5, &s.proof.T_6};
      const char* lab[5] = {"T_1", "T_3", "T_4", "T_5", "T_6"};
      for (int k = 0; k < 5; k++) { *T[k] = G1<C>::from_xy(tpts.data() + (i * 5 + k) * pbts); TP::commit_point(p.transcript_, lab[k], *T[k]); }
      s.u = TP::challenge_scalar(p.transcript_, "u");
      s.x = TP::challenge_scalar(p.transcript_, "x");
      const FE x = s.x;
      s.tb[2] = FE::zero();                                               // t_2_blinding = <wV, v_blinding> (:513)
      for (size_t j = 0; j < s.wV.size(); j++) s.tb[2] = s.tb[2] + s.wV[j] * p.v_blinding_[j];
      auto poly6 = [&](const FE* c) { return x * (c[1] + x * (c[2] + x * (c[3] + x * (c[4] + x * (c[5] + x * c[6]))))); };
      s.proof.t_x = poly6(s.t);
      s.proof.t_x_blinding = poly6(s.tb);
      s.proof.e_blinding = x * (s.i_b + x * (s.o_b + x * s.s_b));          // :537-541 with the second-phase blindings zero
      TP::commit_scalar(p.transcript_, "t_x", s.proof.t_x);
      TP::commit_scalar(p.transcript_, "t_x_blinding", s.proof.t_x_blinding);
      TP::commit_scalar(p.transcript_, "e_blinding", s.proof.e_blinding);
      s.w = TP::challenge_scalar(p.transcript_, "w");
      s.x.to_bytes(xuw.data() + (i * 3) * mb);
      s.u.to_bytes(xuw.data() + (i * 3 + 1) * mb);
      s.w.to_bytes(xuw.data() + (i * 3 + 2) * mb);
      p.transcript_.innerproduct_domain_sep(N);                           // ipp.rs:62
    });
    if ((rc = bpgpu_pbatch_eval(pb, xuw.data()))) return rc;
    tr_.mark("host: u x w; device: eval");

    // ---- IPP rounds (ipp.rs:68-194)
    std::vector<uint8_t> uv(B * 2 * mb), lr(B * 2 * pbts);
    for (size_t k = 0; k < lg; k++) {
      if ((rc = bpgpu_pbatch_ipp_round(pb, k ? uv.data() : nullptr, lr.data()))) return rc;
      pool.run(B, [&](size_t i) {
        State& s = *st[i];
        const G1<C> L = G1<C>::from_xy(lr.data() + (i * 2) * pbts), R = G1<C>::from_xy(lr.data() + (i * 2 + 1) * pbts);
        TP::commit_point(s.prover.transcript_, "L", L);
        TP::commit_point(s.prover.transcript_, "R", R);
        s.proof.ipp_proof.L.push_back(L);
        s.proof.ipp_proof.R.push_back(R);
        s.uk.push_back(TP::challenge_scalar(s.prover.transcript_, "u"));
      });
      // u^-1 for all proofs with ONE inversion (Montgomery's trick); a zero challenge (probability 2^-255) inverts to zero
      // as FieldElement::inverse does
      std::vector<FE> pre(B + 1);
      pre[0] = FE::one();
      for (size_t i = 0; i < B; i++) { const FE& u = st[i]->uk.back(); pre[i + 1] = u.is_zero() ? pre[i] : pre[i] * u; }
      FE inv = pre[B].inverse();
      for (size_t i = B; i-- > 0;) {
        const FE& u = st[i]->uk.back();
        FE ui = FE::zero();
        if (!u.is_zero()) { ui = inv * pre[i]; inv = inv * u; }
        u.to_bytes(uv.data() + (i * 2) * mb);
        ui.to_bytes(uv.data() + (i * 2 + 1) * mb);
      }
    }
    tr_.mark("IPP rounds");
    std::vector<uint8_t> ab(B * 2 * mb);
    if ((rc = bpgpu_pbatch_ipp_finish(pb, uv.data(), ab.data()))) return rc;
    for (size_t i = 0; i < B; i++) {
      State& s = *st[i];
      s.proof.ipp_proof.a = FE::from_bytes(ab.data() + (i * 2) * mb);
      s.proof.ipp_proof.b = FE::from_bytes(ab.data() + (i * 2 + 1) * mb);
      std::vector<uint8_t> bytes = s.proof.to_bytes();
      if (bytes.size() > stride) return BPH_E_BUFFER;
      memcpy(proofs + i * stride, bytes.data(), bytes.size());
    }
    pool.run(B, [&](size_t i) { st[i].reset(); });      // a constraint system is hundreds of small allocations: release them in parallel
    return OK;
  }
};

}  // namespace bph
