// C entry points of the host layer (include/bphost.h): flatten the C++ mirror of the reference's API to bytes.
#include <stdio.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bphost.h"
#include "gadgets.hpp"
#include "prove_batch.hpp"

using namespace bph;

namespace {

template <class C>
std::vector<FieldElement<C>> scalars_from(const uint8_t* be, size_t n) {
  std::vector<FieldElement<C>> v(n);
  for (size_t i = 0; i < n; i++) v[i] = FieldElement<C>::from_bytes(be + i * C::MODBYTES);
  return v;
}

int emit(const std::vector<uint8_t>& bytes, uint8_t* out, size_t cap, size_t* len) {
  if (len) *len = bytes.size();
  if (!out || cap < bytes.size()) return BPH_E_BUFFER;
  memcpy(out, bytes.data(), bytes.size());
  return BPGPU_OK;
}

template <class C>
std::vector<uint8_t> ipp_to_bytes(const InnerProductArgumentProof<C>& p) {
  std::vector<uint8_t> out;
  for (const auto& q : p.L) { auto b = q.to_bytes(); out.insert(out.end(), b.begin(), b.end()); }
  for (const auto& q : p.R) { auto b = q.to_bytes(); out.insert(out.end(), b.begin(), b.end()); }
  append_fe(out, p.a);
  append_fe(out, p.b);
  return out;
}

template <class C>
int ipp_from_bytes(const uint8_t* buf, size_t len, InnerProductArgumentProof<C>* p) {
  const size_t PB = 1 + 2 * C::MODBYTES, SB = C::MODBYTES;
  if (len < 2 * SB || (len - 2 * SB) % (2 * PB)) return BPGPU_E_FORMAT;
  size_t lg = (len - 2 * SB) / (2 * PB);
  p->L.resize(lg);
  p->R.resize(lg);
  size_t off = 0;
  auto point = [&](G1<C>* q) {
    if (buf[off] != 4 || !G1<C>::xy_valid(buf + off + 1)) return false;
    *q = G1<C>::from_xy(buf + off + 1);
    off += PB;
    return true;
  };
  auto scalar = [&](FieldElement<C>* s) {              // canonical (< r), as R1CSProof::from_bytes requires
    *s = FieldElement<C>::from_bytes(buf + off);
    uint8_t chk[C::MODBYTES];
    s->to_bytes(chk);
    const bool ok = memcmp(chk, buf + off, SB) == 0;
    off += SB;
    return ok;
  };
  for (size_t k = 0; k < lg; k++) if (!point(&p->L[k])) return BPGPU_E_FORMAT;
  for (size_t k = 0; k < lg; k++) if (!point(&p->R[k])) return BPGPU_E_FORMAT;
  if (!scalar(&p->a) || !scalar(&p->b)) return BPGPU_E_FORMAT;
  return BPGPU_OK;
}

// every point of a caller-supplied list is a point (ECP::frombytes): else BPGPU_E_FORMAT
template <class C>
int points_valid(const uint8_t* xy, size_t n) {
  for (size_t k = 0; k < n; k++) if (!G1<C>::xy_valid(xy + k * 2 * C::MODBYTES)) return BPGPU_E_FORMAT;
  return BPGPU_OK;
}

template <class C>
Rng<C> make_rng(int mode, uint64_t seed) { return mode == 1 ? Rng<C>(seed, "blind") : Rng<C>(); }

// the verifier's FieldElement::random() (verifier.rs:392): explicit, or one draw of an OS-entropy stream
template <class C>
int verifier_scalar(const uint8_t* r_be, FieldElement<C>* out) {
  if (r_be) { *out = FieldElement<C>::from_bytes(r_be); return BPGPU_OK; }
  Rng<C> os;
  if (!os.ok()) return BPH_E_ENTROPY;
  *out = os.next();
  return BPGPU_OK;
}

template <class C>
int ipp_create_t(bpgpu_ctx* ctx, const char* label, const bpgpu_points* G, const bpgpu_points* H, const uint8_t* Q_xy, const uint8_t* Gf,
                 const uint8_t* Hf, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* proof, size_t cap, size_t* len) {
  FieldElementVector<C> dGf, dHf, da, db;
  int rc;
  if ((rc = FieldElementVector<C>::from_bytes(ctx, Gf, n, &dGf)) || (rc = FieldElementVector<C>::from_bytes(ctx, Hf, n, &dHf)) ||
      (rc = FieldElementVector<C>::from_bytes(ctx, a, n, &da)) || (rc = FieldElementVector<C>::from_bytes(ctx, b, n, &db)))
    return rc;
  G1Vector<C> vG = G1Vector<C>::borrow(ctx, G), vH = G1Vector<C>::borrow(ctx, H);
  Transcript t{std::string(label)};
  InnerProductArgumentProof<C> p;
  if ((rc = IPP<C>::create_ipp(ctx, t, G1<C>::from_xy(Q_xy), dGf, dHf, vG, 0, vH, 0, da, db, n, &p))) return rc;
  return emit(ipp_to_bytes(p), proof, cap, len);
}

template <class C>
int ipp_verify_t(bpgpu_ctx* ctx, const char* label, size_t n, const uint8_t* Gf, const uint8_t* Hf, const uint8_t* P_xy, const uint8_t* Q_xy,
                 const bpgpu_points* G, const bpgpu_points* H, const uint8_t* proof, size_t len) {
  InnerProductArgumentProof<C> p;
  int rc = ipp_from_bytes<C>(proof, len, &p);
  if (rc) return rc;
  if ((rc = points_valid<C>(P_xy, 1)) || (rc = points_valid<C>(Q_xy, 1))) return rc;
  FieldElementVector<C> dGf, dHf;
  if ((rc = FieldElementVector<C>::from_bytes(ctx, Gf, n, &dGf)) || (rc = FieldElementVector<C>::from_bytes(ctx, Hf, n, &dHf))) return rc;
  G1Vector<C> vG = G1Vector<C>::borrow(ctx, G), vH = G1Vector<C>::borrow(ctx, H);
  if (vG.len() < n || vH.len() < n) return BPGPU_E_LEN;
  Transcript t{std::string(label)};
  return IPP<C>::verify_ipp(ctx, n, t, dGf, dHf, G1<C>::from_xy(P_xy), G1<C>::from_xy(Q_xy), vG, 0, vH, 0, p.a, p.b, p.L, p.R);
}

template <class C>
int bound_prove_t(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G, const bpgpu_points* H,
                  uint64_t val, const uint8_t* randomness_be, uint64_t lower, uint64_t upper, size_t bits, int rng_mode, uint64_t seed,
                  uint8_t* proof, size_t cap, size_t* len, uint8_t* comms_xy) {
  if (val < lower || val > upper) return BPH_E_GADGET;
  Rng<C> rng = make_rng<C>(rng_mode, seed);
  if (!rng.ok()) return BPH_E_ENTROPY;
  FieldElement<C> rnd;
  if (randomness_be) rnd = FieldElement<C>::from_bytes(randomness_be);
  G1Vector<C> vG = G1Vector<C>::borrow(ctx, G), vH = G1Vector<C>::borrow(ctx, H);
  R1CSProof<C> p;
  std::vector<G1<C>> comms;
  int rc = gen_proof_of_bounded_num<C>(ctx, val, randomness_be ? &rnd : nullptr, lower, upper, bits, rng, label, G1<C>::from_xy(g_xy),
                                       G1<C>::from_xy(h_xy), vG, vH, &p, &comms);
  if (rc) return rc;
  for (size_t k = 0; k < comms.size(); k++) memcpy(comms_xy + k * 2 * C::MODBYTES, comms[k].xy, 2 * C::MODBYTES);
  return emit(p.to_bytes(), proof, cap, len);
}

template <class C>
int bound_verify_t(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G, const bpgpu_points* H,
                   uint64_t lower, uint64_t upper, size_t bits, const uint8_t* proof, size_t len, const uint8_t* comms_xy, const uint8_t* r_be) {
  R1CSProof<C> p;
  int rc = R1CSProof<C>::from_bytes(proof, len, &p);
  if (rc) return rc;
  if ((rc = points_valid<C>(comms_xy, 3)) || (rc = points_valid<C>(g_xy, 1)) || (rc = points_valid<C>(h_xy, 1))) return rc;
  std::vector<G1<C>> comms(3);
  for (int k = 0; k < 3; k++) comms[k] = G1<C>::from_xy(comms_xy + k * 2 * C::MODBYTES);
  G1Vector<C> vG = G1Vector<C>::borrow(ctx, G), vH = G1Vector<C>::borrow(ctx, H);
  FieldElement<C> vr;
  if ((rc = verifier_scalar<C>(r_be, &vr))) return rc;
  return verify_proof_of_bounded_num<C>(ctx, lower, upper, bits, p, comms, label, G1<C>::from_xy(g_xy), G1<C>::from_xy(h_xy), vG, vH, vr);
}

template <class C>
int range_circuit_csr(size_t m, size_t bits, typename Verifier<C>::CircuitCSR* csr);

// Large range statements: the circuit of (m, bits) is recorded once per context and kept on the device; each proof / check then
// builds its witness and its weights there.  Small ones run the gadget as the reference does: recording
// costs more than it saves.  BPH_RANGE_RECORDED=0 / 1 forces one path (the tests compare the two).
static bool range_use_recorded(size_t m, size_t bits) {
  const char* env = getenv("BPH_RANGE_RECORDED");
  return env ? atoi(env) != 0 : m * bits >= 1024;
}
template <class C>
int range_circuit_cached(bpgpu_ctx* ctx, size_t m, size_t bits, const bpgpu_circuit** out) {
  const uint64_t key = ((uint64_t)0x5241 << 48) | ((uint64_t)bits << 32) | (uint64_t)m;     // "RA" | bits | m
  if ((*out = bpgpu_ctx_circuit_get(ctx, key))) return BPGPU_OK;
  typename Verifier<C>::CircuitCSR csr;
  int rc = range_circuit_csr<C>(m, bits, &csr);
  if (rc) return rc;
  bpgpu_circuit* made = nullptr;
  if ((rc = bpgpu_circuit_create(ctx, csr.n, csr.m, csr.q, csr.row_start.data(), csr.ent_q.data(), csr.ent_c_be.data(), &made))) return rc;
  if ((rc = bpgpu_ctx_circuit_put(ctx, key, made))) { bpgpu_circuit_free(made); return rc; }
  *out = made;
  return BPGPU_OK;
}

template <class C>
int range_prove_t(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G, const bpgpu_points* H,
                  const uint64_t* values, size_t m, size_t bits, int rng_mode, uint64_t seed, uint8_t* proof, size_t cap, size_t* len,
                  uint8_t* comms_xy) {
  Rng<C> rng = make_rng<C>(rng_mode, seed);
  if (!rng.ok()) return BPH_E_ENTROPY;
  G1Vector<C> vG = G1Vector<C>::borrow(ctx, G), vH = G1Vector<C>::borrow(ctx, H);
  R1CSProof<C> p;
  std::vector<G1<C>> comms;
  std::vector<uint64_t> vals(values, values + m);
  int rc;
  if (range_use_recorded(m, bits)) {
    const bpgpu_circuit* circ = nullptr;
    if ((rc = range_circuit_cached<C>(ctx, m, bits, &circ))) return rc;
    rc = gen_proof_of_positive_nums_recorded<C>(ctx, circ, vals, bits, rng, label, G1<C>::from_xy(g_xy), G1<C>::from_xy(h_xy), vG, vH, &p, &comms);
  } else {
    rc = gen_proof_of_positive_nums<C>(ctx, vals, bits, rng, label, G1<C>::from_xy(g_xy), G1<C>::from_xy(h_xy), vG, vH, &p, &comms);
  }
  if (rc) return rc;
  for (size_t k = 0; k < comms.size(); k++) memcpy(comms_xy + k * 2 * C::MODBYTES, comms[k].xy, 2 * C::MODBYTES);
  return emit(p.to_bytes(), proof, cap, len);
}

template <class C>
int range_verify_t(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G, const bpgpu_points* H,
                   size_t m, size_t bits, const uint8_t* proof, size_t len, const uint8_t* comms_xy, const uint8_t* r_be) {
  R1CSProof<C> p;
  int rc = R1CSProof<C>::from_bytes(proof, len, &p);
  if (rc) return rc;
  if ((rc = points_valid<C>(comms_xy, m)) || (rc = points_valid<C>(g_xy, 1)) || (rc = points_valid<C>(h_xy, 1))) return rc;
  std::vector<G1<C>> comms(m);
  for (size_t k = 0; k < m; k++) comms[k] = G1<C>::from_xy(comms_xy + k * 2 * C::MODBYTES);
  G1Vector<C> vG = G1Vector<C>::borrow(ctx, G), vH = G1Vector<C>::borrow(ctx, H);
  FieldElement<C> vr;
  if ((rc = verifier_scalar<C>(r_be, &vr))) return rc;
  if (m > ((size_t)1 << 24) || bits > 64) return BPGPU_E_ARG;
  if (range_use_recorded(m, bits)) {
    const bpgpu_circuit* circ = nullptr;
    if ((rc = range_circuit_cached<C>(ctx, m, bits, &circ))) return rc;
    return verify_proof_of_positive_nums_recorded<C>(ctx, circ, p, comms, label, G1<C>::from_xy(g_xy), G1<C>::from_xy(h_xy), vG, vH, vr);
  }
  return verify_proof_of_positive_nums<C>(ctx, bits, p, comms, label, G1<C>::from_xy(g_xy), G1<C>::from_xy(h_xy), vG, vH, vr);
}

template <class C>
int shuffle_prove_t(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G, const bpgpu_points* H,
                    const uint64_t* x, const uint64_t* y, size_t k, size_t bits, int rng_mode, uint64_t seed, uint8_t* proof, size_t cap,
                    size_t* len, uint8_t* comms_xy) {
  Rng<C> rng = make_rng<C>(rng_mode, seed);
  if (!rng.ok()) return BPH_E_ENTROPY;
  G1Vector<C> vG = G1Vector<C>::borrow(ctx, G), vH = G1Vector<C>::borrow(ctx, H);
  R1CSProof<C> p;
  std::vector<G1<C>> comms;
  int rc = gen_proof_of_shuffle<C>(ctx, std::vector<uint64_t>(x, x + k), std::vector<uint64_t>(y, y + k), bits, rng, label, G1<C>::from_xy(g_xy),
                                   G1<C>::from_xy(h_xy), vG, vH, &p, &comms);
  if (rc) return rc;
  for (size_t j = 0; j < comms.size(); j++) memcpy(comms_xy + j * 2 * C::MODBYTES, comms[j].xy, 2 * C::MODBYTES);
  return emit(p.to_bytes(), proof, cap, len);
}

template <class C>
int shuffle_verify_t(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G, const bpgpu_points* H,
                     size_t k, size_t bits, const uint8_t* proof, size_t len, const uint8_t* comms_xy, const uint8_t* r_be) {
  R1CSProof<C> p;
  int rc = R1CSProof<C>::from_bytes(proof, len, &p);
  if (rc) return rc;
  if ((rc = points_valid<C>(comms_xy, 2 * k)) || (rc = points_valid<C>(g_xy, 1)) || (rc = points_valid<C>(h_xy, 1))) return rc;
  std::vector<G1<C>> comms(2 * k);
  for (size_t j = 0; j < 2 * k; j++) comms[j] = G1<C>::from_xy(comms_xy + j * 2 * C::MODBYTES);
  G1Vector<C> vG = G1Vector<C>::borrow(ctx, G), vH = G1Vector<C>::borrow(ctx, H);
  FieldElement<C> vr;
  if ((rc = verifier_scalar<C>(r_be, &vr))) return rc;
  return verify_proof_of_shuffle<C>(ctx, k, bits, p, comms, label, G1<C>::from_xy(g_xy), G1<C>::from_xy(h_xy), vG, vH, vr);
}

// hash_msg: SHAKE256(msg) squeezed to MODBYTES (G1::from_msg_hash / FieldElement::from_msg_hash)
void hash_msg(const uint8_t* msg, size_t len, int modbytes, uint8_t* out) { shake256(msg, len, out, modbytes); }

// run fn(i, ctx) for i in [0, count) on nctx worker threads (one context each); returns the first non-zero result
template <class F>
static int for_each_proof(bpgpu_ctx* const* ctxs, size_t nctx, size_t count, F fn) {
  std::atomic<size_t> next{0};
  std::atomic<int> err{0};
  auto worker = [&](size_t k) {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= count) break;
      int rc = fn(i, ctxs[k]);
      if (rc) { int z = 0; err.compare_exchange_strong(z, rc); }
    }
  };
  std::vector<std::thread> th;
  for (size_t k = 1; k < nctx; k++) th.emplace_back(worker, k);
  worker(0);
  for (auto& t : th) t.join();
  return err.load();
}

template <class C>
int range_verify_batch_hostscalars_t(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G, bpgpu_points* H,
                         size_t count, size_t m, size_t bits, const uint8_t* proofs, size_t stride, const uint8_t* comms_xy, size_t nthreads,
                         int32_t* verdicts) {
  const size_t mb = C::MODBYTES, pb = 2 * mb;
  const size_t n = m * bits, N = next_power_of_two(n);
  if (bpgpu_points_len(G) < N || bpgpu_points_len(H) < N) {            // InvalidGeneratorsLength for every proof
    for (size_t i = 0; i < count; i++) verdicts[i] = BPGPU_E_GENS_LEN;
    return BPGPU_OK;
  }
  int rc;
  if ((rc = bpgpu_points_precompute(ctx, G)) || (rc = bpgpu_points_precompute(ctx, H))) return rc;
  size_t lg = 0;
  while (((size_t)1 << lg) < N) lg++;
  const size_t F = 2 * N + 2, vn = 6 + m + 5 + 2 * lg;
  const size_t plen = bph_range_proof_len(C::ID, m, bits);
  std::vector<uint8_t> fixed(count * F * mb), vpts(count * vn * pb), vscal(count * vn * mb);
  const G1<C> ident = G1<C>::identity();
  if (nthreads == 0) nthreads = std::thread::hardware_concurrency();
  if (nthreads == 0) nthreads = 1;
  if (nthreads > count) nthreads = count ? count : 1;
  // Host and device overlap: the proofs are cut into slabs; worker threads fill them in index order while this thread
  // hands every completed slab to the device, so the device evaluates slab k while the workers build slab k+1...
  const size_t SLAB = count >= 4096 ? 1024 : 512;   // the Horner stage of a slab is ~3 ms of latency whatever its size
  const size_t nslab = (count + SLAB - 1) / SLAB;
  std::vector<std::atomic<size_t>> done(nslab);
  for (auto& d : done) d.store(0);
  std::atomic<size_t> next{0};
  auto worker = [&]() {
    Rng<C> os;                                      // one entropy stream per worker thread (verifier.rs:392 draws per proof)
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= count) break;
      if (!os.ok()) { verdicts[i] = BPH_E_ENTROPY; done[i / SLAB].fetch_add(1, std::memory_order_release); continue; }
      uint8_t* fo = fixed.data() + i * F * mb;
      uint8_t* po = vpts.data() + i * vn * pb;
      uint8_t* so = vscal.data() + i * vn * mb;
      R1CSProof<C> p;
      int st = R1CSProof<C>::from_bytes(proofs + i * stride, plen, &p);
      if (!st) st = points_valid<C>(comms_xy + i * m * pb, m);
      typename Verifier<C>::VerificationTerms t;
      if (!st) {
        Transcript tr{std::string(label)};
        Verifier<C> verifier(ctx, tr);
        for (size_t k = 0; k < m && !st; k++) {
          Variable var = verifier.commit(G1<C>::from_xy(comms_xy + (i * m + k) * pb));
          st = positive_no_gadget<C>(verifier, AllocatedQuantity<C>{var, false, FieldElement<C>::zero()}, bits);
        }
        if (!st) st = verifier.verification_terms_host(p, N, os.next(), &t);
        if (!st && (t.fixed.size() != F || t.var_points.size() != vn)) st = BPGPU_E_FORMAT;
      }
      verdicts[i] = st;
      if (st) {                                   // keep the device input well formed; the verdict is already decided
        memset(fo, 0, F * mb);
        memset(so, 0, vn * mb);
        for (size_t k = 0; k < vn; k++) memcpy(po + k * pb, ident.xy, pb);
      } else {
        for (size_t k = 0; k < F; k++) t.fixed[k].to_bytes(fo + k * mb);
        for (size_t k = 0; k < vn; k++) { memcpy(po + k * pb, t.var_points[k].xy, pb); t.var_scalars[k].to_bytes(so + k * mb); }
      }
      done[i / SLAB].fetch_add(1, std::memory_order_release);
    }
  };
  Trace tr("range_verify_batch");
  std::vector<std::thread> th;
  for (size_t k = 0; k < nthreads; k++) th.emplace_back(worker);
  // [G[..N] | H[..N] | g | h]
  bpgpu_fixed_run runs[4] = {{G, 0, N, nullptr}, {H, 0, N, nullptr}, {nullptr, 0, 1, g_xy}, {nullptr, 0, 1, h_xy}};
  std::vector<uint8_t> ident_flags(count + 1);
  rc = BPGPU_OK;
  for (size_t k = 0; k < nslab; k++) {
    const size_t lo = k * SLAB, cnt = count - lo < SLAB ? count - lo : SLAB;
    while (done[k].load(std::memory_order_acquire) < cnt) std::this_thread::yield();
    if (rc) continue;                             // after a device error: just let the workers drain
    rc = bpgpu_msm_batch_is_identity(ctx, runs, 4, cnt, fixed.data() + lo * F * mb, vpts.data() + lo * vn * pb, vscal.data() + lo * vn * mb, vn,
                                     ident_flags.data() + lo);
  }
  for (auto& x : th) x.join();
  if (rc) return rc;
  tr.mark("host + device, overlapped");
  for (size_t i = 0; i < count; i++)
    if (verdicts[i] == 0 && !ident_flags[i]) verdicts[i] = BPGPU_E_VERIFY;
  return BPGPU_OK;
}


// circuit of `m` values range-checked to `bits` bits each (m x positive_no_gadget), recorded once for a whole batch
template <class C>
int range_circuit_csr(size_t m, size_t bits, typename Verifier<C>::CircuitCSR* csr) {
  Transcript scratch{std::string("circuit")};
  Verifier<C> rec(nullptr, scratch);
  for (size_t k = 0; k < m; k++) {
    Variable var = rec.commit(G1<C>::identity());
    int st = positive_no_gadget<C>(rec, AllocatedQuantity<C>{var, false, FieldElement<C>::zero()}, bits);
    if (st) return st;
  }
  return rec.export_csr(csr);
}
// bound_check_gadget over three commitments (v, v - min, max - v): gadgets/bound_check.rs:94-129
template <class C>
int bound_circuit_csr(uint64_t lower, uint64_t upper, size_t bits, typename Verifier<C>::CircuitCSR* csr) {
  Transcript scratch{std::string("circuit")};
  Verifier<C> rec(nullptr, scratch);
  std::vector<G1<C>> comms(3, G1<C>::identity());
  int st = verify_bounded_num<C>(lower, upper, bits, comms, rec);
  if (st) return st;
  return rec.export_csr(csr);
}

// the challenges of one proof from a host transcript (mode 1): y, z, u, x, w, u_1..u_lg as big-endian scalars.
// Same appends as Verifier::replay_transcript + IPP::verification_challenges, on the raw proof bytes.
template <class C>
void replay_challenges_host(Transcript t, const uint8_t* proof, const uint8_t* comms_xy, size_t m, size_t lg, size_t N, uint8_t* out_be) {
  const size_t mb = C::MODBYTES, PB = 2 * mb + 1;
  uint8_t tagged[1 + 2 * 48];
  tagged[0] = 4;
  for (size_t j = 0; j < m; j++) { memcpy(tagged + 1, comms_xy + j * 2 * mb, 2 * mb); t.append_message("V", tagged, PB); }
  t.append_u64("m", m);
  auto pt = [&](const char* label, size_t k) { t.append_message(label, proof + k * PB, PB); };
  auto ch = [&](const char* label, size_t slot) {
    uint8_t buf[48];
    t.challenge_bytes(label, buf, mb);
    FieldElement<C>::from_bytes(buf).to_bytes(out_be + slot * mb);
  };
  pt("A_I1", 0); pt("A_O1", 1); pt("S1", 2);
  t.r1cs_1phase_domain_sep();
  pt("A_I2", 3); pt("A_O2", 4); pt("S2", 5);
  ch("y", 0); ch("z", 1);
  pt("T_1", 6); pt("T_3", 7); pt("T_4", 8); pt("T_5", 9); pt("T_6", 10);
  ch("u", 2); ch("x", 3);
  const uint8_t* sc = proof + 11 * PB;
  t.append_message("t_x", sc, mb);
  t.append_message("t_x_blinding", sc + mb, mb);
  t.append_message("e_blinding", sc + 2 * mb, mb);
  ch("w", 4);
  t.innerproduct_domain_sep(N);
  const uint8_t* L = sc + 3 * mb;
  for (size_t k = 0; k < lg; k++) {
    t.append_message("L", L + k * PB, PB);
    t.append_message("R", L + (lg + k) * PB, PB);
    ch("u", 5 + k);
  }
}

// Small statements proved / verified in bulk get WIDE generator tables (16-bit windows, bpgpu_points_precompute_wide: 100 MB
// per BLS12-381 generator): N <= 64 by default (2 x 6.4 GB), BPH_WIDE_TABLES=<max N> moves the limit, 0 turns it off.
static size_t wide_tables_max_n() {
  const char* e = getenv("BPH_WIDE_TABLES");
  return e ? (size_t)atol(e) : 64;
}
static int maybe_wide_tables(bpgpu_ctx* ctx, bpgpu_points* G, bpgpu_points* H, size_t N, size_t count) {
  if (N > wide_tables_max_n() || count < 256 || bpgpu_points_len(G) > 2 * N || bpgpu_points_len(H) > 2 * N) return BPGPU_OK;
  int rc = bpgpu_points_precompute_wide(ctx, G);
  return rc ? rc : bpgpu_points_precompute_wide(ctx, H);
}

// mode 0: transcripts replayed on the device; mode 1: transcripts on `nthreads` host threads, challenges uploaded
template <class C>
int verify_batch_device_t(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G, bpgpu_points* H,
                          const typename Verifier<C>::CircuitCSR& csr, size_t count, const uint8_t* proofs, size_t stride, const uint8_t* comms_xy,
                          int mode, size_t nthreads, int32_t* verdicts) {
  const size_t mb = C::MODBYTES;
  const size_t m = csr.m, N = next_power_of_two(csr.n);
  size_t lg = 0;
  while (((size_t)1 << lg) < N) lg++;
  Trace tr("verify_batch_device");
  bpgpu_circuit* circ = nullptr;
  int rc = bpgpu_circuit_create(ctx, csr.n, csr.m, csr.q, csr.row_start.data(), csr.ent_q.data(), csr.ent_c_be.data(), &circ);
  if (rc) return rc;
  Transcript t0{std::string(label)};
  t0.r1cs_domain_sep();                                   // Verifier::new (verifier.rs:97-108)
  Rng<C> os;                                              // one key per call; proof i takes draw i (verifier.rs:392)
  if (!os.ok()) { bpgpu_circuit_free(circ); return BPH_E_ENTROPY; }
  if (bpgpu_points_len(G) >= N && bpgpu_points_len(H) >= N &&          // tables built once, before the drivers share them
      ((rc = bpgpu_points_precompute(ctx, G)) || (rc = bpgpu_points_precompute(ctx, H)) || (rc = maybe_wide_tables(ctx, G, H, N, count)))) {
    bpgpu_circuit_free(circ);
    return rc;
  }
  // Two slabs in flight on two contexts (the second one cached in the first, bpgpu_ctx_aux): the latency-bound kernels of a
  // slab (per-proof transcripts, the 252-doubling Horner chains) run under the table sums of the other.  Every slab call
  // gets its own key for the verifiers' random scalars (key || slab index).
  const size_t SLABV = getenv("BPGPU_VB_SLAB") && atol(getenv("BPGPU_VB_SLAB")) > 0 ? (size_t)atol(getenv("BPGPU_VB_SLAB")) : 4096;
  const size_t nslab = (count + SLABV - 1) / SLABV;
  size_t maxdrv = nslab >= 8 ? 4 : 2;
  if (const char* e = getenv("BPH_VB_DRIVERS")) maxdrv = (size_t)atol(e) >= 1 && (size_t)atol(e) <= 4 ? (size_t)atol(e) : maxdrv;   // tuning runs
  const size_t ndrv = nslab < maxdrv ? nslab : maxdrv;
  bpgpu_ctx* dctxs[4] = {ctx, nullptr, nullptr, nullptr};
  for (size_t k = 1; k < ndrv; k++)
    if ((rc = bpgpu_ctx_aux(dctxs[k - 1], &dctxs[k]))) { bpgpu_circuit_free(circ); return rc; }
  uint8_t state0[203];
  t0.export_state(state0);
  std::vector<uint8_t> chal;
  const size_t nch = 5 + lg;
  if (mode != 0) {
    chal.resize(count * nch * mb);
    if (nthreads == 0) nthreads = std::thread::hardware_concurrency();
    if (nthreads == 0) nthreads = 1;
    if (nthreads > count) nthreads = count;
    std::atomic<size_t> next{0};
    auto worker = [&]() {
      for (;;) {
        const size_t lo = next.fetch_add(64);
        if (lo >= count) break;
        const size_t hi = lo + 64 < count ? lo + 64 : count;
        for (size_t i = lo; i < hi; i++)
          replay_challenges_host<C>(t0, proofs + i * stride, comms_xy + i * m * 2 * mb, m, lg, N, chal.data() + i * nch * mb);
      }
    };
    std::vector<std::thread> th;
    for (size_t k = 1; k < nthreads; k++) th.emplace_back(worker);
    worker();
    for (auto& x : th) x.join();
    tr.mark("host transcripts");
  }
  std::atomic<int> err{0};
  auto driver = [&](size_t k) {
    for (size_t sl = k; sl < nslab && !err.load(); sl += ndrv) {
      const size_t lo = sl * SLABV, cnt = count - lo < SLABV ? count - lo : SLABV;
      uint8_t key[56];
      memcpy(key, os.key(), os.key_len());
      key[os.key_len()] = (uint8_t)sl;
      key[os.key_len() + 1] = (uint8_t)(sl >> 8);
      const int r = bpgpu_r1cs_verify_batch(dctxs[k], circ, G, H, g_xy, h_xy, cnt, proofs + lo * stride, stride, comms_xy + lo * m * 2 * mb,
                                            mode == 0 ? state0 : nullptr, mode == 0 ? nullptr : chal.data() + lo * nch * mb, key,
                                            os.key_len() + 2, verdicts + lo);
      secure_zero(key, sizeof key);
      if (r) { int z = 0; err.compare_exchange_strong(z, r); }
    }
  };
  {
    std::vector<std::thread> others;
    for (size_t k = 1; k < ndrv; k++) others.emplace_back(driver, k);
    driver(0);
    for (auto& t : others) t.join();
  }
  rc = err.load();
  tr.mark("device");
  bpgpu_circuit_free(circ);
  return rc;
}

template <class C>
int range_verify_batch_t(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G, bpgpu_points* H,
                         size_t count, size_t m, size_t bits, const uint8_t* proofs, size_t stride, const uint8_t* comms_xy, int mode,
                         size_t nthreads, int32_t* verdicts) {
  if (mode == 2) return range_verify_batch_hostscalars_t<C>(ctx, label, g_xy, h_xy, G, H, count, m, bits, proofs, stride, comms_xy, nthreads, verdicts);
  typename Verifier<C>::CircuitCSR csr;
  int rc = range_circuit_csr<C>(m, bits, &csr);
  if (rc) return rc;
  return verify_batch_device_t<C>(ctx, label, g_xy, h_xy, G, H, csr, count, proofs, stride, comms_xy, mode, nthreads, verdicts);
}

template <class C>
int bound_verify_batch_t(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G, bpgpu_points* H,
                         size_t count, uint64_t lower, uint64_t upper, size_t bits, const uint8_t* proofs, size_t stride, const uint8_t* comms_xy,
                         int mode, size_t nthreads, int32_t* verdicts) {
  typename Verifier<C>::CircuitCSR csr;
  int rc = bound_circuit_csr<C>(lower, upper, bits, &csr);
  if (rc) return rc;
  return verify_batch_device_t<C>(ctx, label, g_xy, h_xy, G, H, csr, count, proofs, stride, comms_xy, mode ? 1 : 0, nthreads, verdicts);
}

template <class C>
int circuit_csr_export(const typename Verifier<C>::CircuitCSR& csr, size_t* n, size_t* m, size_t* q, size_t* nnz, uint32_t* row_start,
                       uint32_t* ent_q, uint8_t* ent_c_be) {
  *n = csr.n; *m = csr.m; *q = csr.q; *nnz = csr.ent_q.size();
  if (row_start) memcpy(row_start, csr.row_start.data(), csr.row_start.size() * 4);
  if (ent_q) memcpy(ent_q, csr.ent_q.data(), csr.ent_q.size() * 4);
  if (ent_c_be) memcpy(ent_c_be, csr.ent_c_be.data(), csr.ent_q.size() * C::MODBYTES);
  return BPGPU_OK;
}

// mode 0: the whole slab on the device (bpgpu_pbatch_prove_range: transcripts, draws, flattening and normalisation there too);
// mode 1: lock-step with the transcripts on `nthreads` host threads (prove_batch.hpp)
template <class C>
int range_prove_batch_device_t(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G, bpgpu_points* H,
                               const uint64_t* values, size_t count, size_t m, size_t bits, int rng_mode, uint64_t seed, uint8_t* proofs,
                               size_t stride, uint8_t* comms_xy) {
  const size_t mb = C::MODBYTES, n = m * bits;
  typename Verifier<C>::CircuitCSR csr;
  int rc = range_circuit_csr<C>(m, bits, &csr);
  if (rc) return rc;
  if ((rc = points_valid<C>(g_xy, 1)) || (rc = points_valid<C>(h_xy, 1))) return rc;
  if ((rc = bpgpu_points_precompute(ctx, G)) || (rc = bpgpu_points_precompute(ctx, H)) ||
      (rc = maybe_wide_tables(ctx, G, H, next_power_of_two(n), count)))
    return rc;
  uint8_t state0[203];
  {
    Transcript t0{std::string(label)};
    t0.r1cs_domain_sep();                                  // Prover::new (prover.rs:84-101)
    t0.export_state(state0);
  }
  // one blinding stream per proof (host/curve.hpp Rng): seed + i || "blind", or 32 bytes of OS entropy || "os"
  const size_t klen = rng_mode == 1 ? 13 : 34;
  std::vector<uint8_t> keys(count * klen);
  for (size_t i = 0; i < count; i++) {
    Rng<C> r = rng_mode == 1 ? Rng<C>(seed + i, "blind") : Rng<C>();
    if (!r.ok() || r.key_len() != klen) return BPH_E_ENTROPY;
    memcpy(keys.data() + i * klen, r.key(), klen);
  }
  // Slabs in flight: the per-proof kernels of a slab (transcripts, inversions: a few warps, latency bound) run in the gaps of
  // the table sums of the others.  BPH_PB_SLAB / BPH_PB_DRIVERS override the defaults (tuning runs).
  size_t SLAB = n <= 128 ? 2048 : (n <= 1024 ? 1024 : 256);
  size_t maxdrv = 2;
  if (const char* e = getenv("BPH_PB_SLAB")) SLAB = (size_t)atol(e) ? (size_t)atol(e) : SLAB;
  if (const char* e = getenv("BPH_PB_DRIVERS")) maxdrv = (size_t)atol(e) >= 1 && (size_t)atol(e) <= 4 ? (size_t)atol(e) : maxdrv;
  size_t ndrv = (count + SLAB - 1) / SLAB;
  if (ndrv > maxdrv) ndrv = maxdrv;
  if (ndrv == 0) ndrv = 1;
  size_t B = (count + ndrv - 1) / ndrv;
  if (B > SLAB) B = SLAB;
  const size_t nslab = (count + B - 1) / B;
  bpgpu_ctx* dctxs[4] = {ctx, nullptr, nullptr, nullptr};
  for (size_t k = 1; k < ndrv; k++)
    if ((rc = bpgpu_ctx_aux(dctxs[k - 1], &dctxs[k]))) return rc;      // a chain of cached contexts: ctx -> aux -> aux of aux
  std::atomic<int> err{0};
  auto driver = [&](size_t k, bpgpu_ctx* dctx) {
    bpgpu_pbatch* pb = nullptr;
    bpgpu_circuit* circ = nullptr;
    size_t pb_size = 0;
    int r = bpgpu_circuit_create(dctx, csr.n, csr.m, csr.q, csr.row_start.data(), csr.ent_q.data(), csr.ent_c_be.data(), &circ);
    for (size_t sl = k; !r && sl < nslab && !err.load(); sl += ndrv) {
      const size_t lo = sl * B, cnt = count - lo < B ? count - lo : B;
      if (cnt != pb_size) {
        bpgpu_pbatch_free(pb);
        pb = nullptr;
        r = bpgpu_pbatch_create(dctx, G, H, g_xy, h_xy, cnt, n, &pb);
        pb_size = cnt;
      }
      if (!r)
        r = bpgpu_pbatch_prove_range(pb, circ, values + lo * m, m, bits, state0, keys.data() + lo * klen, klen, proofs + lo * stride, stride,
                                     comms_xy + lo * m * 2 * mb);
    }
    if (r) { int z = 0; err.compare_exchange_strong(z, r); }
    bpgpu_pbatch_free(pb);
    bpgpu_circuit_free(circ);
  };
  {
    std::vector<std::thread> others;
    for (size_t k = 1; k < ndrv; k++) others.emplace_back(driver, k, dctxs[k]);
    driver(0, ctx);
    for (auto& t : others) t.join();
  }
  secure_zero(keys.data(), keys.size());
  return err.load();
}

template <class C>
int range_prove_batch_t(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G, bpgpu_points* H,
                        const uint64_t* values, size_t count, size_t m, size_t bits, int rng_mode, uint64_t seed, int mode, size_t nthreads,
                        uint8_t* proofs, size_t stride, uint8_t* comms_xy) {
  if (mode == 0) return range_prove_batch_device_t<C>(ctx, label, g_xy, h_xy, G, H, values, count, m, bits, rng_mode, seed, proofs, stride, comms_xy);
  if (nthreads == 0) nthreads = std::thread::hardware_concurrency();
  if (nthreads == 0) nthreads = 1;
  // Two drivers (this thread on `ctx`, a second thread on a context of its own on the same device) take alternate slabs:
  // while one slab's transcripts run on the host, the other slab's stage runs on the device.
  const size_t ndrv = (count >= 4096 && nthreads >= 2) ? 2 : 1;      // a second context costs ~20 ms to set up
  const size_t SLAB = 1024;
  size_t B = (count + ndrv - 1) / ndrv;
  if (B > SLAB) B = SLAB;
  const size_t nslab = (count + B - 1) / B;
  const G1<C> g = G1<C>::from_xy(g_xy), h = G1<C>::from_xy(h_xy);
  bpgpu_ctx* ctx2 = nullptr;
  int rc = BPGPU_OK;
  if (ndrv == 2 && (rc = bpgpu_ctx_aux(ctx, &ctx2))) return rc;
  if ((rc = bpgpu_points_precompute(ctx, G)) || (rc = bpgpu_points_precompute(ctx, H))) return rc;
  std::atomic<int> err{0};
  auto driver = [&](size_t k, bpgpu_ctx* dctx) {
    typename BatchProverAccess<C>::Pool pool(std::max<size_t>(1, nthreads / ndrv));
    bpgpu_pbatch* pb = nullptr;
    size_t pb_size = 0;
    for (size_t sl = k; sl < nslab && !err.load(); sl += ndrv) {
      const size_t lo = sl * B, cnt = count - lo < B ? count - lo : B;
      int r = BPGPU_OK;
      if (cnt != pb_size) {                           // the scratch layout depends on the slab size
        bpgpu_pbatch_free(pb);
        pb = nullptr;
        r = bpgpu_pbatch_create(dctx, G, H, g_xy, h_xy, cnt, m * bits, &pb);
        pb_size = cnt;
      }
      if (!r)
        r = BatchProverAccess<C>::prove_slab(dctx, pb, label, g, h, values + lo * m, cnt, m, bits, rng_mode, seed + lo, pool, proofs + lo * stride,
                                             stride, comms_xy + lo * m * 2 * C::MODBYTES);
      if (r) { int z = 0; err.compare_exchange_strong(z, r); }
    }
    bpgpu_pbatch_free(pb);
  };
  if (ndrv == 2) {
    std::thread other(driver, 1, ctx2);
    driver(0, ctx);
    other.join();
  } else {
    driver(0, ctx);
  }
  return err.load();
}

template <class FqP>
void g1_sum_host(const uint8_t* xy, size_t count, int mb, uint8_t* out) {
  using HP = bp::host::HXYZZ<FqP>;
  using F = typename HP::F;
  HP acc = HP::inf();
  for (size_t i = 0; i < count; i++) {
    const uint8_t* p = xy + i * 2 * mb;
    bool ident = p[2 * mb - 1] == 1;                                    // AMCL's identity encoding: x = 0, y = 1
    for (int k = 0; k < 2 * mb - 1 && ident; k++) ident = p[k] == 0;
    if (ident) continue;
    HP q;
    q.x = F::from_be(p, mb);
    q.y = F::from_be(p + mb, mb);
    q.zz = F::one();
    q.zzz = F::one();
    acc.add(q);
  }
  acc.to_xy_be(mb, out);
}

}  // namespace

#define BY_CURVE(ctx, CALL) (bpgpu_ctx_curve(ctx) == BPGPU_BLS12_381 ? CALL(Bls381) : CALL(Bn254))

extern "C" {

int bph_merlin_kat(const char* label, const char* msg_label, const uint8_t* msg, size_t msg_len, const char* ch_label, uint8_t* out,
                   size_t out_len) {
  if (!label || !msg_label || !ch_label || !out) return BPGPU_E_ARG;
  Transcript t{std::string(label)};
  t.append_message(msg_label, msg, msg_len);
  t.challenge_bytes(ch_label, out, out_len);
  return BPGPU_OK;
}

int bph_transcript_kat(int curve, const char* label, const uint8_t* point_xy, const uint8_t* scalar_be, uint8_t* challenge_be) {
  if (!label || !point_xy || !scalar_be || !challenge_be) return BPGPU_E_ARG;
  Transcript t{std::string(label)};
  if (curve == BPGPU_BLS12_381) {
    using C = Bls381;
    t.innerproduct_domain_sep(8);
    TranscriptProtocol<C>::commit_point(t, "P", G1<C>::from_xy(point_xy));
    TranscriptProtocol<C>::commit_scalar(t, "s", FieldElement<C>::from_bytes(scalar_be));
    TranscriptProtocol<C>::challenge_scalar(t, "c").to_bytes(challenge_be);
  } else if (curve == BPGPU_BN254) {
    using C = Bn254;
    t.innerproduct_domain_sep(8);
    TranscriptProtocol<C>::commit_point(t, "P", G1<C>::from_xy(point_xy));
    TranscriptProtocol<C>::commit_scalar(t, "s", FieldElement<C>::from_bytes(scalar_be));
    TranscriptProtocol<C>::challenge_scalar(t, "c").to_bytes(challenge_be);
  } else {
    return BPGPU_E_ARG;
  }
  return BPGPU_OK;
}

int bph_get_generators(bpgpu_ctx* ctx, const char* prefix, size_t n, bpgpu_points** out) {
  if (!ctx || !prefix || !out) return BPGPU_E_ARG;
  const int mb = bpgpu_modbytes(bpgpu_ctx_curve(ctx));
  std::vector<uint8_t> hashes(n * mb + 1);
  for (size_t i = 1; i <= n; i++) {                       // utils/mod.rs:18-21: prefix || i.to_string()
    std::string s = std::string(prefix) + std::to_string(i);
    hash_msg((const uint8_t*)s.data(), s.size(), mb, hashes.data() + (i - 1) * mb);
  }
  return bpgpu_points_from_hashes(ctx, hashes.data(), n, out);
}

int bph_g1_from_msg_hash(bpgpu_ctx* ctx, const uint8_t* msg, size_t msg_len, uint8_t* out_xy) {
  if (!ctx || (!msg && msg_len) || !out_xy) return BPGPU_E_ARG;
  const int mb = bpgpu_modbytes(bpgpu_ctx_curve(ctx));
  uint8_t h[48];
  hash_msg(msg, msg_len, mb, h);
  bpgpu_points* p = nullptr;
  int rc = bpgpu_points_from_hashes(ctx, h, 1, &p);
  if (rc) return rc;
  rc = bpgpu_points_download(ctx, p, 0, 1, out_xy);
  bpgpu_points_free(p);
  return rc;
}

int bph_ipp_create(bpgpu_ctx* ctx, const char* label, const bpgpu_points* G, const bpgpu_points* H, const uint8_t* Q_xy, const uint8_t* Gf,
                   const uint8_t* Hf, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* proof, size_t cap, size_t* len) {
  if (!ctx || !label || !G || !H || !Q_xy || !Gf || !Hf || !a || !b) return BPGPU_E_ARG;
#define CALL(C) ipp_create_t<C>(ctx, label, G, H, Q_xy, Gf, Hf, a, b, n, proof, cap, len)
  return BY_CURVE(ctx, CALL);
#undef CALL
}

int bph_ipp_verify(bpgpu_ctx* ctx, const char* label, size_t n, const uint8_t* Gf, const uint8_t* Hf, const uint8_t* P_xy, const uint8_t* Q_xy,
                   const bpgpu_points* G, const bpgpu_points* H, const uint8_t* proof, size_t len) {
  if (!ctx || !label || !G || !H || !P_xy || !Q_xy || !Gf || !Hf || !proof) return BPGPU_E_ARG;
#define CALL(C) ipp_verify_t<C>(ctx, label, n, Gf, Hf, P_xy, Q_xy, G, H, proof, len)
  return BY_CURVE(ctx, CALL);
#undef CALL
}

int bph_bound_check_prove(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G,
                          const bpgpu_points* H, uint64_t val, const uint8_t* randomness_be, uint64_t lower, uint64_t upper, size_t bits,
                          int rng_mode, uint64_t seed, uint8_t* proof, size_t cap, size_t* len, uint8_t* comms_xy) {
  if (!ctx || !label || !g_xy || !h_xy || !G || !H || !comms_xy) return BPGPU_E_ARG;
#define CALL(C) bound_prove_t<C>(ctx, label, g_xy, h_xy, G, H, val, randomness_be, lower, upper, bits, rng_mode, seed, proof, cap, len, comms_xy)
  return BY_CURVE(ctx, CALL);
#undef CALL
}

int bph_bound_check_verify(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G,
                           const bpgpu_points* H, uint64_t lower, uint64_t upper, size_t bits, const uint8_t* proof, size_t len,
                           const uint8_t* comms_xy, const uint8_t* r_be) {
  if (!ctx || !label || !g_xy || !h_xy || !G || !H || !proof || !comms_xy) return BPGPU_E_ARG;
#define CALL(C) bound_verify_t<C>(ctx, label, g_xy, h_xy, G, H, lower, upper, bits, proof, len, comms_xy, r_be)
  return BY_CURVE(ctx, CALL);
#undef CALL
}

int bph_range_prove(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G, const bpgpu_points* H,
                    const uint64_t* values, size_t m, size_t bits, int rng_mode, uint64_t seed, uint8_t* proof, size_t cap, size_t* len,
                    uint8_t* comms_xy) {
  if (!ctx || !label || !g_xy || !h_xy || !G || !H || (!values && m) || !comms_xy || bits > 64) return BPGPU_E_ARG;
#define CALL(C) range_prove_t<C>(ctx, label, g_xy, h_xy, G, H, values, m, bits, rng_mode, seed, proof, cap, len, comms_xy)
  return BY_CURVE(ctx, CALL);
#undef CALL
}

int bph_range_verify(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G, const bpgpu_points* H,
                     size_t m, size_t bits, const uint8_t* proof, size_t len, const uint8_t* comms_xy, const uint8_t* r_be) {
  if (!ctx || !label || !g_xy || !h_xy || !G || !H || !proof || (!comms_xy && m)) return BPGPU_E_ARG;
#define CALL(C) range_verify_t<C>(ctx, label, g_xy, h_xy, G, H, m, bits, proof, len, comms_xy, r_be)
  return BY_CURVE(ctx, CALL);
#undef CALL
}

int bph_shuffle_prove(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G, const bpgpu_points* H,
                      const uint64_t* x, const uint64_t* y, size_t k, size_t bits, int rng_mode, uint64_t seed, uint8_t* proof, size_t cap,
                      size_t* len, uint8_t* comms_xy) {
  if (!ctx || !label || !g_xy || !h_xy || !G || !H || !x || !y || !k || !comms_xy || bits > 64) return BPGPU_E_ARG;
#define CALL(C) shuffle_prove_t<C>(ctx, label, g_xy, h_xy, G, H, x, y, k, bits, rng_mode, seed, proof, cap, len, comms_xy)
  return BY_CURVE(ctx, CALL);
#undef CALL
}

int bph_shuffle_verify(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G, const bpgpu_points* H,
                       size_t k, size_t bits, const uint8_t* proof, size_t len, const uint8_t* comms_xy, const uint8_t* r_be) {
  if (!ctx || !label || !g_xy || !h_xy || !G || !H || !proof || !k || !comms_xy) return BPGPU_E_ARG;
#define CALL(C) shuffle_verify_t<C>(ctx, label, g_xy, h_xy, G, H, k, bits, proof, len, comms_xy, r_be)
  return BY_CURVE(ctx, CALL);
#undef CALL
}

size_t bph_range_proof_len(int curve, size_t m, size_t bits) {
  const size_t mb = curve == BPGPU_BLS12_381 ? 48 : 32;
  const size_t N = next_power_of_two(m * bits);
  size_t lg = 0;
  while (((size_t)1 << lg) < N) lg++;
  return 11 * (2 * mb + 1) + 3 * mb + 2 * lg * (2 * mb + 1) + 2 * mb;
}

int bph_range_prove_many(bpgpu_ctx* const* ctxs, size_t nctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G,
                         const bpgpu_points* H, const uint64_t* values, size_t count, size_t m, size_t bits, int rng_mode, uint64_t seed,
                         uint8_t* proofs, size_t proof_stride, uint8_t* comms_xy) {
  if (!ctxs || !nctx || !label || !g_xy || !h_xy || !G || !H || !values || !proofs || !comms_xy) return BPGPU_E_ARG;
  const int curve = bpgpu_ctx_curve(ctxs[0]);
  if (proof_stride < bph_range_proof_len(curve, m, bits)) return BPH_E_BUFFER;
  const size_t cstride = m * 2 * bpgpu_modbytes(curve);
  return for_each_proof(ctxs, nctx, count, [&](size_t i, bpgpu_ctx* ctx) {
    size_t len = 0;
    return bph_range_prove(ctx, label, g_xy, h_xy, G, H, values + i * m, m, bits, rng_mode, seed + i, proofs + i * proof_stride, proof_stride,
                           &len, comms_xy + i * cstride);
  });
}

int bph_range_verify_many(bpgpu_ctx* const* ctxs, size_t nctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G,
                          const bpgpu_points* H, size_t count, size_t m, size_t bits, const uint8_t* proofs, size_t proof_stride,
                          const uint8_t* comms_xy, int32_t* verdicts) {
  if (!ctxs || !nctx || !label || !g_xy || !h_xy || !G || !H || !proofs || !comms_xy || !verdicts) return BPGPU_E_ARG;
  const int curve = bpgpu_ctx_curve(ctxs[0]);
  const size_t plen = bph_range_proof_len(curve, m, bits);
  if (proof_stride < plen) return BPH_E_BUFFER;
  const size_t cstride = m * 2 * bpgpu_modbytes(curve);
  for_each_proof(ctxs, nctx, count, [&](size_t i, bpgpu_ctx* ctx) {
    verdicts[i] = bph_range_verify(ctx, label, g_xy, h_xy, G, H, m, bits, proofs + i * proof_stride, plen, comms_xy + i * cstride, nullptr);
    return 0;
  });
  return BPGPU_OK;
}

int bph_range_prove_batch_mode(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G, bpgpu_points* H,
                               const uint64_t* values, size_t count, size_t m, size_t bits, int rng_mode, uint64_t seed, int mode, size_t nthreads,
                               uint8_t* proofs, size_t proof_stride, uint8_t* comms_xy) {
  if (!ctx || !label || !g_xy || !h_xy || !G || !H || (count && (!values || !proofs || !comms_xy)) || !m || !bits || bits > 64 || mode < 0 || mode > 1)
    return BPGPU_E_ARG;
  if (proof_stride < bph_range_proof_len(bpgpu_ctx_curve(ctx), m, bits)) return BPH_E_BUFFER;
  if (count == 0) return BPGPU_OK;
#define CALL(C) range_prove_batch_t<C>(ctx, label, g_xy, h_xy, G, H, values, count, m, bits, rng_mode, seed, mode, nthreads, proofs, proof_stride, comms_xy)
  return BY_CURVE(ctx, CALL);
#undef CALL
}

int bph_range_prove_batch(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G, bpgpu_points* H,
                          const uint64_t* values, size_t count, size_t m, size_t bits, int rng_mode, uint64_t seed, size_t nthreads,
                          uint8_t* proofs, size_t proof_stride, uint8_t* comms_xy) {
  return bph_range_prove_batch_mode(ctx, label, g_xy, h_xy, G, H, values, count, m, bits, rng_mode, seed, 0, nthreads, proofs, proof_stride, comms_xy);
}

int bph_g1_sum(int curve, const uint8_t* points_xy, size_t count, uint8_t* out_xy) {
  if ((!points_xy && count) || !out_xy || (curve != BPGPU_BLS12_381 && curve != BPGPU_BN254)) return BPGPU_E_ARG;
  if (curve == BPGPU_BLS12_381) g1_sum_host<bp::BlsFq>(points_xy, count, 48, out_xy);
  else g1_sum_host<bp::BnFq>(points_xy, count, 32, out_xy);
  return BPGPU_OK;
}

int bph_msm_sharded(bpgpu_ctx* const* ctxs, size_t nctx, const bpgpu_points* const* shards, const size_t* counts, const uint8_t* scalars_be,
                    uint8_t* out_xy) {
  if (!ctxs || !nctx || !shards || !counts || !out_xy) return BPGPU_E_ARG;
  const int curve = bpgpu_ctx_curve(ctxs[0]);
  const size_t mb = (size_t)bpgpu_modbytes(curve);
  std::vector<size_t> off(nctx + 1, 0);
  for (size_t k = 0; k < nctx; k++) {
    if (!ctxs[k] || bpgpu_ctx_curve(ctxs[k]) != curve || (counts[k] && !shards[k])) return BPGPU_E_ARG;
    if (counts[k] && bpgpu_points_len(shards[k]) < counts[k]) return BPGPU_E_LEN;
    off[k + 1] = off[k] + counts[k];
  }
  if (off[nctx] && !scalars_be) return BPGPU_E_ARG;
  std::vector<uint8_t> partial(nctx * 2 * mb);
  std::vector<int> rcs(nctx, BPGPU_OK);
  auto work = [&](size_t k) {
    static const uint8_t none = 0;
    rcs[k] = bpgpu_msm(ctxs[k], shards[k], 0, counts[k], counts[k] ? scalars_be + off[k] * mb : &none, partial.data() + k * 2 * mb);
  };
  std::vector<std::thread> th;
  for (size_t k = 1; k < nctx; k++) th.emplace_back(work, k);
  work(0);
  for (auto& t : th) t.join();
  for (int rc : rcs) if (rc) return rc;
  return bph_g1_sum(curve, partial.data(), nctx, out_xy);
}

int bph_range_verify_batch_mode(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G, bpgpu_points* H,
                                size_t count, size_t m, size_t bits, const uint8_t* proofs, size_t proof_stride, const uint8_t* comms_xy,
                                int mode, size_t nthreads, int32_t* verdicts) {
  if (!ctx || !label || !g_xy || !h_xy || !G || !H || (count && (!proofs || !comms_xy || !verdicts)) || mode < 0 || mode > 2) return BPGPU_E_ARG;
  if (!m || !bits || bits > 64) return BPGPU_E_ARG;
  if (proof_stride < bph_range_proof_len(bpgpu_ctx_curve(ctx), m, bits)) return BPH_E_BUFFER;
  if (count == 0) return BPGPU_OK;
#define CALL(C) range_verify_batch_t<C>(ctx, label, g_xy, h_xy, G, H, count, m, bits, proofs, proof_stride, comms_xy, mode, nthreads, verdicts)
  return BY_CURVE(ctx, CALL);
#undef CALL
}

int bph_range_verify_batch(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G, bpgpu_points* H,
                           size_t count, size_t m, size_t bits, const uint8_t* proofs, size_t proof_stride, const uint8_t* comms_xy,
                           size_t nthreads, int32_t* verdicts) {
  return bph_range_verify_batch_mode(ctx, label, g_xy, h_xy, G, H, count, m, bits, proofs, proof_stride, comms_xy, 0, nthreads, verdicts);
}

int bph_bound_check_verify_batch(bpgpu_ctx* ctx, const char* label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G, bpgpu_points* H,
                                 size_t count, uint64_t lower, uint64_t upper, size_t bits, const uint8_t* proofs, size_t proof_stride,
                                 const uint8_t* comms_xy, int mode, size_t nthreads, int32_t* verdicts) {
  if (!ctx || !label || !g_xy || !h_xy || !G || !H || (count && (!proofs || !comms_xy || !verdicts)) || mode < 0 || mode > 1) return BPGPU_E_ARG;
  if (!bits || bits > 64 || lower > upper) return BPGPU_E_ARG;
  if (proof_stride < bph_range_proof_len(bpgpu_ctx_curve(ctx), 2, bits)) return BPH_E_BUFFER;
  if (count == 0) return BPGPU_OK;
#define CALL(C) bound_verify_batch_t<C>(ctx, label, g_xy, h_xy, G, H, count, lower, upper, bits, proofs, proof_stride, comms_xy, mode, nthreads, verdicts)
  return BY_CURVE(ctx, CALL);
#undef CALL
}

int bph_range_circuit_csr(int curve, size_t m, size_t bits, size_t* n, size_t* m_out, size_t* q, size_t* nnz, uint32_t* row_start, uint32_t* ent_q,
                          uint8_t* ent_coeff_be) {
  if (!n || !m_out || !q || !nnz || !m || !bits || bits > 64 || (curve != BPGPU_BLS12_381 && curve != BPGPU_BN254)) return BPGPU_E_ARG;
  int rc;
  if (curve == BPGPU_BLS12_381) {
    Verifier<Bls381>::CircuitCSR csr;
    if ((rc = range_circuit_csr<Bls381>(m, bits, &csr))) return rc;
    return circuit_csr_export<Bls381>(csr, n, m_out, q, nnz, row_start, ent_q, ent_coeff_be);
  }
  Verifier<Bn254>::CircuitCSR csr;
  if ((rc = range_circuit_csr<Bn254>(m, bits, &csr))) return rc;
  return circuit_csr_export<Bn254>(csr, n, m_out, q, nnz, row_start, ent_q, ent_coeff_be);
}

int bph_bound_check_circuit_csr(int curve, uint64_t lower, uint64_t upper, size_t bits, size_t* n, size_t* m_out, size_t* q, size_t* nnz,
                                uint32_t* row_start, uint32_t* ent_q, uint8_t* ent_coeff_be) {
  if (!n || !m_out || !q || !nnz || !bits || bits > 64 || lower > upper || (curve != BPGPU_BLS12_381 && curve != BPGPU_BN254)) return BPGPU_E_ARG;
  int rc;
  if (curve == BPGPU_BLS12_381) {
    Verifier<Bls381>::CircuitCSR csr;
    if ((rc = bound_circuit_csr<Bls381>(lower, upper, bits, &csr))) return rc;
    return circuit_csr_export<Bls381>(csr, n, m_out, q, nnz, row_start, ent_q, ent_coeff_be);
  }
  Verifier<Bn254>::CircuitCSR csr;
  if ((rc = bound_circuit_csr<Bn254>(lower, upper, bits, &csr))) return rc;
  return circuit_csr_export<Bn254>(csr, n, m_out, q, nnz, row_start, ent_q, ent_coeff_be);
}

int bph_r1cs_replay_challenges(int curve, const char* label, const uint8_t* proof, const uint8_t* comms_xy, size_t m, size_t lg, uint8_t* out_be) {
  if (!label || !proof || (!comms_xy && m) || !out_be || lg >= 32 || (curve != BPGPU_BLS12_381 && curve != BPGPU_BN254)) return BPGPU_E_ARG;
  Transcript t{std::string(label)};
  t.r1cs_domain_sep();
  if (curve == BPGPU_BLS12_381) replay_challenges_host<Bls381>(t, proof, comms_xy, m, lg, (size_t)1 << lg, out_be);
  else replay_challenges_host<Bn254>(t, proof, comms_xy, m, lg, (size_t)1 << lg, out_be);
  return BPGPU_OK;
}

void bph_set_secret_fixed_schedule(int on) { secret_fixed_schedule_flag().store(on ? 1 : 0); }

void bph_r1cs_transcript_state(const char* label, uint8_t* out203) {
  Transcript t{std::string(label)};
  t.r1cs_domain_sep();
  t.export_state(out203);
}

}  // extern "C"
