// Fused Fr vector kernels of the R1CS prover and verifier, and the composite MSM they feed.
//
// Replaces the per-element host loops of /root/reference/src/r1cs/prover.rs:458-486 (blinded vector
// polynomials l(x), r(x)), :524-535 and :552-563 (evaluation at x, padding, G/H factor vectors) and
// of src/r1cs/verifier.rs:341-390 (y^-i * wR, delta, g_scalars, h_scalars).  Constraint flattening
// (prover.rs:142-184) stays on the host: it is a sparse scatter over LinearCombinations.
#include "common.cuh"
#include "host_fp.h"

namespace bp {

// Challenge-derived constants travel as KERNEL ARGUMENTS (2 KB of power table, up to four scalars): no staging buffer,
// no host-to-device copy and no stream synchronisation around them
template <class Fr> struct YTab { Fr t[64]; };       // y^(2^k) [32] | y^-(2^k) [32], Montgomery form
template <class Fr> struct FrArgs4 { Fr a[4]; };

template <class Fr>
__device__ __forceinline__ Fr pow_tab(const Fr* pw, uint32_t e) {
  Fr acc = Fr::one();
  for (int k = 0; e; k++, e >>= 1)
    if (e & 1) acc = acc * pw[k];
  return acc;
}

// prover.rs:469-486 : l1 = a_L + y^-i * wR ; r0 = wO - y^i ; r1 = y^i * a_R + wL ; r3 = y^i * s_R
// (l2 = a_O and l3 = s_L are the inputs themselves).  tab = y^(2^k) [32] | y^-(2^k) [32]
template <class Fr>
__global__ void __launch_bounds__(128) k_r1cs_polys(uint32_t n, const __grid_constant__ YTab<Fr> ytab, const Fr* __restrict__ aL,
                                                    const Fr* __restrict__ aR, const Fr* __restrict__ sR, const Fr* __restrict__ wL,
                                                    const Fr* __restrict__ wR, const Fr* __restrict__ wO, Fr* __restrict__ l1,
                                                    Fr* __restrict__ r0, Fr* __restrict__ r1, Fr* __restrict__ r3) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fr* tab = ytab.t;
  Fr yi = pow_tab(tab, i), yinv = pow_tab(tab + 32, i);
  store_vec(l1 + i, load_vec(aL + i) + yinv * load_vec(wR + i));
  store_vec(r0 + i, load_vec(wO + i) - yi);
  store_vec(r1 + i, yi * load_vec(aR + i) + load_vec(wL + i));
  store_vec(r3 + i, yi * load_vec(sR + i));
}

// prover.rs:524-535,552-563 : l_vec = l(x) | 0^pad ; r_vec = r(x) | (-y^i)_{i >= n} ;
// G_factors = 1^{n1} | u^{N-n1} ; H_factors[i] = y^-i * G_factors[i].   args = {x, u}
template <class Fr>
__global__ void __launch_bounds__(128) k_r1cs_eval(uint32_t n, uint32_t n1, uint32_t N, const __grid_constant__ YTab<Fr> ytab,
                                                   const __grid_constant__ FrArgs4<Fr> fargs, const Fr* __restrict__ l1, const Fr* __restrict__ l2,
                                                   const Fr* __restrict__ l3, const Fr* __restrict__ r0, const Fr* __restrict__ r1,
                                                   const Fr* __restrict__ r3, Fr* __restrict__ lvec, Fr* __restrict__ rvec,
                                                   Fr* __restrict__ Gf, Fr* __restrict__ Hf) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const Fr* tab = ytab.t;
  const Fr x = fargs.a[0], u = fargs.a[1];
  if (i < n) {
    // VecPoly3::eval (vector_poly.rs:99-106) with l.0 = 0 and r.2 = 0
    store_vec(lvec + i, x * (load_vec(l1 + i) + x * (load_vec(l2 + i) + x * load_vec(l3 + i))));
    store_vec(rvec + i, load_vec(r0 + i) + x * (load_vec(r1 + i) + x * (x * load_vec(r3 + i))));
  } else {
    store_vec(lvec + i, Fr::zero());
    store_vec(rvec + i, pow_tab(tab, i).neg());
  }
  Fr gf = i < n1 ? Fr::one() : u;
  store_vec(Gf + i, gf);
  store_vec(Hf + i, pow_tab(tab + 32, i) * gf);
}

// verifier.rs:341-390.  args = {x, a, b, u}.  Outputs gh = g_scalars[N] | h_scalars[N] and
// ywr[i] = y^-i * wR[i] (i < n) for the delta inner product.  wL/wR/wO have length n.
template <class Fr>
__global__ void __launch_bounds__(128) k_r1cs_verifier_scalars(uint32_t n, uint32_t n1, uint32_t N, const __grid_constant__ YTab<Fr> ytab,
                                                               const __grid_constant__ FrArgs4<Fr> fargs, const Fr* __restrict__ wL,
                                                               const Fr* __restrict__ wR, const Fr* __restrict__ wO,
                                                               const Fr* __restrict__ s, Fr* __restrict__ gh, Fr* __restrict__ ywr) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const Fr* tab = ytab.t;
  const Fr x = fargs.a[0], a = fargs.a[1], b = fargs.a[2], u = fargs.a[3];
  const Fr yinv = pow_tab(tab + 32, i);
  Fr wl = Fr::zero(), wr = Fr::zero(), wo = Fr::zero();
  if (i < n) { wl = load_vec(wL + i); wr = load_vec(wR + i); wo = load_vec(wO + i); }
  Fr yw = wr * yinv;
  if (i < n) store_vec(ywr + i, yw);
  Fr g = x * yw - a * load_vec(s + i);
  Fr h = yinv * (x * wl + wo - b * load_vec(s + (N - 1 - i))) - Fr::one();
  if (i >= n1) { g = u * g; h = u * h; }
  store_vec(gh + i, g);
  store_vec(gh + N + i, h);
}

// y^(2^k) | y^-(2^k), k < 32, computed on the host (one inversion, 62 squarings) into a by-value kernel argument.
// Host and device share the limb radix and the Montgomery constant, so the host values are the device values.
template <class Curve>
static void ypow_tables(const uint8_t* y_be, YTab<typename Curve::Fr>* tab) {
  using HF = host::HFp<typename std::conditional<Curve::ID == BPGPU_BLS12_381, BlsFr, BnFr>::type>;
  static_assert(sizeof(HF) == sizeof(typename Curve::Fr), "same layout");
  HF* out = reinterpret_cast<HF*>(tab->t);
  HF cur = HF::from_be(y_be, Curve::MODBYTES), cinv = cur.inv();
  for (int k = 0; k < 32; k++) { out[k] = cur; out[32 + k] = cinv; cur = cur.sqr(); cinv = cinv.sqr(); }
}
template <class Curve>
static void fr_args_host(const uint8_t* be, int cnt, FrArgs4<typename Curve::Fr>* args) {
  using HF = host::HFp<typename std::conditional<Curve::ID == BPGPU_BLS12_381, BlsFr, BnFr>::type>;
  HF* out = reinterpret_cast<HF*>(args->a);
  for (int k = 0; k < 4; k++) out[k] = k < cnt ? HF::from_be(be + (size_t)k * Curve::MODBYTES, Curve::MODBYTES) : HF::zero();
}

template <class Curve>
static int prover_polys_t(bpgpu_ctx* ctx, size_t n, const void* aL, const void* aR, const void* sR, const void* wL, const void* wR,
                          const void* wO, const uint8_t* y_be, void* l1, void* r0, void* r1, void* r3) {
  using Fr = typename Curve::Fr;
  YTab<Fr> tab;
  ypow_tables<Curve>(y_be, &tab);
  if (n) k_r1cs_polys<Fr><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((uint32_t)n, tab, (const Fr*)aL, (const Fr*)aR, (const Fr*)sR,
                                                                              (const Fr*)wL, (const Fr*)wR, (const Fr*)wO, (Fr*)l1, (Fr*)r0,
                                                                              (Fr*)r1, (Fr*)r3);
  ctx->launches++;
  return launch_check(ctx, "k_r1cs_polys");
}

template <class Curve>
static int prover_eval_t(bpgpu_ctx* ctx, size_t n, size_t n1, size_t N, const void* l1, const void* l2, const void* l3, const void* r0,
                         const void* r1, const void* r3, const uint8_t* x_be, const uint8_t* u_be, const uint8_t* y_be, void* lvec,
                         void* rvec, void* Gf, void* Hf) {
  using Fr = typename Curve::Fr;
  YTab<Fr> tab;
  FrArgs4<Fr> args;
  ypow_tables<Curve>(y_be, &tab);
  uint8_t both[2 * 48];
  memcpy(both, x_be, Curve::MODBYTES);
  memcpy(both + Curve::MODBYTES, u_be, Curve::MODBYTES);
  fr_args_host<Curve>(both, 2, &args);
  k_r1cs_eval<Fr><<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>((uint32_t)n, (uint32_t)n1, (uint32_t)N, tab, args, (const Fr*)l1,
                                                                      (const Fr*)l2, (const Fr*)l3, (const Fr*)r0, (const Fr*)r1,
                                                                      (const Fr*)r3, (Fr*)lvec, (Fr*)rvec, (Fr*)Gf, (Fr*)Hf);
  ctx->launches++;
  return launch_check(ctx, "k_r1cs_eval");
}

template <class Curve>
static int verifier_scalars_t(bpgpu_ctx* ctx, size_t n, size_t n1, size_t N, const void* wL, const void* wR, const void* wO, const void* s,
                              const uint8_t* y_be, const uint8_t* xabu_be, void* gh, void* ywr) {
  using Fr = typename Curve::Fr;
  YTab<Fr> tab;
  FrArgs4<Fr> args;
  ypow_tables<Curve>(y_be, &tab);
  fr_args_host<Curve>(xabu_be, 4, &args);
  k_r1cs_verifier_scalars<Fr><<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>((uint32_t)n, (uint32_t)n1, (uint32_t)N, tab, args,
                                                                                  (const Fr*)wL, (const Fr*)wR, (const Fr*)wO,
                                                                                  (const Fr*)s, (Fr*)gh, (Fr*)ywr);
  ctx->launches++;
  return launch_check(ctx, "k_r1cs_verifier_scalars");
}

// ---- composite MSM: concatenates device/host point and scalar sources into one MSM.
// Parts whose points carry window tables (bpgpu_points_precompute, or a host point that is a cached fixed base such as the
// Pedersen h) go through the table path: no doublings, no buckets, no Horner tail.  Everything else is gathered into one
// general Pippenger run.  The two partial results are added in the host finish.
template <class Curve>
static int msm_parts_t(bpgpu_ctx* ctx, const bpgpu_msm_part* parts, size_t np, uint8_t* out_xy) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  size_t total = 0, host_scalars = 0;
  for (size_t k = 0; k < np; k++) { total += parts[k].n; if (!parts[k].scalars) host_scalars += parts[k].n; }
  int rc;
  if ((rc = ctx->parts_pts.reserve(total * sizeof(Affine<Fq>) + 64))) return rc;
  if ((rc = ctx->parts_scl.reserve((total + host_scalars) * sizeof(Fr) + 64))) return rc;
  Affine<Fq>* P = (Affine<Fq>*)ctx->parts_pts.p;
  Fr* S = (Fr*)ctx->parts_scl.p;
  Fr* S2 = S + total;                      // Montgomery copies of host scalars that feed table segments
  TableSeg segs[TBL_MAX_SEGS];
  int nsegs = 0;
  size_t off = 0, off2 = 0;
  // The table path is used when it replaces the general run ENTIRELY.  An MSM that needs a general run anyway (verifier:
  // proof points next to G, H) is cheaper as one general run: the extra launch pair costs latency at small n, and 64 table
  // entries per term cost 2.5x the bucket method's work at large n (measured both ways).
  size_t table_terms = 0, general_terms = 0;
  for (size_t k = 0; k < np; k++) {
    const bpgpu_msm_part& p = parts[k];
    if (p.n == 0) continue;
    bool t = (p.points && p.points->table) || (!p.points && p.n == 1 && fixed_table_lookup(ctx, p.host_points_xy));
    (t ? table_terms : general_terms) += p.n;
  }
  const bool use_tables = general_terms == 0 && table_terms > 0;
  for (size_t k = 0; k < np; k++) {
    const bpgpu_msm_part& p = parts[k];
    if (p.n == 0) continue;
    const void* tbl = nullptr;
    if (use_tables && nsegs < TBL_MAX_SEGS) {
      if (p.points && p.points->table) tbl = (const uint8_t*)p.points->table + p.points_off * TBL_ENTRIES * sizeof(Affine<Fq>);
      else if (!p.points && p.n == 1) tbl = fixed_table_lookup(ctx, p.host_points_xy);
    }
    if (tbl) {
      const void* sc;
      if (p.scalars) sc = (const Fr*)p.scalars->d + p.scalars_off;
      else { if ((rc = scalars_from_host<Curve>(ctx, p.host_scalars_be, p.n, 1, S2 + off2))) return rc; sc = S2 + off2; off2 += p.n; }
      segs[nsegs++] = TableSeg{tbl, sc, (uint32_t)p.n, 1};
      continue;
    }
    if (p.points) {
      BP_CUDA_OK(cudaMemcpyAsync(P + off, (const Affine<Fq>*)p.points->d + p.points_off, p.n * sizeof(Affine<Fq>), cudaMemcpyDeviceToDevice,
                                 ctx->stream));
    } else {
      if ((rc = points_from_host<Curve>(ctx, p.host_points_xy, p.n, P + off))) return rc;
    }
    if (p.scalars) {
      BP_CUDA_OK(cudaMemcpyAsync(S + off, (const Fr*)p.scalars->d + p.scalars_off, p.n * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      if ((rc = scalars_from_host<Curve>(ctx, p.host_scalars_be, p.n, 1, S + off))) return rc;
    }
    off += p.n;
  }
  return msm_mixed_to_host(ctx, segs, nsegs, P, S, true, off, out_xy);
}

// Several independent composite MSMs (prover.rs:347-362: A_I, A_O, S).  When every part of every MSM has window tables
// they are evaluated by ONE launch pair (one group per MSM) with one D2H and one shared inversion; otherwise one by one.
template <class Curve>
static int msm_parts_batch_tables(bpgpu_ctx* ctx, const bpgpu_msm_part* parts, const size_t* counts, size_t nmsm, uint8_t* out_xy, bool* done) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  *done = false;
  if (nmsm > TBL_MAX_GROUPS) return BPGPU_OK;
  size_t np = 0, host_scalars = 0;
  for (size_t m = 0; m < nmsm; m++) np += counts[m];
  size_t live = 0;
  for (size_t k = 0; k < np; k++) {
    const bpgpu_msm_part& p = parts[k];
    if (p.n == 0) continue;
    live++;
    bool t = (p.points && p.points->table) || (!p.points && p.n == 1 && fixed_table_lookup(ctx, p.host_points_xy));
    if (!t) return BPGPU_OK;
    if (!p.scalars) host_scalars += p.n;
  }
  if (live == 0 || live > (size_t)TBL_MAX_SEGS) return BPGPU_OK;
  int rc;
  if ((rc = ctx->parts_scl.reserve(host_scalars * sizeof(Fr) + 64))) return rc;
  Fr* S2 = (Fr*)ctx->parts_scl.p;
  // all host scalars of all parts in ONE upload: gather them on the host first
  std::vector<uint8_t> hs(host_scalars * Curve::MODBYTES + 1);
  size_t ho = 0;
  for (size_t k = 0; k < np; k++)
    if (parts[k].n && !parts[k].scalars) { memcpy(hs.data() + ho * Curve::MODBYTES, parts[k].host_scalars_be, parts[k].n * Curve::MODBYTES); ho += parts[k].n; }
  if (host_scalars && (rc = scalars_from_host<Curve>(ctx, hs.data(), host_scalars, 1, S2))) return rc;
  TableSeg segs[TBL_MAX_SEGS];
  int nsegs = 0;
  size_t k = 0, so = 0;
  for (size_t m = 0; m < nmsm; m++) {
    for (size_t q = 0; q < counts[m]; q++, k++) {
      const bpgpu_msm_part& p = parts[k];
      if (p.n == 0) continue;
      const void* tbl = p.points ? (const uint8_t*)p.points->table + p.points_off * TBL_ENTRIES * sizeof(Affine<Fq>)
                                 : fixed_table_lookup(ctx, p.host_points_xy);
      const void* sc;
      if (p.scalars) sc = (const Fr*)p.scalars->d + p.scalars_off; else { sc = S2 + so; so += p.n; }
      segs[nsegs++] = TableSeg{tbl, sc, (uint32_t)p.n, 1, (int)m};
    }
  }
  uint8_t* outs[TBL_MAX_GROUPS];
  for (size_t m = 0; m < nmsm; m++) outs[m] = out_xy + m * 2 * Curve::MODBYTES;
  rc = msm_tables_to_host(ctx, segs, nsegs, (int)nmsm, outs);
  *done = rc == BPGPU_OK;
  return rc;
}

}  // namespace bp

using namespace bp;

static bool srange(const bpgpu_scalars* s, size_t n) { return s && n <= s->n; }

extern "C" {

int bpgpu_r1cs_prover_polys(bpgpu_ctx* ctx, size_t n, const bpgpu_scalars* a_L, const bpgpu_scalars* a_R, const bpgpu_scalars* s_R,
                            const bpgpu_scalars* wL, const bpgpu_scalars* wR, const bpgpu_scalars* wO, const uint8_t* y_be,
                            bpgpu_scalars** l1, bpgpu_scalars** r0, bpgpu_scalars** r1, bpgpu_scalars** r3) {
  if (!ctx || !y_be || !l1 || !r0 || !r1 || !r3) return BPGPU_E_ARG;
  const bpgpu_scalars* in[6] = {a_L, a_R, s_R, wL, wR, wO};
  for (auto v : in) if (!srange(v, n)) return BPGPU_E_LEN;
  if (n >= (1ull << 31)) return BPGPU_E_ARG;
  bpgpu_scalars** outs[4] = {l1, r0, r1, r3};
  int rc = scalars_alloc_many(ctx, n, 4, outs);
  if (!rc)
    rc = ctx->curve == BPGPU_BLS12_381
             ? prover_polys_t<Bls>(ctx, n, a_L->d, a_R->d, s_R->d, wL->d, wR->d, wO->d, y_be, (*l1)->d, (*r0)->d, (*r1)->d, (*r3)->d)
             : prover_polys_t<Bn>(ctx, n, a_L->d, a_R->d, s_R->d, wL->d, wR->d, wO->d, y_be, (*l1)->d, (*r0)->d, (*r1)->d, (*r3)->d);
  if (rc) for (auto o : outs) { bpgpu_scalars_free(*o); *o = nullptr; }
  return rc;
}

int bpgpu_r1cs_prover_eval(bpgpu_ctx* ctx, size_t n, size_t n1, size_t padded_n, const bpgpu_scalars* l1, const bpgpu_scalars* l2,
                           const bpgpu_scalars* l3, const bpgpu_scalars* r0, const bpgpu_scalars* r1, const bpgpu_scalars* r3,
                           const uint8_t* x_be, const uint8_t* u_be, const uint8_t* y_be, bpgpu_scalars** l_vec, bpgpu_scalars** r_vec,
                           bpgpu_scalars** G_factors, bpgpu_scalars** H_factors) {
  if (!ctx || !x_be || !u_be || !y_be || !l_vec || !r_vec || !G_factors || !H_factors) return BPGPU_E_ARG;
  const bpgpu_scalars* in[6] = {l1, l2, l3, r0, r1, r3};
  for (auto v : in) if (!srange(v, n)) return BPGPU_E_LEN;
  if (n > padded_n || n1 > n || padded_n >= (1ull << 31) || padded_n == 0) return BPGPU_E_ARG;
  bpgpu_scalars** outs[4] = {l_vec, r_vec, G_factors, H_factors};
  int rc = scalars_alloc_many(ctx, padded_n, 4, outs);
  if (!rc)
    rc = ctx->curve == BPGPU_BLS12_381
             ? prover_eval_t<Bls>(ctx, n, n1, padded_n, l1->d, l2->d, l3->d, r0->d, r1->d, r3->d, x_be, u_be, y_be, (*l_vec)->d, (*r_vec)->d,
                                  (*G_factors)->d, (*H_factors)->d)
             : prover_eval_t<Bn>(ctx, n, n1, padded_n, l1->d, l2->d, l3->d, r0->d, r1->d, r3->d, x_be, u_be, y_be, (*l_vec)->d, (*r_vec)->d,
                                 (*G_factors)->d, (*H_factors)->d);
  if (rc) for (auto o : outs) { bpgpu_scalars_free(*o); *o = nullptr; }
  return rc;
}

int bpgpu_r1cs_verifier_scalars(bpgpu_ctx* ctx, size_t n, size_t n1, size_t padded_n, const bpgpu_scalars* wL, const bpgpu_scalars* wR,
                                const bpgpu_scalars* wO, const bpgpu_scalars* s, const uint8_t* y_be, const uint8_t* x_be,
                                const uint8_t* a_be, const uint8_t* b_be, const uint8_t* u_be, bpgpu_scalars** gh_scalars,
                                uint8_t* delta_be) {
  if (!ctx || !y_be || !x_be || !a_be || !b_be || !u_be || !gh_scalars || !delta_be) return BPGPU_E_ARG;
  if (!srange(wL, n) || !srange(wR, n) || !srange(wO, n) || !srange(s, padded_n)) return BPGPU_E_LEN;
  if (n > padded_n || n1 > n || padded_n >= (1ull << 31) || padded_n == 0) return BPGPU_E_ARG;
  *gh_scalars = nullptr;
  const int mb = bpgpu_modbytes(ctx->curve);
  uint8_t args[4 * 48];
  memcpy(args, x_be, mb); memcpy(args + mb, a_be, mb); memcpy(args + 2 * mb, b_be, mb); memcpy(args + 3 * mb, u_be, mb);
  bpgpu_scalars* ywr = nullptr;
  int rc = bpgpu_scalars_alloc(ctx, 2 * padded_n, gh_scalars);
  if (!rc) rc = bpgpu_scalars_alloc(ctx, n, &ywr);
  if (!rc)
    rc = ctx->curve == BPGPU_BLS12_381
             ? verifier_scalars_t<Bls>(ctx, n, n1, padded_n, wL->d, wR->d, wO->d, s->d, y_be, args, (*gh_scalars)->d, ywr->d)
             : verifier_scalars_t<Bn>(ctx, n, n1, padded_n, wL->d, wR->d, wO->d, s->d, y_be, args, (*gh_scalars)->d, ywr->d);
  // delta = <y^-n * wR, wL>  (verifier.rs:350-352)
  if (!rc) rc = bpgpu_fr_inner_product(ctx, ywr, 0, wL, 0, n, delta_be);
  bpgpu_scalars_free(ywr);
  if (rc) { bpgpu_scalars_free(*gh_scalars); *gh_scalars = nullptr; }
  return rc;
}

int bpgpu_msm_parts_batch(bpgpu_ctx* ctx, const bpgpu_msm_part* parts, const size_t* counts, size_t nmsm, uint8_t* out_xy) {
  if (!ctx || !parts || !counts || !out_xy) return BPGPU_E_ARG;
  size_t np = 0;
  for (size_t m = 0; m < nmsm; m++) np += counts[m];
  for (size_t k = 0; k < np; k++) {
    const bpgpu_msm_part& p = parts[k];
    if (p.n == 0) continue;
    if (p.points ? !(p.points_off <= p.points->n && p.n <= p.points->n - p.points_off) : !p.host_points_xy) return BPGPU_E_LEN;
    if (p.scalars ? !(p.scalars_off <= p.scalars->n && p.n <= p.scalars->n - p.scalars_off) : !p.host_scalars_be) return BPGPU_E_LEN;
  }
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  bool done = false;
  int rc = ctx->curve == BPGPU_BLS12_381 ? msm_parts_batch_tables<Bls>(ctx, parts, counts, nmsm, out_xy, &done)
                                        : msm_parts_batch_tables<Bn>(ctx, parts, counts, nmsm, out_xy, &done);
  if (rc || done) return rc;
  const size_t pb = 2 * (size_t)bpgpu_modbytes(ctx->curve);
  size_t k = 0;
  for (size_t m = 0; m < nmsm; m++) {
    if ((rc = bpgpu_msm_parts(ctx, parts + k, counts[m], out_xy + m * pb))) return rc;
    k += counts[m];
  }
  return BPGPU_OK;
}

int bpgpu_msm_parts(bpgpu_ctx* ctx, const bpgpu_msm_part* parts, size_t nparts, uint8_t* out_xy) {
  if (!ctx || (!parts && nparts) || !out_xy) return BPGPU_E_ARG;
  for (size_t k = 0; k < nparts; k++) {
    const bpgpu_msm_part& p = parts[k];
    if (p.n == 0) continue;
    if (p.points ? !(p.points_off <= p.points->n && p.n <= p.points->n - p.points_off) : !p.host_points_xy) return BPGPU_E_LEN;
    if (p.scalars ? !(p.scalars_off <= p.scalars->n && p.n <= p.scalars->n - p.scalars_off) : !p.host_scalars_be) return BPGPU_E_LEN;
  }
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  return ctx->curve == BPGPU_BLS12_381 ? msm_parts_t<Bls>(ctx, parts, nparts, out_xy) : msm_parts_t<Bn>(ctx, parts, nparts, out_xy);
}

}  // extern "C"
