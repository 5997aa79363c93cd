// Batches of independent verification MSMs that share the fixed generators: `batch` sums
//     R_b = sum_i f[b][i] * F_i  +  sum_j v[b][j] * V[b][j]          (b < batch)
// reduced to one byte each: is R_b the identity?
//
// This is Verifier::verify's final check (/root/reference/src/r1cs/verifier.rs:451-456: one MSM per proof, Ok iff
// the result is the identity) for MANY proofs in one call -- BASELINE.json's "batch verification of 4096 independent
// range proofs".  The reference has no batch API; every proof keeps its own verdict (no random-linear-combination
// merging).  F_i are the proof system's generators G[..N], H[..N], g, h (window tables: no doublings); V[b][j] are the
// proof's own points A_I1.., V_k, T_i, L_k, R_k, which need real scalar multiplications:
//   k_batch_fixed : one block per proof, sum of 32 table entries per fixed term, block tree
//   k_batch_multiples / k_batch_windows / k_batch_horner : Straus with 4-bit windows for the proof's own points, laid
//                   out so that each stage has one independent item per lane (see below)
// Nothing returns to the host but `batch` verdict bytes.
#include <stdlib.h>
#include <string.h>

#include "batchsum.cuh"

namespace bp {

// the proof's own points: Straus with 4-bit windows (15 multiples per point, 64 windows), independent of the table geometry
static const int VAR_WINDOWS = 64;
static const int VAR_DIGITS = 15;

// Variable part, three launches so that every stage has one independent item per LANE:
//   k_batch_multiples : thread per (proof, point): the 15 small multiples of the point (14 mixed additions)
//   k_batch_windows   : thread per (proof, 4-bit window): sum over the points of the multiple its digit selects
//   k_batch_horner    : thread per proof: the 252-doubling Horner chain over the 64 window sums, plus the fixed part,
//                       then the verdict.  The chain is serial per proof, but 4096 proofs are 4096 chains: 128 warps
//                       running flat out instead of one lane of a block holding its registers for 3 ms.
template <class Curve>
__global__ void __launch_bounds__(64) k_batch_multiples(size_t total, const Affine<typename Curve::Fq>* __restrict__ pts,
                                                        XYZZ<typename Curve::Fq>* __restrict__ mult) {
  using Fq = typename Curve::Fq;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // (proof, point) flattened
  if (t >= total) return;
  const Affine<Fq> a = load_vec(pts + t);
  XYZZ<Fq> acc = XYZZ<Fq>::from_affine(a);
  XYZZ<Fq>* M = mult + t * VAR_DIGITS;
#pragma unroll 1
  for (int d = 0; d < VAR_DIGITS; d++) {
    store_vec(M + d, acc);
    if (d + 1 < VAR_DIGITS) acc.madd(a);
  }
}

// BLS12-381 only: the scalars of the proofs' own points are split k = k1 + k2 * lambda (curves.cuh), k1 in limbs 0..3 and k2
// in limbs 4..7 of the same slot, so that windows 0..31 are the 4-bit windows of k1 on P and windows 32..63 those of k2 on
// phi(P) = (beta x, y): same number of window sums, HALF the doubling chain in k_batch_horner.
template <class Curve>
__global__ void __launch_bounds__(256) k_glv_split(size_t total, typename Curve::Fr* __restrict__ scal) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  typename Curve::Fr k = load_vec(scal + t), o;
  glv_bls_split(k.v, o.v);
  store_vec(scal + t, o);
}

// wsum[w * batch + b] = sum_i M[b][i][digit_w(s[b][i]) - 1]   (glv: windows >= 32 read k2's digits and add phi of the multiple)
template <class Curve>
__global__ void __launch_bounds__(128) k_batch_windows(size_t batch, uint32_t vn, const typename Curve::Fr* __restrict__ scal,
                                                       const XYZZ<typename Curve::Fq>* __restrict__ mult,
                                                       XYZZ<typename Curve::Fq>* __restrict__ wsum, int glv) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * VAR_WINDOWS) return;
  const size_t b = t / VAR_WINDOWS;
  const uint32_t w = (uint32_t)(t - b * VAR_WINDOWS);
  const Fr* sc = scal + b * vn;
  const XYZZ<Fq>* M = mult + b * (size_t)vn * VAR_DIGITS;
  const bool phi = glv && w >= VAR_WINDOWS / 2;
  Fq beta = Fq::zero();
  if (phi) beta = glv_beta<Curve>();
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
#pragma unroll 1
  for (uint32_t i = 0; i < vn; i++) {
    const uint32_t d = (sc[i].v[w >> 3] >> ((w & 7) * 4)) & 15u;
    if (d) {
      XYZZ<Fq> q = load_vec(M + (size_t)i * VAR_DIGITS + (d - 1));
      if (phi) q.x = Fq::mulc(q.x, beta);
      acc.add(q);
    }
  }
  store_vec(wsum + (size_t)w * batch + b, acc);
}

template <class Curve>
__global__ void __launch_bounds__(32) k_batch_horner(size_t batch, uint32_t vn, const XYZZ<typename Curve::Fq>* __restrict__ wsum,
                                                     const XYZZ<typename Curve::Fq>* __restrict__ fixed_sum,
                                                     uint8_t* __restrict__ is_identity, int glv) {
  using Fq = typename Curve::Fq;
  const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  if (vn) {
    // glv: windows w and w + 32 carry the same weight 16^w -> 31 steps of 4 doublings instead of 63
    const int top = glv ? VAR_WINDOWS / 2 : VAR_WINDOWS;
#pragma unroll 1
    for (int w = top - 1; w >= 0; w--) {
      if (w != top - 1) { acc.dbl(); acc.dbl(); acc.dbl(); acc.dbl(); }
      XYZZ<Fq> q = load_vec(wsum + (size_t)w * batch + b);
      acc.add(q);
      if (glv) { XYZZ<Fq> q2 = load_vec(wsum + (size_t)(w + top) * batch + b); acc.add(q2); }
    }
  }
  XYZZ<Fq> f = load_vec(fixed_sum + b);
  acc.add(f);
  is_identity[b] = acc.is_inf() ? 1 : 0;
}

// the launches for one slab whose inputs are already on the device: fixed scalars d_fs (cnt x F, canonical), the proofs' own
// points d_vp / scalars d_vs (cnt x vn); scratch d_sum (cnt), d_m (cnt x vn x 15), d_w (64 x cnt); verdict bytes to d_v
template <class Curve>
int batch_identity_launch(bpgpu_ctx* ctx, const FixedRuns& runs, uint32_t F, size_t cnt, const void* d_fs, const void* d_vp, const void* d_vs,
                          uint32_t vn, void* d_sum_, void* d_m_, void* d_w_, uint8_t* d_v) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  XYZZ<Fq>*d_sum = (XYZZ<Fq>*)d_sum_, *d_m = (XYZZ<Fq>*)d_m_, *d_w = (XYZZ<Fq>*)d_w_;
  static const bool prof = getenv("BPGPU_PROFILE") != nullptr;
  cudaEvent_t ev[3];
  if (prof) { for (auto& e : ev) cudaEventCreate(&e); cudaEventRecord(ev[0], ctx->stream); }
  if (cnt >= 2048 && F * 8 >= 256)          // thousands of rows: one warp per row (batchsum.cuh)
    k_batch_fixed_warp<Curve><<<(unsigned)((cnt + 3) / 4), 128, 0, ctx->stream>>>(runs, F, (uint32_t)cnt, (const Fr*)d_fs, 0, d_sum);
  else
    k_batch_fixed<Curve><<<(unsigned)cnt, BATCH_FIXED_THREADS, 0, ctx->stream>>>(runs, F, (const Fr*)d_fs, 0, d_sum);
  if (prof) cudaEventRecord(ev[1], ctx->stream);
  const int glv = Curve::ID == BPGPU_BLS12_381 ? 1 : 0;
  if (vn) {
    const size_t np = cnt * vn, nw = cnt * VAR_WINDOWS;
    if (glv) {                                  // d_vs is this call's scratch: split in place
      k_glv_split<Curve><<<(unsigned)((np + 255) / 256), 256, 0, ctx->stream>>>(np, (Fr*)const_cast<void*>(d_vs));
      ctx->launches++;
    }
    k_batch_multiples<Curve><<<(unsigned)((np + 63) / 64), 64, 0, ctx->stream>>>(np, (const Affine<Fq>*)d_vp, d_m);
    k_batch_windows<Curve><<<(unsigned)((nw + 127) / 128), 128, 0, ctx->stream>>>(cnt, vn, (const Fr*)d_vs, d_m, d_w, glv);
    ctx->launches += 2;
  }
  k_batch_horner<Curve><<<(unsigned)((cnt + 31) / 32), 32, 0, ctx->stream>>>(cnt, vn, d_w, d_sum, d_v, glv);
  if (prof) {
    cudaEventRecord(ev[2], ctx->stream);
    cudaEventSynchronize(ev[2]);
    float a, b;
    cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]);
    fprintf(stderr, "[bpgpu batch_identity n=%zu F=%u vn=%u] fixed=%.3f ms var=%.3f ms\n", cnt, F, vn, a, b);
    for (auto& e : ev) cudaEventDestroy(e);
  }
  ctx->launches += 2;
  return launch_check(ctx, "batch_identity");
}
template int batch_identity_launch<Bls>(bpgpu_ctx*, const FixedRuns&, uint32_t, size_t, const void*, const void*, const void*, uint32_t, void*, void*,
                                        void*, uint8_t*);
template int batch_identity_launch<Bn>(bpgpu_ctx*, const FixedRuns&, uint32_t, size_t, const void*, const void*, const void*, uint32_t, void*, void*,
                                       void*, uint8_t*);

template <class Curve>
BatchScratch batch_scratch_layout(size_t slab, uint32_t F, uint32_t vn) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  BatchScratch L;
  size_t off = 0;
  L.fs = off; off += up(slab * F * sizeof(Fr));
  L.vs = off; off += up(slab * vn * sizeof(Fr));
  L.vp = off; off += up(slab * vn * sizeof(Affine<Fq>));
  L.sum = off; off += up(slab * sizeof(XYZZ<Fq>));
  L.v = off; off += up(slab);
  L.m = off; off += up(slab * (size_t)vn * VAR_DIGITS * sizeof(XYZZ<Fq>));
  L.w = off; off += vn ? up(slab * (size_t)VAR_WINDOWS * sizeof(XYZZ<Fq>)) : 0;
  L.total = off + 256;
  return L;
}
template BatchScratch batch_scratch_layout<Bls>(size_t, uint32_t, uint32_t);
template BatchScratch batch_scratch_layout<Bn>(size_t, uint32_t, uint32_t);

template <class Curve>
static int batch_identity_t(bpgpu_ctx* ctx, const FixedRuns& runs, uint32_t F, size_t batch, const uint8_t* fixed_scalars_be,
                            const uint8_t* var_points_xy, const uint8_t* var_scalars_be, uint32_t vn, uint8_t* is_identity) {
  const size_t SLAB = 4096;                               // proofs per launch pair: bounds the multiples scratch
  int rc;
  const size_t slab = batch < SLAB ? batch : SLAB;
  const BatchScratch L = batch_scratch_layout<Curve>(slab, F, vn);
  if ((rc = ctx->msm_b.reserve(L.total))) return rc;
  uint8_t* base = (uint8_t*)ctx->msm_b.p;
  const int mb = Curve::MODBYTES;
  for (size_t lo = 0; lo < batch; lo += slab) {
    const size_t cnt = batch - lo < slab ? batch - lo : slab;
    // canonical (non-Montgomery) scalars: the kernels only take digits
    if ((rc = scalars_from_host<Curve>(ctx, fixed_scalars_be + lo * F * mb, cnt * F, 0, base + L.fs))) return rc;
    if (vn) {
      if ((rc = scalars_from_host<Curve>(ctx, var_scalars_be + lo * vn * mb, cnt * vn, 0, base + L.vs))) return rc;
      if ((rc = points_from_host<Curve>(ctx, var_points_xy + lo * vn * 2 * mb, cnt * vn, base + L.vp))) return rc;
    }
    if ((rc = batch_identity_launch<Curve>(ctx, runs, F, cnt, base + L.fs, base + L.vp, base + L.vs, vn, base + L.sum, base + L.m, base + L.w,
                                           base + L.v)))
      return rc;
    BP_CUDA_OK(cudaMemcpyAsync(is_identity + lo, base + L.v, cnt, cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA_OK(stream_sync(ctx));
    if ((rc = inputs_ok(ctx))) return rc;                 // a var point that is not a point (callers validate per proof first)
  }
  return BPGPU_OK;
}

// the fixed terms of a batch call: device tables (bpgpu_points_precompute) or single host bases (tables cached in the ctx)
int resolve_fixed_runs(bpgpu_ctx* ctx, const bpgpu_fixed_run* runs, size_t nruns, FixedRuns* out, uint32_t* F_out) {
  if (nruns > (size_t)TBL_MAX_SEGS) return BPGPU_E_ARG;
  const size_t asz = ctx->curve == BPGPU_BLS12_381 ? sizeof(Affine<Bls::Fq>) : sizeof(Affine<Bn::Fq>);
  FixedRuns& fr = *out;
  fr.nruns = 0;
  uint32_t F = 0;
  for (size_t k = 0; k < nruns; k++) {
    const bpgpu_fixed_run& r = runs[k];
    if (r.n == 0) continue;
    const void* tbl = nullptr;
    const void* tbl16 = nullptr;
    if (r.points) {
      if (!(r.off <= r.points->n && r.n <= r.points->n - r.off)) return BPGPU_E_LEN;
      if (!r.points->table) return BPGPU_E_ARG;                       // bpgpu_points_precompute first
      tbl = (const uint8_t*)r.points->table + r.off * TBL_ENTRIES * asz;
      if (r.points->table16) tbl16 = (const uint8_t*)r.points->table16 + r.off * TBL16_ENTRIES * asz;
    } else {
      if (!r.host_base_xy || r.n != 1) return BPGPU_E_ARG;
      tbl = fixed_table_lookup(ctx, r.host_base_xy);
      if (!tbl) {                                                     // build (and cache) the single-base table
        bpgpu_fixed_bases* fb = nullptr;
        int rc = bpgpu_fixed_bases_get(ctx, r.host_base_xy, 1, &fb);
        if (rc) return rc;
        tbl = fixed_table_lookup(ctx, r.host_base_xy);
      }
    }
    fr.table[fr.nruns] = tbl;
    fr.table16[fr.nruns] = tbl16;
    fr.start[fr.nruns] = F;
    fr.nruns++;
    F += (uint32_t)r.n;
  }
  fr.start[fr.nruns] = F;
  if (fr.nruns == 0) { fr.nruns = 1; fr.table[0] = nullptr; fr.table16[0] = nullptr; fr.start[0] = 0; fr.start[1] = 0; }
  *F_out = F;
  return BPGPU_OK;
}

}  // namespace bp

using namespace bp;

extern "C" int bpgpu_msm_batch_is_identity(bpgpu_ctx* ctx, const bpgpu_fixed_run* runs, size_t nruns, size_t batch,
                                           const uint8_t* fixed_scalars_be, const uint8_t* var_points_xy, const uint8_t* var_scalars_be,
                                           size_t vn, uint8_t* is_identity) {
  if (!ctx || (!runs && nruns) || (batch && (!is_identity || (!fixed_scalars_be && nruns) || (vn && (!var_points_xy || !var_scalars_be)))))
    return BPGPU_E_ARG;
  if (nruns > (size_t)TBL_MAX_SEGS || vn > 4096) return BPGPU_E_ARG;
  if (batch == 0) return BPGPU_OK;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  FixedRuns fr;
  uint32_t F = 0;
  int rrc = resolve_fixed_runs(ctx, runs, nruns, &fr, &F);
  if (rrc) return rrc;
  return ctx->curve == BPGPU_BLS12_381
             ? batch_identity_t<Bls>(ctx, fr, F, batch, fixed_scalars_be, var_points_xy, var_scalars_be, (uint32_t)vn, is_identity)
             : batch_identity_t<Bn>(ctx, fr, F, batch, fixed_scalars_be, var_points_xy, var_scalars_be, (uint32_t)vn, is_identity);
}
