// Batches of independent verification MSMs that share the fixed generators: `batch` sums
//     R_b = sum_i f[b][i] * F_i  +  sum_j v[b][j] * V[b][j]          (b < batch)
// reduced to one byte each: is R_b the identity?
//
// This is Verifier::verify's final check (/root/reference/src/r1cs/verifier.rs:451-456: one MSM per proof, Ok iff
// the result is the identity) for MANY proofs in one call -- BASELINE.json's "batch verification of 4096 independent
// range proofs".  The reference has no batch API; every proof keeps its own verdict (no random-linear-combination
// merging).  F_i are the proof system's generators G[..N], H[..N], g, h (window tables: no doublings); V[b][j] are the
// proof's own points A_I1.., V_k, T_i, L_k, R_k, which need real scalar multiplications:
//   k_batch_fixed : one block per proof, sum of 64 table entries per fixed term, block tree
//   k_batch_var   : one block of 64 threads per proof: 15 multiples of every point, one 4-bit window per thread,
//                   then ONE thread runs the 252-doubling Horner chain -- ~3 ms of latency, but for thousands of
//                   proofs at once (the chain is what a GPU is bad at per proof and good at per batch)
// Nothing returns to the host but `batch` verdict bytes.
#include <string.h>

#include "common.cuh"

namespace bp {

struct FixedRuns {
  const void* table[TBL_MAX_SEGS];
  uint32_t start[TBL_MAX_SEGS + 1];
  int nruns;
};

template <class Curve>
__global__ void __launch_bounds__(256) k_batch_fixed(FixedRuns runs, uint32_t F, const typename Curve::Fr* __restrict__ scal,
                                                     XYZZ<typename Curve::Fq>* __restrict__ out) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  __shared__ __align__(16) unsigned char smraw[256 * sizeof(XYZZ<Fq>)];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(smraw);
  const size_t b = blockIdx.x;
  const Fr* sc = scal + b * F;
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  for (uint32_t t = threadIdx.x; t < F * 8; t += blockDim.x) {
    const uint32_t p = t >> 3, j = t & 7;
    int rg = 0;
    while (rg + 1 < runs.nruns && p >= runs.start[rg + 1]) rg++;
    const uint32_t row = p - runs.start[rg];
    const uint32_t limb = sc[p].v[j];                       // canonical scalars (not Montgomery)
    if (!limb) continue;
    const Affine<Fq>* tb = (const Affine<Fq>*)runs.table[rg] + ((size_t)row * TBL_WINDOWS + 8 * j) * TBL_DIGITS;
#pragma unroll 1
    for (int k = 0; k < 8; k++) {
      const uint32_t d = (limb >> (4 * k)) & 15u;
      if (d) acc.madd(load_vec_ro(tb + k * TBL_DIGITS + (d - 1)));
    }
  }
  store_vec(sm + threadIdx.x, acc);
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      XYZZ<Fq> a = load_vec(sm + threadIdx.x), c = load_vec(sm + threadIdx.x + o);
      a.add(c);
      store_vec(sm + threadIdx.x, a);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) store_vec(out + b, load_vec(sm));
}

// one block of 64 threads per proof; mult = scratch for vn * 15 XYZZ multiples per proof
template <class Curve>
__global__ void __launch_bounds__(64) k_batch_var(uint32_t vn, const Affine<typename Curve::Fq>* __restrict__ pts,
                                                  const typename Curve::Fr* __restrict__ scal, XYZZ<typename Curve::Fq>* __restrict__ mult,
                                                  const XYZZ<typename Curve::Fq>* __restrict__ fixed_sum, uint8_t* __restrict__ is_identity) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  __shared__ __align__(16) unsigned char smraw[64 * sizeof(XYZZ<Fq>)];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(smraw);
  const size_t b = blockIdx.x;
  const Affine<Fq>* P = pts + b * vn;
  const Fr* sc = scal + b * vn;
  XYZZ<Fq>* M = mult + b * (size_t)vn * TBL_DIGITS;
  // phase 1: multiples 1..15 of every point (thread i takes points i, i+64, ..)
  for (uint32_t i = threadIdx.x; i < vn; i += blockDim.x) {
    const Affine<Fq> a = load_vec(P + i);
    XYZZ<Fq> acc = XYZZ<Fq>::from_affine(a);
    for (int d = 0; d < TBL_DIGITS; d++) {
      store_vec(M + (size_t)i * TBL_DIGITS + d, acc);
      if (d + 1 < TBL_DIGITS) acc.madd(a);
    }
  }
  __syncthreads();
  // phase 2: thread w sums the window-w digits of all points
  {
    const uint32_t w = threadIdx.x;
    XYZZ<Fq> acc = XYZZ<Fq>::inf();
    for (uint32_t i = 0; i < vn; i++) {
      const uint32_t d = (sc[i].v[w >> 3] >> ((w & 7) * 4)) & 15u;
      if (d) { XYZZ<Fq> q = load_vec(M + (size_t)i * TBL_DIGITS + (d - 1)); acc.add(q); }
    }
    store_vec(sm + w, acc);
  }
  __syncthreads();
  // phase 3: Horner over the 64 windows, then the fixed part, then the verdict
  if (threadIdx.x == 0) {
    XYZZ<Fq> acc = load_vec(sm + 63);
    for (int w = 62; w >= 0; w--) {
      acc.dbl(); acc.dbl(); acc.dbl(); acc.dbl();
      XYZZ<Fq> q = load_vec(sm + w);
      acc.add(q);
    }
    XYZZ<Fq> f = load_vec(fixed_sum + b);
    acc.add(f);
    is_identity[b] = acc.is_inf() ? 1 : 0;
  }
}

template <class Curve>
static int batch_identity_t(bpgpu_ctx* ctx, const FixedRuns& runs, uint32_t F, size_t batch, const uint8_t* fixed_scalars_be,
                            const uint8_t* var_points_xy, const uint8_t* var_scalars_be, uint32_t vn, uint8_t* is_identity) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  const size_t SLAB = 4096;                               // proofs per launch pair: bounds the multiples scratch
  int rc;
  const size_t slab = batch < SLAB ? batch : SLAB;
  // scratch: fixed scalars | var scalars | var points | fixed sums | verdicts | multiples
  const size_t sz_fs = (slab * F * sizeof(Fr) + 255) & ~(size_t)255, sz_vs = (slab * vn * sizeof(Fr) + 255) & ~(size_t)255;
  const size_t sz_vp = (slab * vn * sizeof(Affine<Fq>) + 255) & ~(size_t)255, sz_sum = (slab * sizeof(XYZZ<Fq>) + 255) & ~(size_t)255;
  const size_t sz_v = (slab + 255) & ~(size_t)255, sz_m = slab * (size_t)vn * TBL_DIGITS * sizeof(XYZZ<Fq>);
  if ((rc = ctx->msm_b.reserve(sz_fs + sz_vs + sz_vp + sz_sum + sz_v + sz_m + 256))) return rc;
  uint8_t* base = (uint8_t*)ctx->msm_b.p;
  Fr* d_fs = (Fr*)base; base += sz_fs;
  Fr* d_vs = (Fr*)base; base += sz_vs;
  Affine<Fq>* d_vp = (Affine<Fq>*)base; base += sz_vp;
  XYZZ<Fq>* d_sum = (XYZZ<Fq>*)base; base += sz_sum;
  uint8_t* d_v = base; base += sz_v;
  XYZZ<Fq>* d_m = (XYZZ<Fq>*)base;
  const int mb = Curve::MODBYTES;
  for (size_t lo = 0; lo < batch; lo += slab) {
    const size_t cnt = batch - lo < slab ? batch - lo : slab;
    // canonical (non-Montgomery) scalars: the kernels only take digits
    if ((rc = scalars_from_host<Curve>(ctx, fixed_scalars_be + lo * F * mb, cnt * F, 0, d_fs))) return rc;
    if (vn) {
      if ((rc = scalars_from_host<Curve>(ctx, var_scalars_be + lo * vn * mb, cnt * vn, 0, d_vs))) return rc;
      if ((rc = points_from_host<Curve>(ctx, var_points_xy + lo * vn * 2 * mb, cnt * vn, d_vp))) return rc;
    }
    k_batch_fixed<Curve><<<(unsigned)cnt, 256, 0, ctx->stream>>>(runs, F, d_fs, d_sum);
    k_batch_var<Curve><<<(unsigned)cnt, 64, 0, ctx->stream>>>(vn, d_vp, d_vs, d_m, d_sum, d_v);
    ctx->launches += 2;
    if ((rc = launch_check(ctx, "batch_identity"))) return rc;
    BP_CUDA_OK(cudaMemcpyAsync(is_identity + lo, d_v, cnt, cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  }
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
  const size_t asz = ctx->curve == BPGPU_BLS12_381 ? sizeof(Affine<Bls::Fq>) : sizeof(Affine<Bn::Fq>);
  FixedRuns fr;
  fr.nruns = 0;
  uint32_t F = 0;
  for (size_t k = 0; k < nruns; k++) {
    const bpgpu_fixed_run& r = runs[k];
    if (r.n == 0) continue;
    const void* tbl = nullptr;
    if (r.points) {
      if (!(r.off <= r.points->n && r.n <= r.points->n - r.off)) return BPGPU_E_LEN;
      if (!r.points->table) return BPGPU_E_ARG;                       // bpgpu_points_precompute first
      tbl = (const uint8_t*)r.points->table + r.off * TBL_ENTRIES * asz;
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
    fr.start[fr.nruns] = F;
    fr.nruns++;
    F += (uint32_t)r.n;
  }
  fr.start[fr.nruns] = F;
  if (fr.nruns == 0) { fr.nruns = 1; fr.table[0] = nullptr; fr.start[0] = 0; fr.start[1] = 0; }
  return ctx->curve == BPGPU_BLS12_381
             ? batch_identity_t<Bls>(ctx, fr, F, batch, fixed_scalars_be, var_points_xy, var_scalars_be, (uint32_t)vn, is_identity)
             : batch_identity_t<Bn>(ctx, fr, F, batch, fixed_scalars_be, var_points_xy, var_scalars_be, (uint32_t)vn, is_identity);
}
