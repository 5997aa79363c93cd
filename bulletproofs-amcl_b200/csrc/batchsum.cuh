// Batched sums over window tables: R_b = sum_i f[b][i] * F_i for `rows` independent scalar rows over the SAME fixed
// points (concatenated runs of table rows).  Shared by the batched verification (batch.cu) and the lock-step batched
// prover (provebatch.cu).
#pragma once
#include "common.cuh"

namespace bp {

struct FixedRuns {
  const void* table[TBL_MAX_SEGS];
  uint32_t start[TBL_MAX_SEGS + 1];
  int nruns;
};

template <class Curve>
__global__ void __launch_bounds__(256) k_batch_fixed(FixedRuns runs, uint32_t F, const typename Curve::Fr* __restrict__ scal, int mont,
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
    uint32_t limb;                                          // digits come from the canonical integer
    if (mont) { Fr sv = load_vec(sc + p); limb = sv.is_zero() ? 0u : sv.from_mont().v[j]; } else limb = sc[p].v[j];
    if (!limb) continue;
    const Affine<Fq>* tb = (const Affine<Fq>*)runs.table[rg] + ((size_t)row * TBL_WINDOWS + TBL_PER_LIMB * j) * TBL_DIGITS;
#pragma unroll 1
    for (int k = 0; k < TBL_PER_LIMB; k++) {
      const uint32_t d = (limb >> (TBL_BITS * k)) & (uint32_t)TBL_DIGITS;
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


}  // namespace bp
