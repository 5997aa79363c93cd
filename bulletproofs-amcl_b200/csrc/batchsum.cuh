// Batched sums over window tables: R_b = sum_i f[b][i] * F_i for `rows` independent scalar rows over the SAME fixed
// points (concatenated runs of table rows).  Shared by the batched verification (batch.cu) and the lock-step batched
// prover (provebatch.cu).
#pragma once
#include "common.cuh"

namespace bp {

// 128 threads per row: 8 table lookups per (term, limb) item and a 7-level tree -- two to three blocks fit an SM, and a
// thread spends more of its time adding table entries than waiting in the tree (256 threads: one block per SM, 8 levels)
static const int BATCH_FIXED_THREADS = 128;
#ifndef BP_WARP_MADD
#define BP_WARP_MADD madd_c
#endif

struct FixedRuns {
  const void* table[TBL_MAX_SEGS];
  const void* table16[TBL_MAX_SEGS];      // wide (16-bit window) tables of the run, or null
  uint32_t start[TBL_MAX_SEGS + 1];
  int nruns;
};

// batch.cu: the verification MSMs of a slab of proofs from device-resident scalars and points (see batch.cu)
struct BatchScratch { size_t fs, vs, vp, sum, v, m, w, total; };      // byte offsets into one scratch block
template <class Curve> BatchScratch batch_scratch_layout(size_t slab, uint32_t F, uint32_t vn);
template <class Curve>
int batch_identity_launch(bpgpu_ctx* ctx, const FixedRuns& runs, uint32_t F, size_t cnt, const void* d_fs, const void* d_vp, const void* d_vs,
                          uint32_t vn, void* d_sum, void* d_m, void* d_w, uint8_t* d_v);
int resolve_fixed_runs(bpgpu_ctx* ctx, const bpgpu_fixed_run* runs, size_t nruns, FixedRuns* out, uint32_t* F_out);

// ipp_half != 0: the rows are the COMPACT L / R scalar rows of an inner-product round over [G[..N) | H[..N) | g] (N = F - 1):
// term p < N of an even row (L) multiplies H_p if p lies in the left half of its current block (p mod 2*ipp_half <
// ipp_half) and G_p otherwise, of an odd row (R) the other way round; term N multiplies g (ipp.rs:80-104, 148-170 with the
// folded generators expanded, see provebatch.cu).  Every term of such a row is non-zero: no lane idles on the bases a round
// does not touch.
template <class Curve>
__global__ void __launch_bounds__(BATCH_FIXED_THREADS, 3) k_batch_fixed(FixedRuns runs, uint32_t F, const typename Curve::Fr* __restrict__ scal, int mont,
                                                     XYZZ<typename Curve::Fq>* __restrict__ out, uint32_t ipp_half = 0) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  __shared__ __align__(16) unsigned char smraw[BATCH_FIXED_THREADS * sizeof(XYZZ<Fq>)];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(smraw);
  const size_t b = blockIdx.x;
  const Fr* sc = scal + b * F;
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  // long rows are cut into gridDim.y term ranges (one block each; the caller adds the gridDim.y sums of a row)
  const uint32_t per = (F + gridDim.y - 1) / gridDim.y;
  const uint32_t p_lo = blockIdx.y * per, p_hi = min(F, p_lo + per);
  for (uint32_t t = p_lo * 8 + threadIdx.x; t < p_hi * 8; t += blockDim.x) {
    const uint32_t p = t >> 3, j = t & 7;
    uint32_t limb;                                          // digits come from the canonical integer
    if (mont) { Fr sv = load_vec(sc + p); limb = sv.is_zero() ? 0u : sv.from_mont().v[j]; } else limb = sc[p].v[j];
    if (!limb) continue;
    uint32_t base = p;                                      // index into the concatenated runs
    if (ipp_half) {
      const uint32_t N = F - 1;
      if (p < N) {
        const bool left = (p & (2 * ipp_half - 1)) < ipp_half;
        base = (left != (bool)(b & 1)) ? N + p : p;         // L (even row): left -> H ; R (odd row): left -> G
      } else {
        base = 2 * N;
      }
    }
    int rg = 0;
    while (rg + 1 < runs.nruns && base >= runs.start[rg + 1]) rg++;
    const uint32_t row = base - runs.start[rg];
    const Affine<Fq>* tb = (const Affine<Fq>*)runs.table[rg] + ((size_t)row * TBL_WINDOWS + TBL_PER_LIMB * j) * TBL_DIGITS;
#pragma unroll 1
    for (int k = 0; k < TBL_PER_LIMB; k++) {
      const uint32_t d = (limb >> (TBL_BITS * k)) & (uint32_t)TBL_DIGITS;
      if (d) acc.madd(load_vec_ro(tb + k * TBL_DIGITS + (d - 1)));
    }
  }
  store_vec(sm + threadIdx.x, acc);
  __syncthreads();
  for (int o = BATCH_FIXED_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      XYZZ<Fq> a = load_vec(sm + threadIdx.x), c = load_vec(sm + threadIdx.x + o);
      a.add(c);
      store_vec(sm + threadIdx.x, a);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) store_vec(out + b * gridDim.y + blockIdx.y, load_vec(sm));
}

// The same sums with ONE WARP per row (4 rows per block) for launches with thousands of rows: a lane adds 4x as many table
// entries before a 5-level tree instead of a 7-level one, so the tree -- dependent full additions on half, a quarter, ...
// of the lanes -- shrinks from ~40 % of a row's chain to ~10 %.  No row splitting (gridDim.y = 1).
template <class Curve>
__global__ void __launch_bounds__(128, 3) k_batch_fixed_warp(FixedRuns runs, uint32_t F, uint32_t rows, const typename Curve::Fr* __restrict__ scal,
                                                             int mont, XYZZ<typename Curve::Fq>* __restrict__ out, uint32_t ipp_half = 0) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  __shared__ __align__(16) unsigned char smraw[128 * sizeof(XYZZ<Fq>)];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(smraw);
  const uint32_t b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  if (b < rows) {
    const Fr* sc = scal + (size_t)b * F;
    for (uint32_t t = lane; t < F * 8; t += 32) {
      const uint32_t p = t >> 3, j = t & 7;
      uint32_t limb;
      if (mont) { Fr sv = load_vec(sc + p); limb = sv.is_zero() ? 0u : sv.from_mont().v[j]; } else limb = sc[p].v[j];
      if (!limb) continue;
      uint32_t base = p;
      if (ipp_half) {
        const uint32_t N = F - 1;
        if (p < N) {
          const bool left = (p & (2 * ipp_half - 1)) < ipp_half;
          base = (left != (bool)(b & 1)) ? N + p : p;
        } else {
          base = 2 * N;
        }
      }
      int rg = 0;
      while (rg + 1 < runs.nruns && base >= runs.start[rg + 1]) rg++;
      const uint32_t row = base - runs.start[rg];
      // wide tables (two 16-bit windows per limb: half the additions) or the 8-bit ones (four): ONE loop and ONE inlined
      // mixed addition for both -- a second copy of its ~60 KB body makes the kernel miss the instruction cache
      const bool wide = runs.table16[rg] != nullptr;
      const Affine<Fq>* tb = wide ? (const Affine<Fq>*)runs.table16[rg] + ((size_t)row * TBL16_WINDOWS + 2 * j) * TBL16_DIGITS
                                  : (const Affine<Fq>*)runs.table[rg] + ((size_t)row * TBL_WINDOWS + TBL_PER_LIMB * j) * TBL_DIGITS;
      const uint32_t wbits = wide ? 16u : (uint32_t)TBL_BITS, wmask = wide ? 0xffffu : (uint32_t)TBL_DIGITS;
      const uint32_t wstride = wide ? (uint32_t)TBL16_DIGITS : (uint32_t)TBL_DIGITS;
      const int nwin = wide ? 2 : TBL_PER_LIMB;
#pragma unroll 1
      for (int k = 0; k < nwin; k++) {
        const uint32_t d = (limb >> (wbits * k)) & wmask;
        if (d) acc.BP_WARP_MADD(load_vec_ro(tb + (size_t)k * wstride + (d - 1)));
      }
    }
  }
  store_vec(sm + threadIdx.x, acc);
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) {
    if ((int)lane < o) {
      XYZZ<Fq> a = load_vec(sm + threadIdx.x), c = load_vec(sm + threadIdx.x + o);
      a.add(c);
      store_vec(sm + threadIdx.x, a);
    }
    __syncthreads();
  }
  if (lane == 0 && b < rows) store_vec(out + b, load_vec(sm + threadIdx.x));
}

// Pedersen commitments over the pair (g, h): out[r] = s[2r] * g + s[2r+1] * h, scalars CANONICAL integers.  Two terms do not
// fill a 128-thread row block: one warp per commitment, lane w adds the two entries of byte-window w, five tree levels.
template <class Curve>
__global__ void __launch_bounds__(128) k_commit_pair(uint32_t rows, const void* __restrict__ tg, const void* __restrict__ th,
                                                     const typename Curve::Fr* __restrict__ scal, XYZZ<typename Curve::Fq>* __restrict__ out) {
  using Fq = typename Curve::Fq;
  __shared__ __align__(16) unsigned char smraw[128 * sizeof(XYZZ<Fq>)];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(smraw);
  const uint32_t r = blockIdx.x * 4 + (threadIdx.x >> 5), w = threadIdx.x & 31;
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  if (r < rows) {
    const uint32_t dg = (scal[2 * (size_t)r].v[w / TBL_PER_LIMB] >> ((w % TBL_PER_LIMB) * TBL_BITS)) & (uint32_t)TBL_DIGITS;
    const uint32_t dh = (scal[2 * (size_t)r + 1].v[w / TBL_PER_LIMB] >> ((w % TBL_PER_LIMB) * TBL_BITS)) & (uint32_t)TBL_DIGITS;
    if (dg) acc.madd(load_vec_ro((const Affine<Fq>*)tg + (size_t)w * TBL_DIGITS + (dg - 1)));
    if (dh) acc.madd(load_vec_ro((const Affine<Fq>*)th + (size_t)w * TBL_DIGITS + (dh - 1)));
  }
  store_vec(sm + threadIdx.x, acc);
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) {
    if ((int)w < o) {
      XYZZ<Fq> a = load_vec(sm + threadIdx.x), c = load_vec(sm + threadIdx.x + o);
      a.add(c);
      store_vec(sm + threadIdx.x, a);
    }
    __syncthreads();
  }
  if (w == 0 && r < rows) store_vec(out + r, load_vec(sm + threadIdx.x));
}

// out[r] = sum of the `splits` partial sums of row r (in place at parts[r * splits])
template <class Fq>
__global__ void __launch_bounds__(64) k_batch_fixed_combine(uint32_t rows, uint32_t splits, XYZZ<Fq>* __restrict__ parts, XYZZ<Fq>* __restrict__ out) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  XYZZ<Fq> acc = load_vec(parts + (size_t)r * splits);
  for (uint32_t k = 1; k < splits; k++) { XYZZ<Fq> q = load_vec(parts + (size_t)r * splits + k); acc.add(q); }
  store_vec(out + r, acc);
}

}  // namespace bp
