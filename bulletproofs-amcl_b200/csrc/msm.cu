// G1 multi-scalar multiplication for sm_100a: signed-digit Pippenger with a balanced,
// sort-based bucket accumulation.
//
// Replaces amcl_wrapper's Straus/wNAF `multi_scalar_mul_var_time`,
// `inner_product_var_time_with_ref_vecs` and the fixed-window `inner_product_const_time`
// (call sites /root/reference/src/ipp.rs:91,104,158,170,251-253; src/r1cs/verifier.rs:451;
// src/r1cs/prover.rs:347-362).  Output is the normalised affine point, which is unique, so the
// different summation order changes no observable byte.
//
// Pipeline (all on one stream, inputs and every intermediate resident in HBM):
//   k_digits        scalars -> W signed c-bit digits each, histogram per (window, bucket)
//   k_scan          per window: bucket start offsets + offsets of the per-chunk partial sums
//   k_scatter       counting-sort scatter of (point index | sign) by bucket
//   k_chunk_acc     ONE THREAD PER FIXED-SIZE CHUNK of the sorted list: S mixed XYZZ additions
//                   each, whatever the bucket sizes are -> no divergence between lanes and no
//                   serialisation on a hot bucket (0/1 witness scalars put half of all points
//                   in one bucket: helper_constraints/positive_no.rs:18-24)
//   k_merge         one thread per chunk boundary a bucket straddles: its partial sums folded into one; the extra blocks of
//                   the same launch collapse buckets spanning > T chunks with a block-wide tree (giant_blocks)
//   k_reduce_l1     one WARP per 32*L1 buckets: lanes run top-down running sums over their buckets
//                   (two additions per bucket), then a warp suffix-scan turns them into sum (b-base)*B_b
//   k_reduce_l2     per window: combine the segment results (two warps)
//   (host)          Horner over the W window sums + affine normalisation: a ~256-doubling
//                   dependent chain, ~0.1 ms on one CPU core vs milliseconds on one GPU thread
#include <stdlib.h>

#include "common.cuh"

namespace bp {

struct MsmGeom {
  uint32_t n;       // number of terms
  int c;            // window bits
  int W;            // windows in all (W0 per scalar set)
  int W0;           // windows of ONE scalar: several scalar sets over the same points are extra windows (msm_run nsets)
  uint32_t nbp;     // bucket slots per window = 2^(c-1) + 1 (slot 0 unused)
  uint32_t S;       // sorted entries per chunk thread
  uint32_t nchunk;  // chunk threads per window = ceil(n / S)
  uint32_t pcap;    // partial-sum slots per window = nchunk + nbp
  int lgL1;         // log2(buckets per lane) in k_reduce_l1; a warp covers 32 << lgL1 buckets
  uint32_t nseg;    // level-1 warps (segments) per window
  int lgL2;         // log2(segments per lane) in k_reduce_l2
};

static const int GIANT_T = 16;         // buckets with more partials than this are collapsed by giant_blocks (a block tree: ~9 dependent additions)
static const int GIANT_BLOCK = 128;

int msm_window_bits(size_t n) {
  if (n == 0) return 4;
  double best = 1e300;
  int bc = 4;
  for (int c = 3; c <= 18; c++) {
    int W = (256 + c - 1) / c;
    double cost = (double)W * ((double)n * 10.0 + (double)(1u << (c - 1)) * 60.0);
    if (cost < best) { best = cost; bc = c; }
  }
  return bc;
}

// ------------------------------------------------------------------------------------------
template <class Fr>
__global__ void __launch_bounds__(256) k_digits(const Fr* __restrict__ scal0, const Fr* __restrict__ scal1, int mont, MsmGeom g,
                                                uint32_t* __restrict__ digits, uint32_t* __restrict__ hist) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const int set = blockIdx.y;                   // scalar set: its digits are windows [set * W0, (set + 1) * W0)
  const Fr* scal = set ? scal1 : scal0;
  digits += (size_t)set * g.W0 * g.n;
  if (hist) hist += (size_t)set * g.W0 * g.nbp;
  Fr s = load_vec(scal + i);
  if (mont) s = s.from_mont();
  uint32_t limbs[9];
#pragma unroll
  for (int k = 0; k < 8; k++) limbs[k] = s.v[k];
  limbs[8] = 0;
  const uint32_t half = 1u << (g.c - 1);
  const uint32_t mask = (1u << g.c) - 1;
  uint32_t carry = 0;
  for (int w = 0; w < g.W0; w++) {
    uint32_t bit = (uint32_t)w * g.c;
    uint32_t limb = bit >> 5, sh = bit & 31;
    uint32_t d = 0;
    if (limb < 8) {
      uint64_t two = (uint64_t)limbs[limb] | ((uint64_t)limbs[limb + 1] << 32);
      d = (uint32_t)(two >> sh) & mask;
    }
    d += carry;
    uint32_t mag, sign;
    if (d > half) { mag = (mask + 1) - d; sign = 0x80000000u; carry = 1; }
    else { mag = d; sign = 0; carry = 0; }
    digits[(size_t)w * g.n + i] = mag ? (mag | sign) : 0u;
    if (mag && hist) atomicAdd(&hist[(size_t)w * g.nbp + mag], 1u);
  }
}

__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t* total_out) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t wv = warp_tot[lane];
    uint32_t winc = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_tot[lane] = winc - wv;
    if (lane == 31) total = winc;
  }
  __syncthreads();
  uint32_t res = warp_tot[wid] + inc - v;
  *total_out = total;
  __syncthreads();
  return res;
}

// one block of 1024 threads per window
__global__ void __launch_bounds__(1024) k_scan(MsmGeom g, uint32_t* __restrict__ hist, uint32_t* __restrict__ bstart,
                                               uint32_t* __restrict__ cursor, uint32_t* __restrict__ pstart,
                                               uint32_t* __restrict__ giant_count, uint32_t* __restrict__ giant_list) {
  const int w = blockIdx.x;
  uint32_t* h = hist + (size_t)w * g.nbp;
  uint32_t* bs = bstart + (size_t)w * (g.nbp + 1);
  uint32_t* cu = cursor + (size_t)w * g.nbp;
  uint32_t* ps = pstart + (size_t)w * (g.nbp + 1);
  const uint32_t per = (g.nbp + 1023) / 1024;
  const uint32_t lo = threadIdx.x * per;
  const uint32_t hi = min(lo + per, g.nbp);
  uint32_t sum = 0;
  for (uint32_t b = lo; b < hi; b++) sum += h[b];
  uint32_t total;
  uint32_t run = block_excl_scan_1024(sum, &total);
  uint32_t npsum = 0;
  for (uint32_t b = lo; b < hi; b++) {
    uint32_t cnt = h[b];
    bs[b] = run;
    cu[b] = run;
    uint32_t np = cnt ? ((run + cnt - 1) / g.S - run / g.S + 1) : 0;
    h[b] = np;                       // the histogram slot now holds the partial count
    if (np > (uint32_t)GIANT_T) giant_list[atomicAdd(giant_count, 1u)] = (uint32_t)w * g.nbp + b;
    npsum += np;
    run += cnt;
  }
  if (threadIdx.x == 0) bs[g.nbp] = total;
  uint32_t total2;
  uint32_t run2 = block_excl_scan_1024(npsum, &total2);
  for (uint32_t b = lo; b < hi; b++) { ps[b] = run2; run2 += h[b]; }
  if (threadIdx.x == 0) ps[g.nbp] = total2;
}

__global__ void __launch_bounds__(256) k_scatter(MsmGeom g, const uint32_t* __restrict__ digits, uint32_t* __restrict__ cursor,
                                                 uint32_t* __restrict__ sidx) {
  const size_t total = (size_t)g.W * g.n;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    uint32_t d = digits[t];
    if (!d) continue;
    uint32_t w = (uint32_t)(t / g.n), i = (uint32_t)(t - (size_t)w * g.n);
    uint32_t mag = d & 0x7fffffffu;
    uint32_t pos = atomicAdd(&cursor[(size_t)w * g.nbp + mag], 1u);
    sidx[(size_t)w * g.n + pos] = i | (d & 0x80000000u);      // the bucket is implied by the position (bstart)
  }
}

// bucket of sorted entry e: the largest b with bstart[b] <= e (empty buckets share their successor's start, so the
// largest such index is the non-empty bucket that holds e); bs has nbp + 1 entries, bs[nbp] = number of entries > e
__device__ __forceinline__ uint32_t bucket_of(const uint32_t* __restrict__ bs, uint32_t nbp, uint32_t e) {
  uint32_t lo = 0, hi = nbp;                    // invariant: bs[lo] <= e < bs[hi]
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (bs[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

// The inlined form needs 194 registers (2 blocks of 128 threads per SM; forcing it into 168 measured the same in round 1).
// COMPACT: the mixed addition as real calls (XYZZ::madd_c, 128 registers = 4 blocks per SM, the loop stays in the instruction
// cache): what a machine-filling launch wants (2^20 terms: 6.43 -> 5.80 ms, 0.955 of the multiplier pipe).  A launch that does not fill the
// machine is bound by the latency of one thread's chain, where the inlined form (independent products overlap) is faster
// (2^16 terms: 1.44 against 1.73 ms), so the host picks per launch.
template <class Fq, bool COMPACT>
__global__ void __launch_bounds__(128, COMPACT ? 4 : 2) k_chunk_acc(MsmGeom g, const Affine<Fq>* __restrict__ pts,
                                                   const uint32_t* __restrict__ sidx,
                                                   const uint32_t* __restrict__ bstart, const uint32_t* __restrict__ pstart,
                                                   XYZZ<Fq>* __restrict__ partials) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (uint32_t)g.W * g.nchunk) return;
  const uint32_t w = gid / g.nchunk, t = gid - w * g.nchunk;
  const uint32_t* bs = bstart + (size_t)w * (g.nbp + 1);
  const uint32_t* ps = pstart + (size_t)w * (g.nbp + 1);
  const uint32_t cnt = bs[g.nbp];
  const uint32_t e0 = t * g.S;
  if (e0 >= cnt) return;
  const uint32_t e1 = min(e0 + g.S, cnt);
  const uint32_t* idx = sidx + (size_t)w * g.n;
  XYZZ<Fq>* out = partials + (size_t)w * g.pcap;
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  uint32_t cur = bucket_of(bs, g.nbp, e0);
  uint32_t end = bs[cur + 1];                   // first entry past bucket `cur`
  for (uint32_t e = e0; e < e1; e++) {
    if (e >= end) {                             // next bucket: almost always cur + 1, else search past the empty ones
      store_vec(out + ps[cur] + (t - bs[cur] / g.S), acc);
      acc = XYZZ<Fq>::inf();
      cur++;
      end = bs[cur + 1];
      if (e >= end) { cur = bucket_of(bs, g.nbp, e); end = bs[cur + 1]; }
    }
    uint32_t id = idx[e];
    Affine<Fq> P = load_vec_ro(pts + (id & 0x7fffffffu));
    if (id >> 31) P.y = P.y.neg();
    if (COMPACT) acc.madd_c(P); else acc.madd(P);
  }
  store_vec(out + ps[cur] + (t - bs[cur] / g.S), acc);
}

template <class Fq>
__device__ __forceinline__ XYZZ<Fq> block_tree_sum(XYZZ<Fq> v, XYZZ<Fq>* sm) {
  store_vec(sm + threadIdx.x, v);
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      XYZZ<Fq> a = load_vec(sm + threadIdx.x), b = load_vec(sm + threadIdx.x + o);
      a.add(b);
      store_vec(sm + threadIdx.x, a);
    }
    __syncthreads();
  }
  XYZZ<Fq> r = load_vec(sm);
  __syncthreads();
  return r;
}

// Buckets whose points span more than GIANT_T chunks (skewed scalars: 0/1 witnesses, equal
// scalars) are collapsed by a whole block: the total goes to the bucket's first partial slot and
// the bucket's partial count becomes 1, so k_reduce_l1 reads at most GIANT_T partials per bucket.
template <class Fq>
__device__ __forceinline__ void giant_blocks(uint32_t first, uint32_t stride, MsmGeom g, const uint32_t* __restrict__ pstart,
                                             uint32_t* __restrict__ pcount, XYZZ<Fq>* __restrict__ partials,
                                             const uint32_t* __restrict__ giant_count, const uint32_t* __restrict__ giant_list) {
  __shared__ __align__(16) unsigned char smraw[GIANT_BLOCK * sizeof(XYZZ<Fq>)];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(smraw);
  const uint32_t ng = *giant_count;
  for (uint32_t gi = first; gi < ng; gi += stride) {
    const uint32_t gid = giant_list[gi];
    const uint32_t w = gid / g.nbp, b = gid - w * g.nbp;
    const uint32_t* ps = pstart + (size_t)w * (g.nbp + 1);
    const uint32_t p0 = ps[b], p1 = ps[b + 1];
    XYZZ<Fq>* in = partials + (size_t)w * g.pcap;
    XYZZ<Fq> acc = XYZZ<Fq>::inf();
    for (uint32_t k = p0 + threadIdx.x; k < p1; k += blockDim.x) { XYZZ<Fq> q = load_vec(in + k); acc.add(q); }
    XYZZ<Fq> tot = block_tree_sum(acc, sm);
    if (threadIdx.x == 0) { store_vec(in + p0, tot); pcount[gid] = 1; }
    __syncthreads();
  }
}

// Buckets that straddle chunk boundaries have several partial sums (one per chunk they touch).  Phase A of the bucket
// reduction folds them into ONE sum per bucket, in place at the bucket's first partial slot, with full parallelism: the
// thread of the chunk in which the bucket STARTS does it, and only if the bucket continues into the next chunk (so a
// thread has at most one bucket to fold, and -- the chunk length being about two buckets -- almost always exactly one
// addition: no divergence, W * n / S independent threads).  The > GIANT_T cases belong to the giant blocks.
// Afterwards pcount[b] is 0 or 1 for every bucket and k_reduce_l1 is two additions per bucket with no inner loop.
// ONE launch does both: blocks [0, merge_blocks) fold the straddling buckets, the blocks after them collapse the giant
// buckets (the two sets of buckets are disjoint: a merging thread leaves a bucket with more than GIANT_T partials alone
// whether or not its giant block has already finished), so a small MSM does not wait for a block tree before it merges.
template <class Fq>
__global__ void __launch_bounds__(128) k_merge(MsmGeom g, uint32_t merge_blocks, const uint32_t* __restrict__ bstart,
                                               const uint32_t* __restrict__ pstart, uint32_t* __restrict__ pcount,
                                               XYZZ<Fq>* __restrict__ partials, const uint32_t* __restrict__ giant_count,
                                               const uint32_t* __restrict__ giant_list) {
  static_assert(GIANT_BLOCK == 128, "one block shape for both roles");
  if (blockIdx.x >= merge_blocks) {
    giant_blocks<Fq>(blockIdx.x - merge_blocks, gridDim.x - merge_blocks, g, pstart, pcount, partials, giant_count, giant_list);
    return;
  }
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (uint32_t)g.W * g.nchunk) return;
  const uint32_t w = gid / g.nchunk, t = gid - w * g.nchunk;
  const uint32_t* bs = bstart + (size_t)w * (g.nbp + 1);
  const uint32_t cnt = bs[g.nbp];
  const uint64_t e1 = (uint64_t)(t + 1) * g.S;              // first entry of the next chunk
  if (e1 >= cnt) return;
  const uint32_t b = bucket_of(bs, g.nbp, (uint32_t)(e1 - 1));
  if (bs[b + 1] <= e1 || bs[b] / g.S != t) return;          // no straddle, or the bucket began in an earlier chunk
  uint32_t* pc = pcount + (size_t)w * g.nbp;
  const uint32_t np = pc[b];
  if (np <= 1 || np > (uint32_t)GIANT_T) return;            // nothing to fold, or a giant block's bucket (collapsed there, maybe already)
  XYZZ<Fq>* in = partials + (size_t)w * g.pcap + pstart[(size_t)w * (g.nbp + 1) + b];
  XYZZ<Fq> acc = load_vec(in);
  for (uint32_t k = 1; k < np; k++) { XYZZ<Fq> q = load_vec(in + k); acc.add(q); }   // one addition per thread: compact code, see Fp::mulc
  store_vec(in, acc);
  pc[b] = 1;
}

template <class Fq>
__device__ __forceinline__ XYZZ<Fq> shfl_down_xyzz(const XYZZ<Fq>& v, int o) {
  XYZZ<Fq> r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(XYZZ<Fq>) / 4); i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], o);
  return r;
}

// Warp-wide  sum_l ( acc_l + (l * 2^lgL) * run_l )  and  sum_l run_l, results valid on lane 0.
// Uses  sum_l l*run_l = sum_{j>=1} suffix_j  with suffix_j = sum_{l>=j} run_l  (warp suffix scan).
template <class Fq>
__device__ __forceinline__ void warp_weighted_sum(XYZZ<Fq>& acc, XYZZ<Fq>& run, int lgL) {
  const int lane = threadIdx.x & 31;
  XYZZ<Fq> suf = run;
#pragma unroll 1
  for (int o = 1; o < 32; o <<= 1) {
    XYZZ<Fq> t = shfl_down_xyzz(suf, o);
    if (lane + o < 32) suf.add(t);
  }
  run = suf;                                   // lane 0: total
  if (lane >= 1) {
#pragma unroll 1
    for (int k = 0; k < lgL; k++) suf.dbl();
    acc.add(suf);
  }
#pragma unroll 1
  for (int o = 16; o > 0; o >>= 1) {
    XYZZ<Fq> t = shfl_down_xyzz(acc, o);
    if (lane < o) acc.add(t);
  }
}

// Level 1 of the bucket reduction: one warp per 32*L1 consecutive buckets of one window.
// Lane l sums the chunk partials of its L1 buckets (top down) keeping run = sum B_b and
// acc = sum (b - lo + 1) * B_b; the warp then combines lanes.  Output per segment:
//   segA = sum (b - segbase) * B_b ,  segS = sum B_b      (segbase = first bucket - 1)
template <class Fq>
__global__ void __launch_bounds__(128) k_reduce_l1(MsmGeom g, const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ pcount,
                                                   const XYZZ<Fq>* __restrict__ partials, XYZZ<Fq>* __restrict__ segA,
                                                   XYZZ<Fq>* __restrict__ segS) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (uint32_t)g.W * g.nseg) return;
  const uint32_t w = warp / g.nseg, seg = warp - w * g.nseg;
  const uint32_t L1 = 1u << g.lgL1;
  const uint32_t* ps = pstart + (size_t)w * (g.nbp + 1);
  const uint32_t* pc = pcount + (size_t)w * g.nbp;
  const XYZZ<Fq>* in = partials + (size_t)w * g.pcap;
  const uint32_t lo = seg * 32 * L1 + (uint32_t)lane * L1 + 1;
  XYZZ<Fq> run = XYZZ<Fq>::inf(), acc = XYZZ<Fq>::inf();
  if (lo <= g.nbp - 1) {
    const uint32_t hi = min(lo + L1 - 1, g.nbp - 1);
    for (uint32_t b = hi; b >= lo; b--) {
      if (pc[b]) { XYZZ<Fq> q = load_vec(in + ps[b]); run.add(q); }      // k_merge left ONE sum per bucket
      acc.add(run);
    }
  }
  warp_weighted_sum(acc, run, g.lgL1);
  // callers with ONE segment per window pass segA = the window sums P_w and segS = Q_w (weight 0: the identity): no level 2
  if (lane == 0) { store_vec(segA + warp, acc); store_vec(segS + warp, g.nseg == 1 ? XYZZ<Fq>::inf() : run); }
}


// Level 2: per window, combine the nseg segment results (one block of 64 threads per window):
//   P_w = sum_seg segA ,  Q_w = sum_seg seg * segS ;  window sum = P_w + 2^(5+lgL1) * Q_w
// warp 0 computes Q_w, warp 1 computes P_w.  The final shift-and-add is left to the host finish.
template <class Fq>
__global__ void __launch_bounds__(64) k_reduce_l2(MsmGeom g, const XYZZ<Fq>* __restrict__ segA, const XYZZ<Fq>* __restrict__ segS,
                                                  XYZZ<Fq>* __restrict__ winP, XYZZ<Fq>* __restrict__ winQ) {
  const uint32_t w = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t L2 = 1u << g.lgL2;
  const uint32_t lo = (uint32_t)lane * L2;
  if (wid == 0) {
    // weights (seg + 1) via the generic routine, then subtract the plain total once
    XYZZ<Fq> run = XYZZ<Fq>::inf(), acc = XYZZ<Fq>::inf();
    if (lo < g.nseg) {
      const uint32_t hi = min(lo + L2, g.nseg);
      for (uint32_t s = hi; s-- > lo;) { XYZZ<Fq> q = load_vec(segS + (size_t)w * g.nseg + s); run.add(q); acc.add(run); }
    }
    warp_weighted_sum(acc, run, g.lgL2);
    if (lane == 0) { XYZZ<Fq> neg = run.neg(); acc.add(neg); store_vec(winQ + w, acc); }
  } else {
    XYZZ<Fq> acc = XYZZ<Fq>::inf();
    if (lo < g.nseg) {
      const uint32_t hi = min(lo + L2, g.nseg);
      for (uint32_t s = lo; s < hi; s++) { XYZZ<Fq> q = load_vec(segA + (size_t)w * g.nseg + s); acc.add(q); }
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
      XYZZ<Fq> t = shfl_down_xyzz(acc, o);
      if (lane < o) acc.add(t);
    }
    if (lane == 0) store_vec(winP + w, acc);
  }
}

// ------------------------------------------------------------------------------------------
// SMALL MSMs (n <= SMALL_MAX_N): one block per window does everything after the digits -- counting sort in shared memory,
// balanced accumulation, folding of the buckets that straddle threads, bucket reduction -- in ONE launch.  The general
// pipeline above is six launches of latency-bound kernels for such an input (0.27 ms whatever n is); the verification MSM
// of a proof, the IPP rounds after the generators were materialised and every MSM below a few thousand terms are this shape.
//   phase 1  histogram of the window's digits, exclusive scan (one warp), scatter of (index | sign) into shared memory
//   phase 2  thread t accumulates sorted entries [t S, (t+1) S), S >= ceil(entries / threads): a run that covers its whole
//            bucket goes to bsum[b]; a run of a bucket that began in an earlier thread goes to head[t], one that continues
//            in the next thread to tail[t] (at most one of each per thread)
//   phase 3  thread b folds bucket b = tail[t0] + head[t0+1 .. t1] when it spans few threads; wider ones (skewed digits,
//            the short top window) are summed by a warp each
//   phase 4  warp 0: lanes run top-down sums over their buckets, warp suffix scan (as k_reduce_l1 with one segment)
static const uint32_t SMALL_MAX_N = 4096;
static const int SMALL_T = 256;
static const uint32_t SMALL_SPAN = 6;

template <class Fq>
__global__ void __launch_bounds__(SMALL_T) k_msm_small(MsmGeom g, const Affine<Fq>* __restrict__ pts, const uint32_t* __restrict__ digits,
                                                       XYZZ<Fq>* __restrict__ scratch, XYZZ<Fq>* __restrict__ winP, XYZZ<Fq>* __restrict__ winQ) {
  extern __shared__ uint32_t sm32[];
  const uint32_t nb = g.nbp - 1, n = g.n, w = blockIdx.x, tid = threadIdx.x, T = blockDim.x;      // T = 64, 128 or 256 >= nb
  const int lane = tid & 31, wid = tid >> 5;
  uint32_t* cnt = sm32;                 // [nb + 2]  histogram, then the scatter cursors
  uint32_t* bs = cnt + nb + 2;          // [nb + 2]  bs[b] = first sorted entry of bucket b (1..nb); bs[nb + 1] = entries
  uint32_t* wide = bs + nb + 2;         // [nb + 1]  [0] = count, then the buckets left to the warps
  uint32_t* ent = wide + nb + 1;        // [n]       (index | sign) sorted by bucket
  const uint32_t* dg = digits + (size_t)w * n;
  for (uint32_t b = tid; b < nb + 2; b += T) cnt[b] = 0;
  if (tid == 0) wide[0] = 0;
  __syncthreads();
  for (uint32_t i = tid; i < n; i += T) {
    const uint32_t d = dg[i];
    if (d) atomicAdd(&cnt[d & 0x7fffffffu], 1u);
  }
  __syncthreads();
  if (tid < 32) {
    const uint32_t K = (nb + 31) / 32, lo = 1 + tid * K;
    uint32_t sum = 0;
    for (uint32_t k = 0; k < K; k++) if (lo + k <= nb) sum += cnt[lo + k];
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    uint32_t run = inc - sum;
    for (uint32_t k = 0; k < K; k++) if (lo + k <= nb) { bs[lo + k] = run; run += cnt[lo + k]; }
    if (tid == 31) bs[nb + 1] = inc;
  }
  __syncthreads();
  for (uint32_t b = 1 + tid; b <= nb; b += T) cnt[b] = bs[b];
  __syncthreads();
  for (uint32_t i = tid; i < n; i += T) {
    const uint32_t d = dg[i];
    if (d) ent[atomicAdd(&cnt[d & 0x7fffffffu], 1u)] = i | (d & 0x80000000u);
  }
  __syncthreads();
  const uint32_t total = bs[nb + 1];
  XYZZ<Fq>* head = scratch + (size_t)w * (2 * SMALL_T + g.nbp);
  XYZZ<Fq>* tail = head + SMALL_T;
  XYZZ<Fq>* bsum = tail + SMALL_T;      // [nbp], indexed by bucket
  // chunk length: a thread's chain is S mixed additions, then the bucket's thread folds ~avg / S partial sums (full additions,
  // 1.4 x the cost): S ~ sqrt(1.4 avg) minimises the sum; never fewer entries per thread than 256 threads need to cover them
  uint32_t S = (total + T - 1) / T;
  {
    const uint32_t avg14 = (total * 14u) / (10u * (nb ? nb : 1u));          // 1.4 * entries per bucket
    uint32_t r = 1;
    while ((r + 1) * (r + 1) <= avg14) r++;
    if (S < r) S = r;
    if (S == 0) S = 1;
  }
  {
    const uint32_t e0 = tid * S, e1 = min(e0 + S, total);
    if (e0 < e1) {
      uint32_t lo = 1, hi = nb + 1;                   // bs[lo] <= e0 < bs[hi]: the largest such lo is the bucket of e0
      while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (bs[mid] <= e0) lo = mid; else hi = mid; }
      uint32_t cur = lo, end = bs[cur + 1];
      XYZZ<Fq> acc = XYZZ<Fq>::inf();
      for (uint32_t e = e0; e <= e1; e++) {
        if (e == e1 || e >= end) {                    // the run of bucket `cur` is over: where does it go?
          if (bs[cur] < e0) store_vec(head + tid, acc);
          else if (end > e1) store_vec(tail + tid, acc);
          else store_vec(bsum + cur, acc);
          if (e == e1) break;
          acc = XYZZ<Fq>::inf();
          do { cur++; end = bs[cur + 1]; } while (e >= end);
        }
        const uint32_t id = ent[e];
        Affine<Fq> P = load_vec_ro(pts + (id & 0x7fffffffu));
        if (id >> 31) P.y = P.y.neg();
        acc.madd(P);
      }
    }
  }
  __syncthreads();
  if (tid >= 1 && tid <= nb) {
    const uint32_t b = tid, s0 = bs[b], s1 = bs[b + 1];
    if (s1 > s0) {
      const uint32_t t0 = s0 / S, t1 = (s1 - 1) / S;
      if (t1 - t0 > SMALL_SPAN) wide[1 + atomicAdd(&wide[0], 1u)] = b;
      else if (t1 > t0) {
        XYZZ<Fq> acc = load_vec(tail + t0);
        for (uint32_t t = t0 + 1; t <= t1; t++) { XYZZ<Fq> q = load_vec(head + t); acc.add(q); }
        store_vec(bsum + b, acc);
      }
    }
  }
  __syncthreads();
  for (uint32_t gi = wid; gi < wide[0]; gi += T / 32) {
    const uint32_t b = wide[1 + gi], s0 = bs[b], s1 = bs[b + 1];
    const uint32_t t0 = s0 / S, t1 = (s1 - 1) / S;
    XYZZ<Fq> acc = XYZZ<Fq>::inf();
    if (lane == 0) acc = load_vec(tail + t0);
    for (uint32_t t = t0 + 1 + lane; t <= t1; t += 32) { XYZZ<Fq> q = load_vec(head + t); acc.add(q); }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
      XYZZ<Fq> t = shfl_down_xyzz(acc, o);
      if (lane < o) acc.add(t);
    }
    if (lane == 0) store_vec(bsum + b, acc);
  }
  __syncthreads();
  if (wid == 0) {
    int lgL = 0;
    while ((32u << lgL) < nb) lgL++;
    const uint32_t L1 = 1u << lgL, lo = (uint32_t)lane * L1 + 1;
    XYZZ<Fq> run = XYZZ<Fq>::inf(), acc = XYZZ<Fq>::inf();
    if (lo <= nb) {
      const uint32_t hi = min(lo + L1 - 1, nb);
      for (uint32_t b = hi; b >= lo; b--) {
        if (bs[b + 1] > bs[b]) { XYZZ<Fq> q = load_vec(bsum + b); run.add(q); }
        acc.add(run);
      }
    }
    warp_weighted_sum(acc, run, lgL);
    if (lane == 0) { store_vec(winP + w, acc); store_vec(winQ + w, XYZZ<Fq>::inf()); }
  }
}

// ------------------------------------------------------------------------------------------
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct StageTimer {
  bool on;
  cudaStream_t st;
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> names;
  bool verbose;
  StageTimer(cudaStream_t s, bool enable) : on(enable || getenv("BPGPU_PROFILE") != nullptr), st(s), verbose(getenv("BPGPU_PROFILE") != nullptr) { mark("start"); }
  void mark(const char* name) {
    if (!on) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st);
    ev.push_back(e); names.push_back(name);
  }
  void report(const MsmGeom& g, bpgpu_ctx* ctx) {
    if (!on) return;
    cudaEventSynchronize(ev.back());
    for (size_t i = 1; i < ev.size() && i <= 8; i++) { float ms; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]); ctx->stage_ms_sum[i - 1] += ms; }
    ctx->stage_runs++;
    if (!verbose) { for (auto e : ev) cudaEventDestroy(e); return; }
    fprintf(stderr, "[bpgpu msm n=%u c=%d W=%d S=%u lgL1=%d nseg=%u]", g.n, g.c, g.W, g.S, g.lgL1, g.nseg);
    for (size_t i = 1; i < ev.size(); i++) { float ms; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]); fprintf(stderr, " %s=%.3f", names[i], ms); }
    float tot; cudaEventElapsedTime(&tot, ev.front(), ev.back());
    fprintf(stderr, " total=%.3fms\n", tot);
    for (auto e : ev) cudaEventDestroy(e);
  }
};

// Leaves the W per-window sums (XYZZ, Montgomery form) in ctx scratch; the caller combines them
// (Horner over windows + affine normalisation) on the host, see msm_finish_host in api.cu.
// d_scalars2 (optional): a SECOND scalar vector over the same points (the L and R of an IPP round): its windows are
// appended to the first one's, every later stage just sees 2 * W0 windows, and one pipeline run yields both sums.
template <class Curve>
int msm_run(bpgpu_ctx* ctx, const Affine<typename Curve::Fq>* d_points, const void* d_scalars, bool scalars_mont, size_t n,
            MsmResult* res, const void* d_scalars2) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  cudaStream_t st = ctx->stream;
  res->W = 0; res->c = 0; res->qshift = 0; res->d_winsum = nullptr; res->nsets = d_scalars2 ? 2 : 1;
  if (n == 0) return BPGPU_OK;
  if (n >= (1ull << 31)) return BPGPU_E_ARG;
  MsmGeom g;
  g.n = (uint32_t)n;
  g.c = msm_window_bits(n);
  // tuning overrides for tools/msm_sweep.py (unset in normal use)
  static const char* env_c = getenv("BPGPU_C");
  static const char* env_s = getenv("BPGPU_S");
  static const char* env_l1 = getenv("BPGPU_LGL1");
  if (env_c) g.c = atoi(env_c);
  g.W0 = (Curve::SCALAR_BITS + 1 + g.c - 1) / g.c;
  g.W = g.W0 * (d_scalars2 ? 2 : 1);
  g.nbp = (1u << (g.c - 1)) + 1;
  {
    // Chunk length: long enough that a bucket is split into few partial sums (the reduction reads them all), short enough
    // that W*n/S chunk threads still fill the machine (a chunk is a serial chain of S mixed additions, ~10 us each).
    // Measured at 2^20 (tools/msm_sweep.py): S = 64 beats 32 by 5 %; 96 and 128 lose in k_chunk_acc.
    uint64_t fill = ((uint64_t)g.W * n) / ((uint64_t)ctx->sm_count * 8 * 32);
    uint64_t s = fill;
    g.S = (uint32_t)(s < 8 ? 8 : (s > 64 ? 64 : s));   // below 8 a bucket splits into so many partial sums that merging them dominates
    if (env_s) g.S = (uint32_t)atoi(env_s);
  }
  g.nchunk = (g.n + g.S - 1) / g.S;
  g.pcap = g.nchunk + g.nbp;
  {
    uint32_t nb = g.nbp - 1;                       // 2^(c-1)
    int lg = 0;                                    // buckets per lane: up to 16, but keep >= ~600 warps busy
    while ((32u << (lg + 1)) <= nb && lg < 4 && ((uint64_t)g.W * nb >> (lg + 1 + 5)) >= 600) lg++;
    if (env_l1) lg = atoi(env_l1);
    g.lgL1 = lg;
    g.nseg = (nb + (32u << lg) - 1) / (32u << lg);
    int lg2 = 0;
    while ((32u << lg2) < g.nseg) lg2++;
    g.lgL2 = lg2;
  }

  static const bool small_on = !(getenv("BPGPU_SMALL") && atoi(getenv("BPGPU_SMALL")) == 0);
  // measured (tools/small_msm_ab.py, tools/proof_rate.py): the one-launch kernel wins from a few hundred terms up on both curves
  // and at every small size on BN254; on BLS12-381 (175 registers x the block's threads held to the end) the 154-term
  // verification MSM of a 64-bit range proof is no faster alone (0.69 against 0.66 ms) and 10 % slower with 16 contexts
  const bool small_fits = n <= SMALL_MAX_N && g.c <= 8 && (Fq::N == 8 || n > 256);
  if (small_on && small_fits) {
    // ---- small input: digits, then ONE block per window (k_msm_small)
    const size_t szd = align256((size_t)g.W * n * 4);
    const size_t szs = align256((size_t)g.W * (2 * SMALL_T + g.nbp) * sizeof(XYZZ<Fq>));
    const size_t szw = align256((size_t)2 * g.W * sizeof(XYZZ<Fq>));
    int rc;
    if ((rc = ctx->msm_a.reserve(szd)) || (rc = ctx->msm_b.reserve(szs + szw))) return rc;
    uint32_t* digits = (uint32_t*)ctx->msm_a.p;
    XYZZ<Fq>* scratch = (XYZZ<Fq>*)ctx->msm_b.p;
    XYZZ<Fq>* winsum = (XYZZ<Fq>*)((uint8_t*)ctx->msm_b.p + szs);
    StageTimer tm(st, ctx->profile != 0 && n >= (size_t)ctx->profile);
    k_digits<Fr><<<dim3((g.n + 255) / 256, d_scalars2 ? 2 : 1), 256, 0, st>>>((const Fr*)d_scalars, (const Fr*)d_scalars2, scalars_mont ? 1 : 0, g, digits,
                                                                               (uint32_t*)nullptr);
    tm.mark("digits");
    if (ctx->wait_points) {
      ctx->wait_points = false;
      BP_CUDA_OK(cudaStreamWaitEvent(st, ctx->points_ready, 0));
    }
    const size_t smem = ((size_t)3 * (g.nbp + 2) + n) * sizeof(uint32_t);
    // block size by input size: the block's registers are held until its warp 0 has finished the bucket reduction, so a
    // small input (the 154-term verification MSM of a 64-bit range proof) gets 128 threads and two or more blocks share an SM
    int T = n <= 1024 ? 128 : SMALL_T;
    if ((uint32_t)T < g.nbp) T = SMALL_T;              // one thread per bucket in the folding phase (nbp <= 129)
    k_msm_small<Fq><<<g.W, T, smem, st>>>(g, d_points, digits, scratch, winsum, winsum + g.W);
    tm.mark("small");
    ctx->launches += 2;
    g.S = 0; g.lgL1 = 0; g.nseg = 1;
    res->W = g.W0; res->c = g.c; res->qshift = 5; res->d_winsum = winsum;
    int lrc = launch_check(ctx, "msm_small");
    tm.report(g, ctx);
    return lrc;
  }

  // ---- scratch layout
  const size_t sz_digits = align256((size_t)g.W * n * 4);
  const size_t sz_hist = align256((size_t)g.W * g.nbp * 4);
  const size_t sz_bstart = align256((size_t)g.W * (g.nbp + 1) * 4);
  const size_t sz_giant = align256(((size_t)g.W * g.nbp + 1) * 4);
  int rc;
  if ((rc = ctx->msm_a.reserve(sz_digits * 2 + sz_hist * 2 + sz_bstart * 2 + sz_giant))) return rc;
  uint8_t* base = (uint8_t*)ctx->msm_a.p;
  uint32_t* digits = (uint32_t*)base; base += sz_digits;
  uint32_t* sidx = (uint32_t*)base; base += sz_digits;
  uint32_t* hist = (uint32_t*)base; base += sz_hist;
  uint32_t* cursor = (uint32_t*)base; base += sz_hist;
  uint32_t* bstart = (uint32_t*)base; base += sz_bstart;
  uint32_t* pstart = (uint32_t*)base; base += sz_bstart;
  uint32_t* giant = (uint32_t*)base;  // [0] = count, [1..] = list
  const size_t sz_part = align256((size_t)g.W * g.pcap * sizeof(XYZZ<Fq>));
  const size_t sz_seg = align256((size_t)g.W * g.nseg * sizeof(XYZZ<Fq>));
  const size_t sz_wsum = align256((size_t)2 * g.W * sizeof(XYZZ<Fq>));
  if ((rc = ctx->msm_b.reserve(sz_part + 2 * sz_seg + sz_wsum))) return rc;
  uint8_t* b2 = (uint8_t*)ctx->msm_b.p;
  XYZZ<Fq>* partials = (XYZZ<Fq>*)b2; b2 += sz_part;
  XYZZ<Fq>* segA = (XYZZ<Fq>*)b2; b2 += sz_seg;
  XYZZ<Fq>* segS = (XYZZ<Fq>*)b2; b2 += sz_seg;
  XYZZ<Fq>* winsum = (XYZZ<Fq>*)b2;                            // [0, W): P_w   [W, 2W): Q_w

  StageTimer tm(st, ctx->profile != 0 && n >= (size_t)ctx->profile);   // profile = smallest n that is recorded
  BP_CUDA_OK(cudaMemsetAsync(hist, 0, sz_hist, st));
  BP_CUDA_OK(cudaMemsetAsync(giant, 0, 4, st));

  k_digits<Fr><<<dim3((g.n + 255) / 256, d_scalars2 ? 2 : 1), 256, 0, st>>>((const Fr*)d_scalars, (const Fr*)d_scalars2, scalars_mont ? 1 : 0, g, digits, hist);
  tm.mark("digits");
  k_scan<<<g.W, 1024, 0, st>>>(g, hist, bstart, cursor, pstart, giant, giant + 1);
  tm.mark("scan");
  {
    size_t total = (size_t)g.W * n;
    size_t blocks = (total + 255) / 256;
    size_t cap = (size_t)ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
    k_scatter<<<(unsigned)blocks, 256, 0, st>>>(g, digits, cursor, sidx);
    tm.mark("scatter");
  }
  if (ctx->wait_points) {                    // points still arriving on the copy queue (bpgpu_msm_refs)
    ctx->wait_points = false;
    BP_CUDA_OK(cudaStreamWaitEvent(st, ctx->points_ready, 0));
  }
  {
    uint32_t threads = (uint32_t)g.W * g.nchunk;
    if (threads >= (uint32_t)ctx->sm_count * 3 * 128)           // at least one resident wave of the compact form: throughput bound
      k_chunk_acc<Fq, true><<<(threads + 127) / 128, 128, 0, st>>>(g, d_points, sidx, bstart, pstart, partials);
    else
      k_chunk_acc<Fq, false><<<(threads + 127) / 128, 128, 0, st>>>(g, d_points, sidx, bstart, pstart, partials);
    tm.mark("chunk_acc");
  }
  tm.mark("giant");                          // (kept as a stage name: its work now rides the merge launch)
  {
    uint32_t threads = (uint32_t)g.W * g.nchunk;
    const uint32_t mb = (threads + 127) / 128;
    k_merge<Fq><<<mb + (uint32_t)ctx->sm_count * 2, 128, 0, st>>>(g, mb, bstart, pstart, hist, partials, giant, giant + 1);
  }
  tm.mark("merge");
  const int qshift = 5 + g.lgL1;
  {
    uint32_t warps = (uint32_t)g.W * g.nseg;
    if (g.nseg == 1) {                         // small windows: level 1 writes the window sums itself
      k_reduce_l1<Fq><<<(warps * 32 + 127) / 128, 128, 0, st>>>(g, pstart, hist, partials, winsum, winsum + g.W);
      tm.mark("reduce_l1");
      tm.mark("reduce_l2");
    } else {
      k_reduce_l1<Fq><<<(warps * 32 + 127) / 128, 128, 0, st>>>(g, pstart, hist, partials, segA, segS);
      tm.mark("reduce_l1");
      k_reduce_l2<Fq><<<g.W, 64, 0, st>>>(g, segA, segS, winsum, winsum + g.W);
      tm.mark("reduce_l2");
      ctx->launches += 1;
    }
  }
  ctx->launches += 6;
  res->W = g.W0; res->c = g.c; res->qshift = qshift; res->d_winsum = winsum;    // P at [0, nsets*W), Q at [nsets*W, 2*nsets*W)
  int lrc = launch_check(ctx, "msm");
  tm.report(g, ctx);
  return lrc;
}

template int msm_run<Bls>(bpgpu_ctx*, const Affine<Bls::Fq>*, const void*, bool, size_t, MsmResult*, const void*);
template int msm_run<Bn>(bpgpu_ctx*, const Affine<Bn::Fq>*, const void*, bool, size_t, MsmResult*, const void*);

}  // namespace bp
