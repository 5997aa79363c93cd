// Fixed-base commitments: out_i = sum_j s_ij * B_j for a handful of bases B_0..B_{k-1} that stay the same across
// thousands of calls -- the Pedersen pair (g, h) of the R1CS prover.
//
// Replaces commitment::commit_to_field_element(g, h, v, r) = g.binary_scalar_mul(h, v, r) and `g * w`
// (/root/reference/src/r1cs/prover.rs:123 V_i, :496-500 T_1..T_6, :550 Q; gadgets/poseidon_hash.rs:49-62), which AMCL
// computes with a 255-step double-and-add each.  A doubling chain is the one thing a GPU thread is bad at (measured:
// 9 us per dependent XYZZ doubling, tools/latency_probe.py), so for fixed bases the doublings are done ONCE:
// the table T[j][w][d-1] = d * 2^(8w) * B_j (w < 32, d = 1..255, affine) turns every later commitment into a sum of
// <= 32k table entries, reduced by a block-wide tree: no doublings, one launch for a whole batch.
#include <stdlib.h>
#include <string.h>
#include <thread>

#include "common.cuh"
#include "host_fp.h"

namespace bp {

static const int FB_WINDOWS = TBL_WINDOWS;      // 8-bit unsigned windows cover 256 bits
static const int FB_DIGITS = TBL_DIGITS;
static const int FB_GROUP = 15;                 // multiples normalised together by one thread (one shared inversion)
static const int FB_GROUPS = TBL_DIGITS / FB_GROUP;   // 17

}  // namespace bp

struct bpgpu_fixed_bases {
  bpgpu_ctx* ctx;
  size_t k;
  void* table;                   // Affine[k][32][255]
  std::vector<uint8_t> key;      // the k bases as X||Y bytes (cache key)
};

namespace bp {

template <class Fq>
__device__ __forceinline__ XYZZ<Fq> block_tree_sum_256(const XYZZ<Fq>& v, XYZZ<Fq>* sm) {
  store_vec(sm + threadIdx.x, v);
  __syncthreads();
  for (int o = (int)blockDim.x / 2; o > 0; o >>= 1) {      // blockDim.x = 64 or 256
    if ((int)threadIdx.x < o) {
      XYZZ<Fq> a = load_vec(sm + threadIdx.x), b = load_vec(sm + threadIdx.x + o);
      a.add(b);
      store_vec(sm + threadIdx.x, a);
    }
    __syncthreads();
  }
  return load_vec(sm);
}

// pow2[j][w] = 2^(8w) * B_j : one thread per base, a serial chain of 248 doublings (one-off)
template <class Fq>
__global__ void k_fb_pow2(const Affine<Fq>* __restrict__ bases, size_t k, XYZZ<Fq>* __restrict__ pow2) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  XYZZ<Fq> p = XYZZ<Fq>::from_affine(load_vec(bases + j));
  for (int w = 0; w < FB_WINDOWS; w++) {
    store_vec(pow2 + (size_t)j * FB_WINDOWS + w, p);
    if (w + 1 < FB_WINDOWS) for (int b = 0; b < TBL_BITS; b++) p.dbl();
  }
}

// table[j][w][d-1] = d * pow2[j][w], normalised to affine: one thread per (j, w, group of 15 consecutive multiples); the
// 15 multiples of a thread share ONE field inversion (Montgomery's trick on the ZZZ coordinates)
template <class Fq>
__global__ void __launch_bounds__(64) k_fb_multiples(const XYZZ<Fq>* __restrict__ pow2, size_t total, Affine<Fq>* __restrict__ table) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;             // (j * W + w) * FB_GROUPS + g
  if (t >= total * FB_GROUPS) return;
  const size_t jw = t / FB_GROUPS;
  const uint32_t g = (uint32_t)(t - jw * FB_GROUPS);
  const XYZZ<Fq> base = load_vec(pow2 + jw);
  Affine<Fq>* out = table + jw * FB_DIGITS + (size_t)g * FB_GROUP;      // multiples 15g+1 .. 15g+15
  if (base.is_inf()) {
    for (int d = 0; d < FB_GROUP; d++) store_vec(out + d, Affine<Fq>::inf());
    return;
  }
  XYZZ<Fq> m[FB_GROUP];
  Fq pre[FB_GROUP];                        // pre[d] = product of zzz_0 .. zzz_d over the finite multiples
  XYZZ<Fq> acc = g ? mul_small(base, g * FB_GROUP + 1) : base;
  Fq run = Fq::one();
#pragma unroll 1
  for (int d = 0; d < FB_GROUP; d++) {
    m[d] = acc;
    if (!acc.is_inf()) run = run * acc.zzz;
    pre[d] = run;
    if (d + 1 < FB_GROUP) acc.add(base);
  }
  Fq inv = run.inv();
#pragma unroll 1
  for (int d = FB_GROUP - 1; d >= 0; d--) {
    if (m[d].is_inf()) { store_vec(out + d, Affine<Fq>::inf()); continue; }      // only for points of small order: never on these curves
    Fq i3 = d ? inv * pre[d - 1] : inv;     // 1 / zzz_d
    inv = inv * m[d].zzz;
    Fq i1 = i3 * m[d].zz;                   // 1 / z
    Affine<Fq> a;
    a.x = m[d].x * i1.sqr();
    a.y = m[d].y * i3;
    store_vec(out + d, a);
  }
}

// 8 threads per term: thread (p, j) adds the table entries of the 4 byte-windows of scalar limb j of term p; block tree;
// one XYZZ per block.
// blockIdx.y selects the group (independent sum) the block works for.
struct TableSegs {
  const void* table[TBL_MAX_SEGS]; const void* scal[TBL_MAX_SEGS]; const uint32_t* rows[TBL_MAX_SEGS];
  uint32_t mont[TBL_MAX_SEGS];
  uint32_t start[TBL_MAX_SEGS + 1];            // first term of the segment within its group
  uint32_t gfirst[TBL_MAX_GROUPS + 1];         // first segment of each group
  uint32_t gtotal[TBL_MAX_GROUPS];             // terms per group
};
// Grid-stride over (term, limb) items with SMALL blocks (6 tree levels instead of the 8 of a 256-thread block: every level
// is a dependent full addition on a shrinking number of lanes): a thread adds the 4 byte-window entries of each of its
// items, then the block tree leaves one XYZZ per block.  The next table entry is requested before the current one is
// added: entries of a multi-GB table set are DRAM misses.
static const int TS_THREADS = 64;
// FIXED = the fixed-schedule variant for SECRET scalars (bpgpu_ctx_set_fixed_schedule; the reference commits to the witness
// with inner_product_const_time, prover.rs:347-362): every (term, limb) item performs exactly 4 table loads and 4 mixed
// additions whatever the digits are -- a zero digit loads entry 1 and discards the sum by a lane-wise select; zero limbs and
// zero scalars are not skipped.  What still depends on the scalars: the table ADDRESSES (one entry is fetched per digit, as
// AMCL's windowed multiplication selects one -- but a GPU fetch is not a masked scan) and the first addition into an empty
// accumulator.  See DESIGN.md "secret scalars".
template <class Curve, bool FIXED>
__global__ void __launch_bounds__(TS_THREADS) k_table_sum(TableSegs segs, XYZZ<typename Curve::Fq>* __restrict__ partial) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  __shared__ __align__(16) unsigned char smraw[TS_THREADS * sizeof(XYZZ<Fq>)];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(smraw);
  const uint32_t grp = blockIdx.y;
  const uint32_t items = segs.gtotal[grp] * 8;
  const uint32_t last = segs.gfirst[grp + 1] - 1;
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  for (uint32_t it = blockIdx.x * TS_THREADS + threadIdx.x; it < items; it += gridDim.x * TS_THREADS) {
    const uint32_t p = it >> 3, j = it & 7;
    uint32_t sg = segs.gfirst[grp];
    while (sg < last && p >= segs.start[sg + 1]) sg++;
    const uint32_t idx = p - segs.start[sg];
    uint32_t limb;
    if (segs.mont[sg]) { Fr sc = load_vec((const Fr*)segs.scal[sg] + idx); limb = (!FIXED && sc.is_zero()) ? 0u : sc.from_mont().v[j]; }
    else limb = ((const Fr*)segs.scal[sg])[idx].v[j];
    if (!FIXED && !limb) continue;
    const uint32_t row = segs.rows[sg] ? segs.rows[sg][idx] : idx;
    const Affine<Fq>* tb = (const Affine<Fq>*)segs.table[sg] + ((size_t)row * TBL_WINDOWS + TBL_PER_LIMB * j) * TBL_DIGITS;
    if (FIXED) {
#pragma unroll 1
      for (int k = 0; k < TBL_PER_LIMB; k++) {
        const uint32_t dk = (limb >> (TBL_BITS * k)) & (uint32_t)TBL_DIGITS;
        const Affine<Fq> e = load_vec_ro(tb + k * TBL_DIGITS + (dk ? dk - 1 : 0));
        XYZZ<Fq> t = acc;
        t.madd(e);
        acc = select(dk != 0, t, acc);
      }
      continue;
    }
    uint32_t d = limb & (uint32_t)TBL_DIGITS;
    Affine<Fq> cur = d ? load_vec_ro(tb + (d - 1)) : Affine<Fq>::inf();
#pragma unroll 1
    for (int k = 0; k < TBL_PER_LIMB; k++) {
      Affine<Fq> nxt = Affine<Fq>::inf();
      if (k + 1 < TBL_PER_LIMB) {
        const uint32_t dn = (limb >> (TBL_BITS * (k + 1))) & (uint32_t)TBL_DIGITS;
        if (dn) nxt = load_vec_ro(tb + (k + 1) * TBL_DIGITS + (dn - 1));
      }
      acc.madd(cur);                                   // madd of the identity (0, 0) is a no-op
      cur = nxt;
    }
  }
  XYZZ<Fq> tot = block_tree_sum_256(acc, sm);
  if (threadIdx.x == 0) store_vec(partial + (size_t)grp * gridDim.x + blockIdx.x, tot);
}

// block g sums the `count` block results of group g into out[g]
template <class Fq>
__global__ void __launch_bounds__(256) k_table_sum_final(const XYZZ<Fq>* __restrict__ partial, uint32_t count, XYZZ<Fq>* __restrict__ out) {
  __shared__ __align__(16) unsigned char smraw[256 * sizeof(XYZZ<Fq>)];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(smraw);
  const XYZZ<Fq>* src = partial + (size_t)blockIdx.x * count;
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  for (uint32_t k = threadIdx.x; k < count; k += blockDim.x) { XYZZ<Fq> q = load_vec(src + k); acc.add(q); }
  XYZZ<Fq> tot = block_tree_sum_256(acc, sm);
  if (threadIdx.x == 0) store_vec(out + blockIdx.x, tot);
}

// one block of FB_WINDOWS (32) threads per commitment: thread w sums the k table entries of window w, then a block tree
template <class Fq>
__global__ void __launch_bounds__(FB_WINDOWS) k_fb_commit(const Affine<Fq>* __restrict__ table, int k, const ScalarInt* __restrict__ scalars,
                                                          XYZZ<Fq>* __restrict__ out) {
  __shared__ __align__(16) unsigned char smraw[FB_WINDOWS * sizeof(XYZZ<Fq>)];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(smraw);
  const int w = threadIdx.x;
  const size_t inst = blockIdx.x;
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  for (int j = 0; j < k; j++) {
    const uint32_t limb = scalars[inst * k + j].v[w / TBL_PER_LIMB];
    const uint32_t d = (limb >> ((w % TBL_PER_LIMB) * TBL_BITS)) & (uint32_t)TBL_DIGITS;
    if (d) acc.madd(load_vec_ro(table + ((size_t)j * FB_WINDOWS + w) * FB_DIGITS + (d - 1)));
  }
  store_vec(sm + w, acc);
  __syncthreads();
  for (int o = FB_WINDOWS / 2; o > 0; o >>= 1) {
    if (w < o) {
      XYZZ<Fq> a = load_vec(sm + w), b = load_vec(sm + w + o);
      a.add(b);
      store_vec(sm + w, a);
    }
    __syncthreads();
  }
  if (w == 0) store_vec(out + inst, load_vec(sm));
}

// the same for a SMALL batch whose scalars arrive as kernel arguments: the 2 scalars of one commitment or the 10 of
// T_1..T_6 need no staging copy and no conversion launch
static const int FB_INLINE_SCALARS = 16;
struct InlineScalars { ScalarInt s[FB_INLINE_SCALARS]; };
template <class Fq>
__global__ void __launch_bounds__(FB_WINDOWS) k_fb_commit_inline(const Affine<Fq>* __restrict__ table, int k, InlineScalars sc,
                                                                 XYZZ<Fq>* __restrict__ out) {
  __shared__ __align__(16) unsigned char smraw[FB_WINDOWS * sizeof(XYZZ<Fq>)];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(smraw);
  const int w = threadIdx.x;
  const size_t inst = blockIdx.x;
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  for (int j = 0; j < k; j++) {
    const uint32_t limb = sc.s[inst * k + j].v[w / TBL_PER_LIMB];
    const uint32_t d = (limb >> ((w % TBL_PER_LIMB) * TBL_BITS)) & (uint32_t)TBL_DIGITS;
    if (d) acc.madd(load_vec_ro(table + ((size_t)j * FB_WINDOWS + w) * FB_DIGITS + (d - 1)));
  }
  store_vec(sm + w, acc);
  __syncthreads();
  for (int o = FB_WINDOWS / 2; o > 0; o >>= 1) {
    if (w < o) {
      XYZZ<Fq> a = load_vec(sm + w), b = load_vec(sm + w + o);
      a.add(b);
      store_vec(sm + w, a);
    }
    __syncthreads();
  }
  if (w == 0) store_vec(out + inst, load_vec(sm));
}

// Builds the window tables of n affine points already on the device.  One-off: a 248-doubling chain per point (all
// points in parallel), then 17 threads per (point, window) with 14 additions and one inversion each.
template <class Curve>
int build_tables(bpgpu_ctx* ctx, const void* d_affine, size_t n, void** table_out) {
  using Fq = typename Curve::Fq;
  *table_out = nullptr;
  if (n == 0) return BPGPU_OK;
  const size_t tbytes = n * TBL_ENTRIES * sizeof(Affine<Fq>);
  if (tbytes > ((size_t)96 << 30)) return BPGPU_E_ARG;                       // tables are for generator sets, not for 2^20-point inputs
  void* d_pow2 = nullptr;
  void* table = nullptr;
  BP_CUDA_OK(cudaMalloc(&table, tbytes));
  // the XYZZ powers are scratch: process the points in slabs of at most 2^16
  const size_t slab = n < 65536 ? n : 65536;
  if (cudaMalloc(&d_pow2, slab * FB_WINDOWS * sizeof(XYZZ<Fq>)) != cudaSuccess) { cudaFree(table); return BPGPU_E_CUDA; }
  int rc = BPGPU_OK;
  for (size_t lo = 0; lo < n && !rc; lo += slab) {
    const size_t cnt = n - lo < slab ? n - lo : slab;
    k_fb_pow2<Fq><<<(unsigned)((cnt + 63) / 64), 64, 0, ctx->stream>>>((const Affine<Fq>*)d_affine + lo, cnt, (XYZZ<Fq>*)d_pow2);
    const size_t total = cnt * FB_WINDOWS;
    k_fb_multiples<Fq><<<(unsigned)((total * FB_GROUPS + 63) / 64), 64, 0, ctx->stream>>>((const XYZZ<Fq>*)d_pow2, total,
                                                                               (Affine<Fq>*)table + lo * TBL_ENTRIES);
    ctx->launches += 2;
    rc = launch_check(ctx, "build_tables");
  }
  if (!rc && stream_sync(ctx) != cudaSuccess) rc = BPGPU_E_CUDA;
  cudaFree(d_pow2);
  if (rc) { cudaFree(table); return rc; }
  *table_out = table;
  return BPGPU_OK;
}
template int build_tables<Bls>(bpgpu_ctx*, const void*, size_t, void**);
template int build_tables<Bn>(bpgpu_ctx*, const void*, size_t, void**);

// wide[pW][d-1] = d * 2^(16 W) * B_p for d = 1..65535 = the sum of the two byte-window entries T8[p][2W+1][d >> 8] and
// T8[p][2W][d & 255]: ONE mixed addition per entry, no doublings.  One thread per (p, W, 15 consecutive d) -- 65535 = 15 * 4369 --
// and one shared inversion per thread (Montgomery's trick), as k_fb_multiples.
static const int WIDE_GROUP = 15;
static const uint32_t WIDE_GROUPS = TBL16_DIGITS / WIDE_GROUP;     // 4369
template <class Fq>
__global__ void __launch_bounds__(64) k_fb_wide(const Affine<Fq>* __restrict__ table8, size_t npw, Affine<Fq>* __restrict__ table16) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // pW * WIDE_GROUPS + g
  if (t >= npw * WIDE_GROUPS) return;
  const size_t pw = t / WIDE_GROUPS;
  const uint32_t g = (uint32_t)(t - pw * WIDE_GROUPS);
  const size_t pt = pw / TBL16_WINDOWS, W = pw - pt * TBL16_WINDOWS;
  const Affine<Fq>* lo8 = table8 + (pt * TBL_WINDOWS + 2 * W) * TBL_DIGITS;
  const Affine<Fq>* hi8 = lo8 + TBL_DIGITS;
  Affine<Fq>* out = table16 + pw * TBL16_DIGITS + (size_t)g * WIDE_GROUP;
  XYZZ<Fq> m[WIDE_GROUP];
  Fq pre[WIDE_GROUP];
  Fq run = Fq::one();
#pragma unroll 1
  for (int k = 0; k < WIDE_GROUP; k++) {
    const uint32_t d = g * WIDE_GROUP + 1 + (uint32_t)k, dl = d & 255u, dh = d >> 8;
    XYZZ<Fq> acc = dh ? XYZZ<Fq>::from_affine(load_vec_ro(hi8 + (dh - 1))) : XYZZ<Fq>::inf();
    if (dl) acc.madd(load_vec_ro(lo8 + (dl - 1)));
    m[k] = acc;
    if (!acc.is_inf()) run = run * acc.zzz;
    pre[k] = run;
  }
  Fq inv = run.inv();
#pragma unroll 1
  for (int k = WIDE_GROUP - 1; k >= 0; k--) {
    if (m[k].is_inf()) { store_vec(out + k, Affine<Fq>::inf()); continue; }     // the identity among the bases
    Fq i3 = k ? inv * pre[k - 1] : inv;     // 1 / zzz_k
    inv = inv * m[k].zzz;
    Fq i1 = i3 * m[k].zz;                   // 1 / z
    Affine<Fq> a;
    a.x = m[k].x * i1.sqr();
    a.y = m[k].y * i3;
    store_vec(out + k, a);
  }
}

template <class Curve>
int build_tables16(bpgpu_ctx* ctx, const void* table8, size_t n, void** table16_out) {
  using Fq = typename Curve::Fq;
  *table16_out = nullptr;
  if (n == 0 || !table8) return BPGPU_E_ARG;
  const size_t bytes = n * TBL16_ENTRIES * sizeof(Affine<Fq>);
  if (bytes > ((size_t)48 << 30)) return BPGPU_E_ARG;                        // a few dozen generators, not a circuit of thousands
  void* t16 = nullptr;
  if (cudaMalloc(&t16, bytes) != cudaSuccess) { cudaGetLastError(); return BPGPU_E_CUDA; }
  const size_t npw = n * TBL16_WINDOWS, threads = npw * WIDE_GROUPS;
  k_fb_wide<Fq><<<(unsigned)((threads + 63) / 64), 64, 0, ctx->stream>>>((const Affine<Fq>*)table8, npw, (Affine<Fq>*)t16);
  ctx->launches++;
  int rc = launch_check(ctx, "k_fb_wide");
  if (!rc && stream_sync(ctx) != cudaSuccess) rc = BPGPU_E_CUDA;
  if (rc) { cudaFree(t16); return rc; }
  *table16_out = t16;
  return BPGPU_OK;
}
template int build_tables16<Bls>(bpgpu_ctx*, const void*, size_t, void**);
template int build_tables16<Bn>(bpgpu_ctx*, const void*, size_t, void**);

template <class Curve>
int table_sum_run(bpgpu_ctx* ctx, const TableSeg* segs, int nsegs, int ngroups, int* host_partials) {
  using Fq = typename Curve::Fq;
  if (nsegs > TBL_MAX_SEGS || ngroups < 1 || ngroups > TBL_MAX_GROUPS) return BPGPU_E_ARG;
  TableSegs ts;
  int q = 0;
  uint32_t maxtotal = 0;
  for (int g = 0; g < ngroups; g++) {
    ts.gfirst[g] = q;
    uint32_t total = 0;
    for (int k = 0; k < nsegs; k++) {
      if (segs[k].group != g || segs[k].n == 0) continue;
      ts.table[q] = segs[k].table; ts.scal[q] = segs[k].scalars; ts.mont[q] = segs[k].mont ? 1u : 0u; ts.rows[q] = segs[k].rows;
      ts.start[q] = total;
      total += segs[k].n;
      q++;
    }
    if (q == (int)ts.gfirst[g]) {                    // empty group: one dummy segment with no terms
      ts.table[q] = nullptr; ts.scal[q] = nullptr; ts.mont[q] = 0; ts.start[q] = 0; ts.rows[q] = nullptr;
      q++;
    }
    ts.gtotal[g] = total;
    if (total > maxtotal) maxtotal = total;
  }
  ts.gfirst[ngroups] = q;
  ts.start[q] = 0;
  if (q > TBL_MAX_SEGS) return BPGPU_E_ARG;
  // Every level of a block tree is one dependent XYZZ addition: these sums are latency bound, so blocks are small (6 tree
  // levels) and, when the caller finishes on the host anyway (host_partials) and there are few of them, the block results go
  // to the host as they are -- it adds a few dozen points in less time than a second launch takes.
  // one (term, limb) item per thread while that fits the machine (these sums are bound by the LATENCY of a thread's chain:
  // ~5 us per dependent mixed addition on BN254, ~10 us on BLS12-381); beyond sm_count * 8 blocks the grid strides
  const uint32_t items = maxtotal * 8;
  const uint32_t bs = TS_THREADS;
  static const uint32_t ipt_env = [] { const char* e = getenv("BPGPU_TS_ITEMS"); return e ? (uint32_t)atoi(e) : 0u; }();
  // items per thread: one for small sums (pure latency); from a few thousand items on, three -- the 63 tree additions of a
  // block cost as much issue time as its 256 accumulating ones, and a third of the blocks shortens a proof at n = 1024 too
  // (4.45 -> 4.1 ms) while 16 concurrent contexts gain 55 % (1234 -> 1917 proofs/s); BPGPU_TS_ITEMS overrides
  const uint32_t ipt = ipt_env ? ipt_env : (items >= 4096 ? 3u : 1u);
  uint32_t blocks = (items + bs * ipt - 1) / (bs * ipt);
  static const uint32_t per_sm = [] { const char* e = getenv("BPGPU_TS_BLOCKS_PER_SM"); return e ? (uint32_t)atoi(e) : 4u; }();
  uint32_t cap = (uint32_t)ctx->sm_count * per_sm / (uint32_t)ngroups;      // all groups together: one resident wave
  if (cap == 0) cap = 1;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  int rc = ctx->tbl_part.reserve(((size_t)blocks * ngroups + TBL_MAX_GROUPS) * sizeof(XYZZ<Fq>));
  if (rc) return rc;
  XYZZ<Fq>* out = (XYZZ<Fq>*)ctx->tbl_part.p;           // [0, ngroups) = results, then per-block sums
  if (host_partials) *host_partials = 0;
  static const bool prof = getenv("BPGPU_PROFILE") != nullptr;
  cudaEvent_t ev[2];
  if (prof) { for (auto& e : ev) cudaEventCreate(&e); cudaEventRecord(ev[0], ctx->stream); }
  struct ProfEnd {
    bool on; cudaEvent_t* ev; cudaStream_t st; uint32_t terms, blocks; int groups;
    ~ProfEnd() {
      if (!on) return;
      cudaEventRecord(ev[1], st); cudaEventSynchronize(ev[1]);
      float ms; cudaEventElapsedTime(&ms, ev[0], ev[1]);
      fprintf(stderr, "[bpgpu table_sum terms=%u groups=%d blocks=%u] %.3f ms\n", terms, groups, blocks, ms);
      cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
    }
  } prof_end{prof, ev, ctx->stream, maxtotal, blocks, ngroups};
  auto launch = [&](dim3 grid, XYZZ<Fq>* dst) {
    if (ctx->fixed_schedule) k_table_sum<Curve, true><<<grid, bs, 0, ctx->stream>>>(ts, dst);
    else k_table_sum<Curve, false><<<grid, bs, 0, ctx->stream>>>(ts, dst);
  };
  if (blocks == 1) {
    launch(dim3(1, ngroups), out);
    ctx->launches += 1;
  } else if (host_partials && blocks * ngroups <= (uint32_t)TBL_HOST_PARTIALS) {
    launch(dim3(blocks, ngroups), out + TBL_MAX_GROUPS);
    ctx->launches += 1;
    *host_partials = (int)blocks;                       // group g: out[TBL_MAX_GROUPS + g * blocks + b], b < blocks
  } else {
    launch(dim3(blocks, ngroups), out + TBL_MAX_GROUPS);
    k_table_sum_final<Fq><<<ngroups, 256, 0, ctx->stream>>>(out + TBL_MAX_GROUPS, blocks, out);
    ctx->launches += 2;
  }
  return launch_check(ctx, "table_sum");
}
template int table_sum_run<Bls>(bpgpu_ctx*, const TableSeg*, int, int, int*);
template int table_sum_run<Bn>(bpgpu_ctx*, const TableSeg*, int, int, int*);

template <class Curve>
static int fb_build(bpgpu_fixed_bases* fb, const uint8_t* bases_xy) {
  using Fq = typename Curve::Fq;
  bpgpu_ctx* ctx = fb->ctx;
  void* d_bases = nullptr;
  BP_CUDA_OK(cudaMalloc(&d_bases, fb->k * sizeof(Affine<Fq>)));
  int rc = points_from_host<Curve>(ctx, bases_xy, fb->k, d_bases);
  if (!rc) rc = build_tables<Curve>(ctx, d_bases, fb->k, &fb->table);      // synchronises
  cudaFree(d_bases);
  if (!rc && (rc = inputs_ok(ctx))) { cudaFree(fb->table); fb->table = nullptr; }   // a base that is not a point: BPGPU_E_FORMAT
  return rc;
}

// XYZZ results -> affine X||Y bytes with ONE field inversion for the whole batch (Montgomery's trick), on the host
template <class FqParams>
static void normalise_batch_host(const uint8_t* xyzz_bytes, size_t count, int modbytes, uint8_t* out_xy) {
  using HP = host::HXYZZ<FqParams>;
  using F = typename HP::F;
  const HP* p = reinterpret_cast<const HP*>(xyzz_bytes);
  std::vector<F> prefix(count + 1);
  prefix[0] = F::one();
  for (size_t i = 0; i < count; i++) prefix[i + 1] = p[i].is_inf() ? prefix[i] : prefix[i] * p[i].zzz;
  F inv = prefix[count].inv();
  for (size_t i = count; i-- > 0;) {
    uint8_t* o = out_xy + i * 2 * modbytes;
    if (p[i].is_inf()) { memset(o, 0, 2 * modbytes); o[2 * modbytes - 1] = 1; continue; }
    F i3 = inv * prefix[i];            // 1 / zzz_i
    inv = inv * p[i].zzz;
    F i1 = i3 * p[i].zz;               // 1 / z
    (p[i].x * i1.sqr()).to_be(o, modbytes);
    (p[i].y * i3).to_be(o + modbytes, modbytes);
  }
}

// in place: entry g of `xyzz_bytes` becomes the sum of entries [g * per, (g + 1) * per)
template <class FqParams>
void host_sum_partials(uint8_t* xyzz_bytes, int ngroups, int per) {
  using HP = host::HXYZZ<FqParams>;
  HP* p = reinterpret_cast<HP*>(xyzz_bytes);
  for (int g = 0; g < ngroups; g++) {
    HP acc = p[(size_t)g * per];
    for (int b = 1; b < per; b++) acc.add(p[(size_t)g * per + b]);
    p[g] = acc;
  }
}
template void host_sum_partials<BlsFq>(uint8_t*, int, int);
template void host_sum_partials<BnFq>(uint8_t*, int, int);

// large batches (the lock-step prover reads back thousands of sums per stage) are cut into chunks of >= 128 points, one
// host thread and one inversion each
void normalise_points_host(int curve, const uint8_t* xyzz_bytes, size_t count, uint8_t* out_xy) {
  const size_t mb = curve == BPGPU_BLS12_381 ? 48 : 32;
  const size_t psz = curve == BPGPU_BLS12_381 ? sizeof(XYZZ<Bls::Fq>) : sizeof(XYZZ<Bn::Fq>);
  auto run = [&](size_t lo, size_t cnt) {
    if (curve == BPGPU_BLS12_381) normalise_batch_host<BlsFq>(xyzz_bytes + lo * psz, cnt, 48, out_xy + lo * 2 * mb);
    else normalise_batch_host<BnFq>(xyzz_bytes + lo * psz, cnt, 32, out_xy + lo * 2 * mb);
  };
  size_t nt = std::thread::hardware_concurrency();
  if (nt > 16) nt = 16;
  if (nt > count / 128) nt = count / 128;
  if (nt <= 1) { run(0, count); return; }
  std::vector<std::thread> th;
  const size_t per = (count + nt - 1) / nt;
  for (size_t k = 1; k < nt; k++) {
    const size_t lo = k * per;
    if (lo >= count) break;
    th.emplace_back(run, lo, count - lo < per ? count - lo : per);
  }
  run(0, per < count ? per : count);
  for (auto& t : th) t.join();
}

}  // namespace bp
namespace bp {
int msm_tables_to_host(bpgpu_ctx* ctx, const TableSeg* segs, int nsegs, int ngroups, uint8_t* const* outs_xy) {
  const bool bls = ctx->curve == BPGPU_BLS12_381;
  const size_t psz = bls ? sizeof(XYZZ<Bls::Fq>) : sizeof(XYZZ<Bn::Fq>);
  const int mb = bpgpu_modbytes(ctx->curve);
  int hp = 0;
  int rc = bls ? table_sum_run<Bls>(ctx, segs, nsegs, ngroups, &hp) : table_sum_run<Bn>(ctx, segs, nsegs, ngroups, &hp);
  if (rc) return rc;
  if (hp) {
    BP_CUDA_OK(cudaMemcpyAsync(ctx->pinned, (const uint8_t*)ctx->tbl_part.p + TBL_MAX_GROUPS * psz, (size_t)ngroups * hp * psz,
                               cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA_OK(stream_sync(ctx));
    if (bls) host_sum_partials<BlsFq>(ctx->pinned, ngroups, hp);
    else host_sum_partials<BnFq>(ctx->pinned, ngroups, hp);
  } else {
    BP_CUDA_OK(cudaMemcpyAsync(ctx->pinned, ctx->tbl_part.p, ngroups * psz, cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA_OK(stream_sync(ctx));
  }
  if ((rc = inputs_ok(ctx))) return rc;
  uint8_t tmp[TBL_MAX_GROUPS * 2 * 48];
  if (bls) normalise_batch_host<BlsFq>(ctx->pinned, ngroups, mb, tmp);
  else normalise_batch_host<BnFq>(ctx->pinned, ngroups, mb, tmp);
  for (int g = 0; g < ngroups; g++) memcpy(outs_xy[g], tmp + (size_t)g * 2 * mb, 2 * mb);
  return BPGPU_OK;
}

template <class Curve>
static int fb_commit(bpgpu_fixed_bases* fb, const uint8_t* scalars_be, size_t count, uint8_t* out_xy) {
  using Fq = typename Curve::Fq;
  using FqParams = typename std::conditional<Curve::ID == BPGPU_BLS12_381, BlsFq, BnFq>::type;
  bpgpu_ctx* ctx = fb->ctx;
  const size_t ns = count * fb->k;
  int rc;
  if ((rc = ctx->msm_c.reserve(ns * 32 + 32))) return rc;
  if ((rc = ctx->msm_e.reserve(count * sizeof(XYZZ<Fq>) + 32))) return rc;
  if (ns <= (size_t)FB_INLINE_SCALARS) {
    // scalars on the ABI are canonical (< r), big endian, MODBYTES long: the low 32 bytes are the 8 limbs
    InlineScalars sc;
    memset(&sc, 0, sizeof sc);
    for (size_t i = 0; i < ns; i++) {
      const uint8_t* be = scalars_be + i * Curve::MODBYTES;
      for (int k = 0; k < 8; k++) {
        const uint8_t* p = be + Curve::MODBYTES - 4 * (k + 1);
        sc.s[i].v[k] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
      }
    }
    k_fb_commit_inline<Fq><<<(unsigned)count, FB_WINDOWS, 0, ctx->stream>>>((const Affine<Fq>*)fb->table, (int)fb->k, sc, (XYZZ<Fq>*)ctx->msm_e.p);
  } else {
    if ((rc = scalars_from_host<Curve>(ctx, scalars_be, ns, 0, ctx->msm_c.p))) return rc;
    k_fb_commit<Fq><<<(unsigned)count, FB_WINDOWS, 0, ctx->stream>>>((const Affine<Fq>*)fb->table, (int)fb->k, (const ScalarInt*)ctx->msm_c.p,
                                                           (XYZZ<Fq>*)ctx->msm_e.p);
  }
  ctx->launches++;
  if ((rc = launch_check(ctx, "k_fb_commit"))) return rc;
  const size_t bytes = count * sizeof(XYZZ<Fq>);
  std::vector<uint8_t> big;
  uint8_t* stage = ctx->pinned;
  if (bytes > ctx->pinned_cap / 2) { big.resize(bytes); stage = big.data(); }
  BP_CUDA_OK(cudaMemcpyAsync(stage, ctx->msm_e.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA_OK(stream_sync(ctx));
  normalise_batch_host<FqParams>(stage, count, Curve::MODBYTES, out_xy);
  return BPGPU_OK;
}

// table of a host point that is one of the bases of a cached fixed-base set (e.g. h of the Pedersen pair)
const void* fixed_table_lookup(bpgpu_ctx* ctx, const uint8_t* xy) {
  const size_t pb = 2 * (size_t)bpgpu_modbytes(ctx->curve);
  const size_t asz = ctx->curve == BPGPU_BLS12_381 ? sizeof(Affine<Bls::Fq>) : sizeof(Affine<Bn::Fq>);
  for (bpgpu_fixed_bases* fb : ctx->fb_cache)
    for (size_t j = 0; j < fb->k; j++)
      if (memcmp(fb->key.data() + j * pb, xy, pb) == 0) return (const uint8_t*)fb->table + j * TBL_ENTRIES * asz;
  return nullptr;
}

}  // namespace bp

using namespace bp;

extern "C" {

void bpgpu_fixed_bases_free(bpgpu_fixed_bases* fb) {
  if (!fb) return;
  cudaSetDevice(fb->ctx->device);
  if (fb->table) cudaFree(fb->table);
  delete fb;
}

int bpgpu_fixed_bases_create(bpgpu_ctx* ctx, const uint8_t* bases_xy, size_t k, bpgpu_fixed_bases** out) {
  if (!ctx || !bases_xy || !out || k == 0 || k > 64) return BPGPU_E_ARG;
  *out = nullptr;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  bpgpu_fixed_bases* fb = new (std::nothrow) bpgpu_fixed_bases();
  if (!fb) return BPGPU_E_CUDA;
  fb->ctx = ctx; fb->k = k; fb->table = nullptr;
  fb->key.assign(bases_xy, bases_xy + k * 2 * bpgpu_modbytes(ctx->curve));
  int rc = ctx->curve == BPGPU_BLS12_381 ? fb_build<Bls>(fb, bases_xy) : fb_build<Bn>(fb, bases_xy);
  if (rc) { bpgpu_fixed_bases_free(fb); return rc; }
  *out = fb;
  return BPGPU_OK;
}

int bpgpu_fixed_bases_commit(bpgpu_ctx* ctx, bpgpu_fixed_bases* fb, const uint8_t* scalars_be, size_t count, uint8_t* out_xy) {
  if (!ctx || !fb || fb->ctx != ctx || (!scalars_be && count) || (!out_xy && count)) return BPGPU_E_ARG;
  if (count == 0) return BPGPU_OK;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  return ctx->curve == BPGPU_BLS12_381 ? fb_commit<Bls>(fb, scalars_be, count, out_xy) : fb_commit<Bn>(fb, scalars_be, count, out_xy);
}

// ctx-owned cache: the table for a given list of bases is built on first use and lives as long as the ctx
int bpgpu_fixed_bases_get(bpgpu_ctx* ctx, const uint8_t* bases_xy, size_t k, bpgpu_fixed_bases** out) {
  if (!ctx || !bases_xy || !out || k == 0) return BPGPU_E_ARG;
  const size_t len = k * 2 * bpgpu_modbytes(ctx->curve);
  for (bpgpu_fixed_bases* fb : ctx->fb_cache)
    if (fb->k == k && fb->key.size() == len && memcmp(fb->key.data(), bases_xy, len) == 0) { *out = fb; return BPGPU_OK; }
  if (ctx->fb_cache.size() >= 16) { bpgpu_fixed_bases_free(ctx->fb_cache.front()); ctx->fb_cache.erase(ctx->fb_cache.begin()); }
  int rc = bpgpu_fixed_bases_create(ctx, bases_xy, k, out);
  if (!rc) ctx->fb_cache.push_back(*out);
  return rc;
}

}  // extern "C"
