// Fixed-base commitments: out_i = sum_j s_ij * B_j for a handful of bases B_0..B_{k-1} that stay the same across
// thousands of calls -- the Pedersen pair (g, h) of the R1CS prover.
//
// Replaces commitment::commit_to_field_element(g, h, v, r) = g.binary_scalar_mul(h, v, r) and `g * w`
// (/root/reference/src/r1cs/prover.rs:123 V_i, :496-500 T_1..T_6, :550 Q; gadgets/poseidon_hash.rs:49-62), which AMCL
// computes with a 255-step double-and-add each.  A doubling chain is the one thing a GPU thread is bad at (measured:
// 9 us per dependent XYZZ doubling, tools/latency_probe.py), so for fixed bases the doublings are done ONCE:
// the table T[j][w][d-1] = d * 2^(4w) * B_j (w < 64, d = 1..15, affine) turns every later commitment into a sum of
// <= 64k table entries, reduced by a block-wide tree: no doublings, one launch for a whole batch.
#include <string.h>

#include "common.cuh"
#include "host_fp.h"

namespace bp {

static const int FB_WINDOWS = 64;      // 4-bit unsigned windows cover 256 bits
static const int FB_DIGITS = 15;

}  // namespace bp

struct bpgpu_fixed_bases {
  bpgpu_ctx* ctx;
  size_t k;
  void* table;                   // Affine[k][64][15]
  std::vector<uint8_t> key;      // the k bases as X||Y bytes (cache key)
};

namespace bp {

// pow2[j][w] = 2^(4w) * B_j : one thread per base, a serial chain of 252 doublings (one-off)
template <class Fq>
__global__ void k_fb_pow2(const Affine<Fq>* __restrict__ bases, int k, XYZZ<Fq>* __restrict__ pow2) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  XYZZ<Fq> p = XYZZ<Fq>::from_affine(load_vec(bases + j));
  for (int w = 0; w < FB_WINDOWS; w++) {
    store_vec(pow2 + (size_t)j * FB_WINDOWS + w, p);
    if (w + 1 < FB_WINDOWS) for (int b = 0; b < 4; b++) p.dbl();
  }
}

// table[j][w][d-1] = d * pow2[j][w], normalised to affine: one thread per (j, w)
template <class Fq>
__global__ void __launch_bounds__(64) k_fb_multiples(const XYZZ<Fq>* __restrict__ pow2, int total, Affine<Fq>* __restrict__ table) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const XYZZ<Fq> base = load_vec(pow2 + t);
  XYZZ<Fq> acc = base;
  for (int d = 1; d <= FB_DIGITS; d++) {
    store_vec(table + (size_t)t * FB_DIGITS + (d - 1), acc.to_affine());
    if (d < FB_DIGITS) acc.add(base);
  }
}

// one block of 64 threads per commitment: thread w sums the k table entries of window w, then a block tree
template <class Fq>
__global__ void __launch_bounds__(64) k_fb_commit(const Affine<Fq>* __restrict__ table, int k, const ScalarInt* __restrict__ scalars,
                                                  XYZZ<Fq>* __restrict__ out) {
  __shared__ __align__(16) unsigned char smraw[64 * sizeof(XYZZ<Fq>)];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(smraw);
  const int w = threadIdx.x;
  const size_t inst = blockIdx.x;
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  for (int j = 0; j < k; j++) {
    const uint32_t limb = scalars[inst * k + j].v[w >> 3];
    const uint32_t d = (limb >> ((w & 7) * 4)) & 15u;
    if (d) acc.madd(load_vec_ro(table + ((size_t)j * FB_WINDOWS + w) * FB_DIGITS + (d - 1)));
  }
  store_vec(sm + w, acc);
  __syncthreads();
  for (int o = 32; o > 0; o >>= 1) {
    if (w < o) {
      XYZZ<Fq> a = load_vec(sm + w), b = load_vec(sm + w + o);
      a.add(b);
      store_vec(sm + w, a);
    }
    __syncthreads();
  }
  if (w == 0) store_vec(out + inst, load_vec(sm));
}

template <class Curve>
static int fb_build(bpgpu_fixed_bases* fb, const uint8_t* bases_xy) {
  using Fq = typename Curve::Fq;
  bpgpu_ctx* ctx = fb->ctx;
  const int k = (int)fb->k;
  void* d_bases = nullptr;
  void* d_pow2 = nullptr;
  BP_CUDA_OK(cudaMalloc(&d_bases, k * sizeof(Affine<Fq>)));
  BP_CUDA_OK(cudaMalloc(&d_pow2, (size_t)k * FB_WINDOWS * sizeof(XYZZ<Fq>)));
  BP_CUDA_OK(cudaMalloc(&fb->table, (size_t)k * FB_WINDOWS * FB_DIGITS * sizeof(Affine<Fq>)));
  int rc = points_from_host<Curve>(ctx, bases_xy, k, d_bases);
  if (!rc) {
    k_fb_pow2<Fq><<<(k + 31) / 32, 32, 0, ctx->stream>>>((const Affine<Fq>*)d_bases, k, (XYZZ<Fq>*)d_pow2);
    const int total = k * FB_WINDOWS;
    k_fb_multiples<Fq><<<(total + 63) / 64, 64, 0, ctx->stream>>>((const XYZZ<Fq>*)d_pow2, total, (Affine<Fq>*)fb->table);
    ctx->launches += 2;
    rc = launch_check(ctx, "fixed_bases build");
  }
  if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = BPGPU_E_CUDA;
  cudaFree(d_bases);
  cudaFree(d_pow2);
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

template <class Curve>
static int fb_commit(bpgpu_fixed_bases* fb, const uint8_t* scalars_be, size_t count, uint8_t* out_xy) {
  using Fq = typename Curve::Fq;
  using FqParams = typename std::conditional<Curve::ID == BPGPU_BLS12_381, BlsFq, BnFq>::type;
  bpgpu_ctx* ctx = fb->ctx;
  const size_t ns = count * fb->k;
  int rc;
  if ((rc = ctx->msm_c.reserve(ns * 32 + 32))) return rc;
  if ((rc = ctx->msm_e.reserve(count * sizeof(XYZZ<Fq>) + 32))) return rc;
  if ((rc = scalars_from_host<Curve>(ctx, scalars_be, ns, 0, ctx->msm_c.p))) return rc;
  k_fb_commit<Fq><<<(unsigned)count, 64, 0, ctx->stream>>>((const Affine<Fq>*)fb->table, (int)fb->k, (const ScalarInt*)ctx->msm_c.p,
                                                         (XYZZ<Fq>*)ctx->msm_e.p);
  ctx->launches++;
  if ((rc = launch_check(ctx, "k_fb_commit"))) return rc;
  const size_t bytes = count * sizeof(XYZZ<Fq>);
  std::vector<uint8_t> big;
  uint8_t* stage = ctx->pinned;
  if (bytes > ctx->pinned_cap / 2) { big.resize(bytes); stage = big.data(); }
  BP_CUDA_OK(cudaMemcpyAsync(stage, ctx->msm_e.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  normalise_batch_host<FqParams>(stage, count, Curve::MODBYTES, out_xy);
  return BPGPU_OK;
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
