// Lock-step proving of a BATCH of independent R1CS proofs of one circuit shape (n multipliers, one phase).
//
// A single small proof is a chain of ~55 stream operations with a host transcript step between most of them
// (Prover::prove, /root/reference/src/r1cs/prover.rs:322-593, and the rounds of IPP::create_ipp, src/ipp.rs:68-194):
// many contexts in parallel saturate the driver's launch path at a few thousand proofs per second while the GPU idles.
// Here B proofs advance TOGETHER: every device stage is one launch (or one launch + one table-sum launch) for all B
// proofs, the host runs the B transcripts between stages.  Per proof the arithmetic is exactly the single-proof path's
// (same kernels' formulas, same tables), so proof i of a batch is byte-identical to the proof a single call produces.
//
//   commit3   : witness upload, s_L / s_R drawn on the device (one counter-mode stream per proof), A_I, A_O, S
//   polys     : l(x), r(x) coefficient vectors and t_1..t_6                      (prover.rs:458-488)
//   eval      : l_vec, r_vec, G_factors, H_factors -> IPP state                  (prover.rs:524-563)
//   ipp_round : [fold] + L / R scalar rows + cross products, then all 2B sums    (ipp.rs:68-194)
//   ipp_finish: last fold, a and b of every proof
#include <string.h>

#include "batchsum.cuh"
#include "host_fp.h"

struct bpgpu_pbatch {
  bpgpu_ctx* ctx;
  size_t B, n, N;
  bp::FixedRuns runs1, runs2;        // [G[..n) | H[..n) | h] and [G[..N) | H[..N) | g]
  void* mem;                         // one allocation for everything below
  void *W, *S, *blind, *rows1, *wts, *ytab, *polys, *tout, *params, *vecs, *rows2, *uv, *about, *sums, *keys, *ctr0, *parts;
  // a row of F terms is summed by ceil(F / 512) blocks: long rows need more than one block to fill the machine, but every
  // block ends in a 7-level tree, so its 512 terms (32 table additions per thread on average) should outweigh that
  static uint32_t splits_for(size_t F) { return (uint32_t)((F + 511) / 512); }
  size_t n_dev;                      // IPP vector length currently on the device
  bool started;
};

namespace bp {

template <class Fr>
__device__ __forceinline__ Fr pb_pow_tab(const Fr* pw, uint32_t e) {
  Fr acc = Fr::one();
  for (int k = 0; e; k++, e >>= 1)
    if (e & 1) acc = acc * pw[k];
  return acc;
}

// scalar rows of A_I = <a_L,G> + <a_R,H> + i_b h ; A_O = <a_O,G> + o_b h ; S = <s_L,G> + <s_R,H> + s_b h  (prover.rs:347-362)
template <class Fr>
__global__ void __launch_bounds__(128) k_pb_pack3(uint32_t B, uint32_t n, const Fr* __restrict__ W, const Fr* __restrict__ S,
                                                  const Fr* __restrict__ blind, Fr* __restrict__ rows) {
  const uint32_t F = 2 * n + 1;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)3 * B * F) return;
  const uint32_t r = (uint32_t)(t / F), c = (uint32_t)(t - (size_t)r * F);
  const uint32_t b = r / 3, k = r - 3 * b;
  const Fr* w = W + (size_t)b * 3 * n;
  const Fr* s = S + (size_t)b * 2 * n;
  Fr v = Fr::zero();
  if (c == 2 * n) v = load_vec(blind + (size_t)b * 3 + k);
  else if (k == 0) v = load_vec(w + c);                          // a_L | a_R are adjacent
  else if (k == 1) { if (c < n) v = load_vec(w + 2 * n + c); }   // a_O, nothing on H
  else v = load_vec(s + c);                                      // s_L | s_R are adjacent
  store_vec(rows + t, v);
}

// prover.rs:469-486 per proof: l1 = a_L + y^-i wR ; r0 = wO - y^i ; r1 = y^i a_R + wL ; r3 = y^i s_R
template <class Fr>
__global__ void __launch_bounds__(128) k_pb_polys(uint32_t n, const Fr* __restrict__ ytab, const Fr* __restrict__ W, const Fr* __restrict__ S,
                                                  const Fr* __restrict__ wts, Fr* __restrict__ polys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= n) return;
  const Fr* tab = ytab + (size_t)b * 64;
  const Fr *w = W + (size_t)b * 3 * n, *s = S + (size_t)b * 2 * n, *q = wts + (size_t)b * 3 * n;
  Fr* o = polys + (size_t)b * 4 * n;
  const Fr yi = pb_pow_tab(tab, i), yinv = pb_pow_tab(tab + 32, i);
  store_vec(o + i, load_vec(w + i) + yinv * load_vec(q + n + i));                    // l1
  store_vec(o + n + i, load_vec(q + 2 * n + i) - yi);                                // r0
  store_vec(o + 2 * n + i, yi * load_vec(w + n + i) + load_vec(q + i));              // r1
  store_vec(o + 3 * n + i, yi * load_vec(s + n + i));                                // r3
}

// VecPoly3::special_inner_product (vector_poly.rs:79-97), one block per proof
template <class Fr>
__global__ void __launch_bounds__(128) k_pb_tpoly(uint32_t n, const Fr* __restrict__ W, const Fr* __restrict__ S, const Fr* __restrict__ polys,
                                                  Fr* __restrict__ tout) {
  __shared__ __align__(16) unsigned char smraw[128 * sizeof(Fr)];
  Fr* sm = reinterpret_cast<Fr*>(smraw);
  const uint32_t b = blockIdx.x;
  const Fr *w = W + (size_t)b * 3 * n, *s = S + (size_t)b * 2 * n, *p = polys + (size_t)b * 4 * n;
  Fr t[6];
  for (int k = 0; k < 6; k++) t[k] = Fr::zero();
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const Fr l1 = load_vec(p + i), l2 = load_vec(w + 2 * n + i), l3 = load_vec(s + i);
    const Fr r0 = load_vec(p + n + i), r1 = load_vec(p + 2 * n + i), r3 = load_vec(p + 3 * n + i);
    t[0] = t[0] + l1 * r0;
    t[1] = t[1] + l1 * r1 + l2 * r0;
    t[2] = t[2] + l2 * r1 + l3 * r0;
    t[3] = t[3] + l1 * r3 + l3 * r1;
    t[4] = t[4] + l2 * r3;
    t[5] = t[5] + l3 * r3;
  }
  for (int k = 0; k < 6; k++) {
    store_vec(sm + threadIdx.x, t[k]);
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) store_vec(sm + threadIdx.x, load_vec(sm + threadIdx.x) + load_vec(sm + threadIdx.x + o));
      __syncthreads();
    }
    if (threadIdx.x == 0) store_vec(tout + (size_t)b * 6 + k, load_vec(sm));
    __syncthreads();
  }
}

// prover.rs:524-535,552-563 per proof -> IPP state a = l_vec, b = r_vec, sG = G_factors, sH = H_factors (n1 = n: one phase)
template <class Fr>
__global__ void __launch_bounds__(128) k_pb_eval(uint32_t n, uint32_t N, const Fr* __restrict__ ytab, const Fr* __restrict__ params,
                                                 const Fr* __restrict__ W, const Fr* __restrict__ S, const Fr* __restrict__ polys,
                                                 Fr* __restrict__ vecs) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= N) return;
  const Fr* tab = ytab + (size_t)b * 64;
  const Fr x = load_vec(params + (size_t)b * 4), u = load_vec(params + (size_t)b * 4 + 1);
  const Fr *w = W + (size_t)b * 3 * n, *s = S + (size_t)b * 2 * n, *p = polys + (size_t)b * 4 * n;
  Fr* v = vecs + (size_t)b * 4 * N;
  if (i < n) {
    store_vec(v + i, x * (load_vec(p + i) + x * (load_vec(w + 2 * n + i) + x * load_vec(s + i))));
    store_vec(v + N + i, load_vec(p + n + i) + x * (load_vec(p + 2 * n + i) + x * (x * load_vec(p + 3 * n + i))));
  } else {
    store_vec(v + i, Fr::zero());
    store_vec(v + N + i, pb_pow_tab(tab, i).neg());
  }
  const Fr gf = i < n ? Fr::one() : u;
  store_vec(v + 2 * N + i, gf);
  store_vec(v + 3 * N + i, pb_pow_tab(tab + 32, i) * gf);
}

// ytab[b] = y^(2^k) [32] | y^-(2^k) [32] from yy[b] = {y, y^-1}: two threads per proof, 31 squarings each
template <class Fr>
__global__ void __launch_bounds__(128) k_pb_ytab(uint32_t B, const Fr* __restrict__ yy, Fr* __restrict__ ytab) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * B) return;
  Fr cur = load_vec(yy + t);
  Fr* out = ytab + (size_t)(t >> 1) * 64 + (t & 1) * 32;
  for (int k = 0; k < 32; k++) { store_vec(out + k, cur); cur = cur.sqr(); }
}

template <class Fr>
__device__ __forceinline__ void pb_fold_element(uint32_t i, uint32_t n_cur, const Fr& u, const Fr& ui, Fr* a, Fr* b, Fr* sG, Fr* sH, Fr* ab_out) {
  const uint32_t half = n_cur >> 1;
  const uint32_t p = i & (n_cur - 1);
  Fr g = load_vec(sG + i), h = load_vec(sH + i);
  if (p < half) { g = g * ui; h = h * u; } else { g = g * u; h = h * ui; }
  store_vec(sG + i, g);
  store_vec(sH + i, h);
  if (i < half) {
    const Fr al = load_vec(a + i), ar = load_vec(a + i + half), bl = load_vec(b + i), br = load_vec(b + i + half);
    const Fr na = al * u + ui * ar, nb = bl * ui + u * br;
    store_vec(a + i, na);
    store_vec(b + i, nb);
    if (half == 1 && ab_out) { store_vec(ab_out, na); store_vec(ab_out + 1, nb); }
  }
}

// one block per proof: [fold by (u, u^-1)] -> scalar rows of L and R over [G | H | g] (2N+1 each, zero where a base is
// not used this round) -> cross products c_L, c_R times w (Q = w * g)        (ipp.rs:77-104,115-130,145-188)
template <class Fr>
__global__ void __launch_bounds__(256) k_pb_ipp_round(uint32_t N, uint32_t n_in, int do_fold, const Fr* __restrict__ uv,
                                                      const Fr* __restrict__ params, Fr* __restrict__ vecs, Fr* __restrict__ rows) {
  __shared__ __align__(16) unsigned char smraw[256 * sizeof(Fr)];
  Fr* sm = reinterpret_cast<Fr*>(smraw);
  const uint32_t bq = blockIdx.x;
  Fr* a = vecs + (size_t)bq * 4 * N;
  Fr *b = a + N, *sG = a + 2 * N, *sH = a + 3 * N;
  Fr* sclL = rows + (size_t)(2 * bq) * (2 * N + 1);
  Fr* sclR = sclL + (2 * N + 1);
  uint32_t n_cur = n_in;
  if (do_fold) {
    const Fr u = load_vec(uv + (size_t)bq * 2), ui = load_vec(uv + (size_t)bq * 2 + 1);
    for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) pb_fold_element(i, n_cur, u, ui, a, b, sG, sH, (Fr*)nullptr);
    n_cur >>= 1;
    __syncthreads();
  }
  const uint32_t half = n_cur >> 1;
  const Fr zero = Fr::zero();
  for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
    const uint32_t p = i & (n_cur - 1);
    const Fr g = load_vec(sG + i), h = load_vec(sH + i);
    if (p < half) {                     // left half: H_L takes b_R (L), G_L takes a_R (R)
      store_vec(sclL + i, zero);
      store_vec(sclL + N + i, load_vec(b + p + half) * h);
      store_vec(sclR + i, load_vec(a + p + half) * g);
      store_vec(sclR + N + i, zero);
    } else {                            // right half: G_R takes a_L (L), H_R takes b_L (R)
      store_vec(sclL + i, load_vec(a + p - half) * g);
      store_vec(sclL + N + i, zero);
      store_vec(sclR + i, zero);
      store_vec(sclR + N + i, load_vec(b + p - half) * h);
    }
  }
  Fr cl = Fr::zero(), cr = Fr::zero();
  for (uint32_t j = threadIdx.x; j < half; j += blockDim.x) {
    const Fr al = load_vec(a + j), ar = load_vec(a + j + half), bl = load_vec(b + j), br = load_vec(b + j + half);
    cl = cl + al * br;
    cr = cr + ar * bl;
  }
  const Fr wq = load_vec(params + (size_t)bq * 4 + 2);
  for (int pass = 0; pass < 2; pass++) {
    store_vec(sm + threadIdx.x, pass == 0 ? cl : cr);
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) store_vec(sm + threadIdx.x, load_vec(sm + threadIdx.x) + load_vec(sm + threadIdx.x + o));
      __syncthreads();
    }
    if (threadIdx.x == 0) store_vec((pass == 0 ? sclL : sclR) + 2 * N, load_vec(sm) * wq);
    __syncthreads();
  }
}

template <class Fr>
__global__ void __launch_bounds__(128) k_pb_last_fold(uint32_t N, uint32_t n_in, const Fr* __restrict__ uv, Fr* __restrict__ vecs,
                                                      Fr* __restrict__ about) {
  const uint32_t bq = blockIdx.x;
  Fr* a = vecs + (size_t)bq * 4 * N;
  const Fr u = load_vec(uv + (size_t)bq * 2), ui = load_vec(uv + (size_t)bq * 2 + 1);
  for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) pb_fold_element(i, n_in, u, ui, a, a + N, a + 2 * N, a + 3 * N, about + (size_t)bq * 2);
}

static inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

// rows x F scalar matrix (Montgomery) over `runs` -> rows affine points on the host
template <class Curve>
static int pb_sums(bpgpu_pbatch* pb, const FixedRuns& runs, uint32_t F, size_t rows, const void* d_rows, uint8_t* out_xy) {
  using Fq = typename Curve::Fq;
  bpgpu_ctx* ctx = pb->ctx;
  const uint32_t splits = pb->splits_for(F);
  if (splits == 1) {
    k_batch_fixed<Curve><<<(unsigned)rows, BATCH_FIXED_THREADS, 0, ctx->stream>>>(runs, F, (const typename Curve::Fr*)d_rows, 1, (XYZZ<Fq>*)pb->sums);
  } else {
    k_batch_fixed<Curve><<<dim3((unsigned)rows, splits), BATCH_FIXED_THREADS, 0, ctx->stream>>>(runs, F, (const typename Curve::Fr*)d_rows, 1,
                                                                                               (XYZZ<Fq>*)pb->parts);
    k_batch_fixed_combine<Fq><<<(unsigned)((rows + 63) / 64), 64, 0, ctx->stream>>>((uint32_t)rows, splits, (XYZZ<Fq>*)pb->parts, (XYZZ<Fq>*)pb->sums);
    ctx->launches++;
  }
  ctx->launches++;
  int rc = launch_check(ctx, "pb_sums");
  if (rc) return rc;
  std::vector<uint8_t> host(rows * sizeof(XYZZ<Fq>));
  BP_CUDA_OK(cudaMemcpyAsync(host.data(), pb->sums, host.size(), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA_OK(stream_sync(ctx));
  normalise_points_host(ctx->curve, host.data(), rows, out_xy);
  return BPGPU_OK;
}

// host big-endian scalars -> Montgomery values on the host (same radix as the device), then one copy
template <class Curve>
static int pb_upload_mont(bpgpu_pbatch* pb, const uint8_t* be, size_t count, void* dst) {
  using HF = host::HFp<typename std::conditional<Curve::ID == BPGPU_BLS12_381, BlsFr, BnFr>::type>;
  std::vector<HF> tmp(count);
  for (size_t i = 0; i < count; i++) tmp[i] = HF::from_be(be + i * Curve::MODBYTES, Curve::MODBYTES);
  BP_CUDA_OK(cudaMemcpyAsync(dst, tmp.data(), count * sizeof(HF), cudaMemcpyHostToDevice, pb->ctx->stream));
  BP_CUDA_OK(stream_sync(pb->ctx));                      // tmp is a local
  return BPGPU_OK;
}

template <class Curve>
static int pb_commit3_t(bpgpu_pbatch* pb, const uint8_t* witness_be, const uint8_t* keys, size_t key_len, const uint64_t* ctr0,
                        const uint8_t* blind_be, uint8_t* out_xy) {
  using Fr = typename Curve::Fr;
  bpgpu_ctx* ctx = pb->ctx;
  const size_t B = pb->B, n = pb->n;
  int rc;
  if ((rc = scalars_from_host<Curve>(ctx, witness_be, B * 3 * n, 1, pb->W))) return rc;
  if ((rc = pb_upload_mont<Curve>(pb, blind_be, B * 3, pb->blind))) return rc;
  std::vector<uint8_t> kb(B * 64, 0);
  for (size_t b = 0; b < B; b++) memcpy(kb.data() + b * 64, keys + b * key_len, key_len);
  BP_CUDA_OK(cudaMemcpyAsync(pb->keys, kb.data(), kb.size(), cudaMemcpyHostToDevice, ctx->stream));
  BP_CUDA_OK(cudaMemcpyAsync(pb->ctr0, ctr0, B * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  BP_CUDA_OK(stream_sync(ctx));                          // kb is a local
  if ((rc = fr_random_batch_run<Curve>(ctx, (const uint8_t*)pb->keys, (uint32_t)key_len, (const uint64_t*)pb->ctr0, B, 2 * n, pb->S))) return rc;
  const size_t total = 3 * B * (2 * n + 1);
  k_pb_pack3<Fr><<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>((uint32_t)B, (uint32_t)n, (const Fr*)pb->W, (const Fr*)pb->S,
                                                                          (const Fr*)pb->blind, (Fr*)pb->rows1);
  ctx->launches++;
  return pb_sums<Curve>(pb, pb->runs1, (uint32_t)(2 * n + 1), 3 * B, pb->rows1, out_xy);
}

template <class Curve>
static int pb_polys_t(bpgpu_pbatch* pb, const uint8_t* weights_be, const uint8_t* y_be, uint8_t* t_be) {
  using Fr = typename Curve::Fr;
  using HF = host::HFp<typename std::conditional<Curve::ID == BPGPU_BLS12_381, BlsFr, BnFr>::type>;
  bpgpu_ctx* ctx = pb->ctx;
  const size_t B = pb->B, n = pb->n;
  int rc;
  if ((rc = scalars_from_host<Curve>(ctx, weights_be, B * 3 * n, 1, pb->wts))) return rc;
  if ((rc = pb_upload_mont<Curve>(pb, y_be, B * 2, pb->uv))) return rc;          // {y, y^-1} per proof (uv is free until the IPP)
  k_pb_ytab<Fr><<<(unsigned)((2 * B + 127) / 128), 128, 0, ctx->stream>>>((uint32_t)B, (const Fr*)pb->uv, (Fr*)pb->ytab);
  ctx->launches++;
  k_pb_polys<Fr><<<dim3((unsigned)((n + 127) / 128), (unsigned)B), 128, 0, ctx->stream>>>((uint32_t)n, (const Fr*)pb->ytab, (const Fr*)pb->W,
                                                                                         (const Fr*)pb->S, (const Fr*)pb->wts, (Fr*)pb->polys);
  k_pb_tpoly<Fr><<<(unsigned)B, 128, 0, ctx->stream>>>((uint32_t)n, (const Fr*)pb->W, (const Fr*)pb->S, (const Fr*)pb->polys, (Fr*)pb->tout);
  ctx->launches += 2;
  if ((rc = launch_check(ctx, "pb_polys"))) return rc;
  std::vector<HF> t(B * 6);
  BP_CUDA_OK(cudaMemcpyAsync(t.data(), pb->tout, t.size() * sizeof(HF), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA_OK(stream_sync(ctx));
  for (size_t i = 0; i < B * 6; i++) t[i].to_be(t_be + i * Curve::MODBYTES, Curve::MODBYTES);
  return BPGPU_OK;
}

template <class Curve>
static int pb_eval_t(bpgpu_pbatch* pb, const uint8_t* xuw_be) {
  using Fr = typename Curve::Fr;
  bpgpu_ctx* ctx = pb->ctx;
  const size_t B = pb->B, n = pb->n, N = pb->N;
  // params[b] = {x, u, w, 0}
  std::vector<uint8_t> be(B * 4 * Curve::MODBYTES, 0);
  for (size_t b = 0; b < B; b++) memcpy(be.data() + b * 4 * Curve::MODBYTES, xuw_be + b * 3 * Curve::MODBYTES, 3 * Curve::MODBYTES);
  int rc = pb_upload_mont<Curve>(pb, be.data(), B * 4, pb->params);
  if (rc) return rc;
  k_pb_eval<Fr><<<dim3((unsigned)((N + 127) / 128), (unsigned)B), 128, 0, ctx->stream>>>((uint32_t)n, (uint32_t)N, (const Fr*)pb->ytab,
                                                                                        (const Fr*)pb->params, (const Fr*)pb->W, (const Fr*)pb->S,
                                                                                        (const Fr*)pb->polys, (Fr*)pb->vecs);
  ctx->launches++;
  pb->n_dev = N;
  pb->started = false;
  return launch_check(ctx, "k_pb_eval");
}

template <class Curve>
static int pb_round_t(bpgpu_pbatch* pb, const uint8_t* uv_be, uint8_t* out_xy) {
  using Fr = typename Curve::Fr;
  bpgpu_ctx* ctx = pb->ctx;
  const size_t B = pb->B, N = pb->N;
  int rc;
  const int do_fold = pb->started ? 1 : 0;
  if (do_fold && (rc = pb_upload_mont<Curve>(pb, uv_be, B * 2, pb->uv))) return rc;
  k_pb_ipp_round<Fr><<<(unsigned)B, 256, 0, ctx->stream>>>((uint32_t)N, (uint32_t)pb->n_dev, do_fold, (const Fr*)pb->uv, (const Fr*)pb->params,
                                                          (Fr*)pb->vecs, (Fr*)pb->rows2);
  ctx->launches++;
  if (do_fold) pb->n_dev >>= 1;
  pb->started = true;
  return pb_sums<Curve>(pb, pb->runs2, (uint32_t)(2 * N + 1), 2 * B, pb->rows2, out_xy);
}

template <class Curve>
static int pb_finish_t(bpgpu_pbatch* pb, const uint8_t* uv_be, uint8_t* ab_be) {
  using Fr = typename Curve::Fr;
  using HF = host::HFp<typename std::conditional<Curve::ID == BPGPU_BLS12_381, BlsFr, BnFr>::type>;
  bpgpu_ctx* ctx = pb->ctx;
  const size_t B = pb->B, N = pb->N;
  int rc;
  std::vector<HF> ab(B * 2);
  if (N == 1) {                                          // no round ran: a, b are l_vec[0], r_vec[0]
    for (size_t b = 0; b < B; b++) {
      BP_CUDA_OK(cudaMemcpyAsync(&ab[2 * b], (const Fr*)pb->vecs + b * 4, sizeof(HF), cudaMemcpyDeviceToHost, ctx->stream));
      BP_CUDA_OK(cudaMemcpyAsync(&ab[2 * b + 1], (const Fr*)pb->vecs + b * 4 + 1, sizeof(HF), cudaMemcpyDeviceToHost, ctx->stream));
    }
  } else {
    if (pb->n_dev != 2 || !pb->started) return BPGPU_E_ARG;
    if ((rc = pb_upload_mont<Curve>(pb, uv_be, B * 2, pb->uv))) return rc;
    k_pb_last_fold<Fr><<<(unsigned)B, 128, 0, ctx->stream>>>((uint32_t)N, 2u, (const Fr*)pb->uv, (Fr*)pb->vecs, (Fr*)pb->about);
    ctx->launches++;
    if ((rc = launch_check(ctx, "k_pb_last_fold"))) return rc;
    BP_CUDA_OK(cudaMemcpyAsync(ab.data(), pb->about, ab.size() * sizeof(HF), cudaMemcpyDeviceToHost, ctx->stream));
    pb->n_dev = 1;
  }
  BP_CUDA_OK(stream_sync(ctx));
  for (size_t i = 0; i < B * 2; i++) ab[i].to_be(ab_be + i * Curve::MODBYTES, Curve::MODBYTES);
  return BPGPU_OK;
}

}  // namespace bp

using namespace bp;

extern "C" {

int bpgpu_pbatch_create(bpgpu_ctx* ctx, bpgpu_points* G, bpgpu_points* H, const uint8_t* g_xy, const uint8_t* h_xy, size_t batch, size_t n,
                        bpgpu_pbatch** out) {
  if (!ctx || !G || !H || !g_xy || !h_xy || !out || !batch || !n) return BPGPU_E_ARG;
  *out = nullptr;
  size_t N = 1;
  while (N < n) N <<= 1;
  if (G->n < N || H->n < N) return BPGPU_E_GENS_LEN;                   // InvalidGeneratorsLength (prover.rs:332-334,379-381)
  if (batch * (2 * N + 1) >= (1ull << 31) || 2 * N + 1 > (1u << 20)) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = bpgpu_points_precompute(ctx, G)) || (rc = bpgpu_points_precompute(ctx, H))) return rc;
  uint8_t pair[4 * 48];
  const size_t pbytes = 2 * (size_t)bpgpu_modbytes(ctx->curve);
  memcpy(pair, g_xy, pbytes);
  memcpy(pair + pbytes, h_xy, pbytes);
  bpgpu_fixed_bases* fb = nullptr;
  if ((rc = bpgpu_fixed_bases_get(ctx, pair, 2, &fb))) return rc;      // tables of the Pedersen pair (cached in the ctx)
  const void *tg = fixed_table_lookup(ctx, g_xy), *th = fixed_table_lookup(ctx, h_xy);
  if (!tg || !th) return BPGPU_E_ARG;
  bpgpu_pbatch* pb = new (std::nothrow) bpgpu_pbatch();
  if (!pb) return BPGPU_E_CUDA;
  memset(pb, 0, sizeof *pb);
  pb->ctx = ctx; pb->B = batch; pb->n = n; pb->N = N;
  pb->runs1.nruns = 3; pb->runs1.table[0] = G->table; pb->runs1.table[1] = H->table; pb->runs1.table[2] = th;
  pb->runs1.start[0] = 0; pb->runs1.start[1] = (uint32_t)n; pb->runs1.start[2] = (uint32_t)(2 * n); pb->runs1.start[3] = (uint32_t)(2 * n + 1);
  pb->runs2.nruns = 3; pb->runs2.table[0] = G->table; pb->runs2.table[1] = H->table; pb->runs2.table[2] = tg;
  pb->runs2.start[0] = 0; pb->runs2.start[1] = (uint32_t)N; pb->runs2.start[2] = (uint32_t)(2 * N); pb->runs2.start[3] = (uint32_t)(2 * N + 1);
  const size_t fr = 32, xz = ctx->curve == BPGPU_BLS12_381 ? sizeof(XYZZ<Bls::Fq>) : sizeof(XYZZ<Bn::Fq>);
  const size_t B = batch;
  const size_t sizes[17] = {B * 3 * n * fr, B * 2 * n * fr, B * 3 * fr, 3 * B * (2 * n + 1) * fr, B * 3 * n * fr, B * 64 * fr, B * 4 * n * fr, B * 6 * fr,
                            B * 4 * fr, B * 4 * N * fr, 2 * B * (2 * N + 1) * fr, B * 2 * fr, B * 2 * fr, 3 * B * xz, B * 64, B * 8,
                            3 * B * xz * bpgpu_pbatch::splits_for(2 * N + 1)};
  size_t total = 0;
  for (size_t s : sizes) total += up256(s);
  if (dev_alloc(ctx, &pb->mem, total) != cudaSuccess) { delete pb; return BPGPU_E_CUDA; }   // stream-ordered pool: no driver round trip
  void** slots[17] = {&pb->W, &pb->S, &pb->blind, &pb->rows1, &pb->wts, &pb->ytab, &pb->polys, &pb->tout, &pb->params, &pb->vecs, &pb->rows2,
                      &pb->uv, &pb->about, &pb->sums, &pb->keys, &pb->ctr0, &pb->parts};
  uint8_t* p = (uint8_t*)pb->mem;
  for (int k = 0; k < 17; k++) { *slots[k] = p; p += up256(sizes[k]); }
  *out = pb;
  return BPGPU_OK;
}

void bpgpu_pbatch_free(bpgpu_pbatch* pb) {
  if (!pb) return;
  cudaSetDevice(pb->ctx->device);
  if (pb->mem) dev_free(pb->ctx, pb->mem);
  delete pb;
}

#define PB_DISPATCH(pb, CALL) ((pb)->ctx->curve == BPGPU_BLS12_381 ? CALL(Bls) : CALL(Bn))

int bpgpu_pbatch_commit3(bpgpu_pbatch* pb, const uint8_t* witness_be, const uint8_t* keys, size_t key_len, const uint64_t* ctr0,
                         const uint8_t* blind_be, uint8_t* out_xy) {
  if (!pb || !witness_be || !keys || !ctr0 || !blind_be || !out_xy || key_len > 64) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(pb->ctx->device));
#define CALL(C) pb_commit3_t<C>(pb, witness_be, keys, key_len, ctr0, blind_be, out_xy)
  return PB_DISPATCH(pb, CALL);
#undef CALL
}

int bpgpu_pbatch_polys(bpgpu_pbatch* pb, const uint8_t* weights_be, const uint8_t* y_be, uint8_t* t_be) {
  if (!pb || !weights_be || !y_be || !t_be) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(pb->ctx->device));
#define CALL(C) pb_polys_t<C>(pb, weights_be, y_be, t_be)
  return PB_DISPATCH(pb, CALL);
#undef CALL
}

int bpgpu_pbatch_eval(bpgpu_pbatch* pb, const uint8_t* xuw_be) {
  if (!pb || !xuw_be) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(pb->ctx->device));
#define CALL(C) pb_eval_t<C>(pb, xuw_be)
  return PB_DISPATCH(pb, CALL);
#undef CALL
}

int bpgpu_pbatch_ipp_round(bpgpu_pbatch* pb, const uint8_t* uv_be, uint8_t* out_xy) {
  if (!pb || !out_xy || pb->n_dev < 2 || (pb->started && !uv_be)) return BPGPU_E_ARG;
  if (pb->started && pb->n_dev < 4) return BPGPU_E_ARG;               // nothing left to split after this fold
  BP_CUDA_OK(cudaSetDevice(pb->ctx->device));
#define CALL(C) pb_round_t<C>(pb, uv_be, out_xy)
  return PB_DISPATCH(pb, CALL);
#undef CALL
}

int bpgpu_pbatch_ipp_finish(bpgpu_pbatch* pb, const uint8_t* uv_be, uint8_t* ab_be) {
  if (!pb || !ab_be || (pb->N > 1 && !uv_be)) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(pb->ctx->device));
#define CALL(C) pb_finish_t<C>(pb, uv_be, ab_be)
  return PB_DISPATCH(pb, CALL);
#undef CALL
}

}  // extern "C"
