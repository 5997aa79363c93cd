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
#include <stdlib.h>
#include <string.h>

#include "batchsum.cuh"
#include "host_fp.h"
#include "verify_core.cuh"

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
  // device-transcript mode (bpgpu_pbatch_prove_range): transcripts, blinding draws and proof records live on the device too
  void* dmem;
  size_t dm, dplen;
  void *tr, *vrows, *vblind, *trows, *tbl, *ztab, *values, *dkeys, *dsums, *dproofs, *dcomms, *dstate;
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

// one block per proof: [fold by (u, u^-1)] -> COMPACT scalar rows of L and R (N + 1 terms each: position i contributes
// exactly one base to L and one to R, see k_batch_fixed's ipp_half; the last term is the cross product times w, Q = w * g)
// (ipp.rs:77-104,115-130,145-188)
template <class Fr>
__global__ void __launch_bounds__(256) k_pb_ipp_round(uint32_t N, uint32_t n_in, int do_fold, const Fr* __restrict__ uv,
                                                      const Fr* __restrict__ params, Fr* __restrict__ vecs, Fr* __restrict__ rows) {
  __shared__ __align__(16) unsigned char smraw[256 * sizeof(Fr)];
  Fr* sm = reinterpret_cast<Fr*>(smraw);
  const uint32_t bq = blockIdx.x;
  Fr* a = vecs + (size_t)bq * 4 * N;
  Fr *b = a + N, *sG = a + 2 * N, *sH = a + 3 * N;
  Fr* sclL = rows + (size_t)(2 * bq) * (N + 1);
  Fr* sclR = sclL + (N + 1);
  uint32_t n_cur = n_in;
  if (do_fold) {
    const Fr u = load_vec(uv + (size_t)bq * 2), ui = load_vec(uv + (size_t)bq * 2 + 1);
    for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) pb_fold_element(i, n_cur, u, ui, a, b, sG, sH, (Fr*)nullptr);
    n_cur >>= 1;
    __syncthreads();
  }
  const uint32_t half = n_cur >> 1;
  for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
    const uint32_t p = i & (n_cur - 1);
    const Fr g = load_vec(sG + i), h = load_vec(sH + i);
    if (p < half) {                     // left half: H_L takes b_R (L), G_L takes a_R (R)
      store_vec(sclL + i, load_vec(b + p + half) * h);
      store_vec(sclR + i, load_vec(a + p + half) * g);
    } else {                            // right half: G_R takes a_L (L), H_R takes b_L (R)
      store_vec(sclL + i, load_vec(a + p - half) * g);
      store_vec(sclR + i, load_vec(b + p - half) * h);
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
    if (threadIdx.x == 0) store_vec((pass == 0 ? sclL : sclR) + N, load_vec(sm) * wq);
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

// rows x F scalar matrix (Montgomery) over `runs` -> rows XYZZ sums at d_out (device), no synchronisation
template <class Curve>
static int pb_sums_launch(bpgpu_pbatch* pb, const FixedRuns& runs, uint32_t F, size_t rows, const void* d_rows, void* d_out, uint32_t ipp_half = 0) {
  using Fq = typename Curve::Fq;
  bpgpu_ctx* ctx = pb->ctx;
  const uint32_t splits = pb->splits_for(F);
  static const long warp_rows = getenv("BPGPU_WARP_ROWS") ? atol(getenv("BPGPU_WARP_ROWS")) : 2048;   // tuning override
  if (splits == 1 && (long)rows >= warp_rows && F * 8 >= 256) {
    // thousands of rows: one warp per row keeps every lane adding for longer (batchsum.cuh)
    k_batch_fixed_warp<Curve><<<(unsigned)((rows + 3) / 4), 128, 0, ctx->stream>>>(runs, F, (uint32_t)rows, (const typename Curve::Fr*)d_rows, 1,
                                                                                  (XYZZ<Fq>*)d_out, ipp_half);
  } else if (splits == 1) {
    k_batch_fixed<Curve><<<(unsigned)rows, BATCH_FIXED_THREADS, 0, ctx->stream>>>(runs, F, (const typename Curve::Fr*)d_rows, 1, (XYZZ<Fq>*)d_out,
                                                                                 ipp_half);
  } else {
    k_batch_fixed<Curve><<<dim3((unsigned)rows, splits), BATCH_FIXED_THREADS, 0, ctx->stream>>>(runs, F, (const typename Curve::Fr*)d_rows, 1,
                                                                                               (XYZZ<Fq>*)pb->parts, ipp_half);
    k_batch_fixed_combine<Fq><<<(unsigned)((rows + 63) / 64), 64, 0, ctx->stream>>>((uint32_t)rows, splits, (XYZZ<Fq>*)pb->parts, (XYZZ<Fq>*)d_out);
    ctx->launches++;
  }
  ctx->launches++;
  return launch_check(ctx, "pb_sums");
}

// the same sums as affine points on the host (one D2H, one shared inversion there)
template <class Curve>
static int pb_sums(bpgpu_pbatch* pb, const FixedRuns& runs, uint32_t F, size_t rows, const void* d_rows, uint8_t* out_xy, uint32_t ipp_half = 0) {
  using Fq = typename Curve::Fq;
  bpgpu_ctx* ctx = pb->ctx;
  int rc = pb_sums_launch<Curve>(pb, runs, F, rows, d_rows, pb->sums, ipp_half);
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
  return pb_sums<Curve>(pb, pb->runs2, (uint32_t)(N + 1), 2 * B, pb->rows2, out_xy, (uint32_t)(pb->n_dev >> 1));
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


// ======================================================================================================================
// Device-transcript mode (for SLABS): the B transcripts, the blinding draws, the flattened
// constraints and the affine normalisation of every commitment run on the device as well, so a slab of proofs is ~45
// launches with NO host round trip between the upload of the values and the download of the finished proof records.
// (For ONE proof the host transcript stays: a normalisation is a dependent field inversion, ~0.5 ms on a GPU thread against
// 25 us on a host core; for a slab it is one inversion per 4 points on thousands of threads at once.)
// Byte-exactness: same draws (host/curve.hpp Rng positions), same transcript (verify_core.cuh tr_stage_*), same group
// elements as gen_proof_of_positive_nums -- tests/test_gpu_host.py::test_lock_step_* compare the records byte for byte.
// ======================================================================================================================

// per proof: transcript start, the value blindings and i_b, o_b, s_b (draws 0 .. m+2 of the proof's stream), the rows of the
// V commitments, the identity points A_I2 = A_O2 = S2 of a one-phase proof (prover.rs:429)
template <class Curve>
__global__ void __launch_bounds__(64) k_pd_init(uint32_t B, uint32_t m, const uint8_t* __restrict__ state0, const uint64_t* __restrict__ values,
                                                const uint8_t* __restrict__ keys, uint32_t klen, StrobeHD* __restrict__ tr,
                                                typename Curve::Fr* __restrict__ vrows, typename Curve::Fr* __restrict__ vblind,
                                                typename Curve::Fr* __restrict__ blind3, uint64_t* __restrict__ ctr0,
                                                uint8_t* __restrict__ proofs, uint32_t plen) {
  using Fr = typename Curve::Fr;
  using PL = ProofLayout<Curve>;
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  StrobeHD t;
  t.load(state0);
  tr[b] = t;
  const uint8_t* key = keys + (size_t)b * 64;
  for (uint32_t j = 0; j < m; j++) {
    const uint64_t v = values[(size_t)b * m + j];
    Fr fv = Fr::zero();
    fv.v[0] = (uint32_t)v; fv.v[1] = (uint32_t)(v >> 32);
    const Fr bl = fr_stream_draw<Curve>(key, klen, j);
    vrows[((size_t)b * m + j) * 2] = fv;                     // canonical integers: k_commit_pair takes digits
    vrows[((size_t)b * m + j) * 2 + 1] = bl.from_mont();
    vblind[(size_t)b * m + j] = bl;
  }
  for (uint32_t k = 0; k < 3; k++) blind3[(size_t)b * 3 + k] = fr_stream_draw<Curve>(key, klen, m + k);   // prover.rs:336-338
  ctr0[b] = m + 3;                                                                                       // s_L, s_R: the next 2n draws
  uint8_t* rec = proofs + (size_t)b * plen;
  for (uint32_t k = 3; k < 6; k++) {
    uint8_t* p = rec + PL::point(k);
    p[0] = 4;
    for (uint32_t i = 1; i < PL::PB; i++) p[i] = 0;
    p[PL::PB - 1] = 1;
  }
}

// positive_no_gadget's assignments (positive_no.rs:18-24): multiplier i of value j = i / bits holds (1 - bit, bit, 0)
template <class Fr>
__global__ void __launch_bounds__(128) k_pd_witness(uint32_t B, uint32_t m, uint32_t bits, const uint64_t* __restrict__ values, Fr* __restrict__ W) {
  const uint32_t n = m * bits;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)B * n) return;
  const uint32_t b = (uint32_t)(t / n), i = (uint32_t)(t - (size_t)b * n);
  const uint32_t j = i / bits, k = i - j * bits;
  const bool bit = (values[(size_t)b * m + j] >> k) & 1;
  Fr* w = W + (size_t)b * 3 * n;
  store_vec(w + i, bit ? Fr::zero() : Fr::one());
  store_vec(w + n + i, bit ? Fr::one() : Fr::zero());
  store_vec(w + 2 * n + i, Fr::zero());
}

// XYZZ sums -> canonical affine bytes at their place in the proof records (or the commitment list): row r goes to
// dst + (r / per) * stride + off + (r % per) * pitch, as [0x04] || X || Y.  PD_NORM points per thread share one inversion.
static const int PD_NORM = 4;
template <class Curve>
__global__ void __launch_bounds__(64) k_pd_normalise(uint32_t rows, const XYZZ<typename Curve::Fq>* __restrict__ src, uint32_t per,
                                                     uint8_t* __restrict__ dst, size_t stride, uint32_t off, uint32_t pitch, int tag) {
  using Fq = typename Curve::Fq;
  constexpr int MB = Curve::MODBYTES;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t r0 = t * PD_NORM;
  if (r0 >= rows) return;
  const uint32_t cnt = min((uint32_t)PD_NORM, rows - r0);
  Fq pre[PD_NORM + 1];
  pre[0] = Fq::one();
  for (uint32_t i = 0; i < cnt; i++) {
    const Fq zzz = load_vec(&src[r0 + i].zzz);
    pre[i + 1] = zzz.is_zero() ? pre[i] : Fq::mulc(pre[i], zzz);
  }
  Fq inv = pre[cnt].inv();
  for (uint32_t i = cnt; i-- > 0;) {
    const uint32_t r = r0 + i;
    uint8_t* o = dst + (size_t)(r / per) * stride + off + (size_t)(r % per) * pitch;
    if (tag) *o++ = 4;
    const XYZZ<Fq> p = load_vec(src + r);
    if (p.is_inf()) {                                   // AMCL's identity (0, 1)
      for (int k = 0; k < 2 * MB; k++) o[k] = 0;
      o[2 * MB - 1] = 1;
      continue;
    }
    const Fq i3 = Fq::mulc(inv, pre[i]);                // 1 / zzz
    inv = Fq::mulc(inv, p.zzz);
    const Fq i1 = Fq::mulc(i3, p.zz);                   // 1 / z
    const Fq ax = Fq::mulc(p.x, Fq::mulc(i1, i1)).from_mont(), ay = Fq::mulc(p.y, i3).from_mont();
    hd_limbs_to_be<Fq::N>(ax.v, MB, o);
    hd_limbs_to_be<Fq::N>(ay.v, MB, o + MB);
  }
}

// V_j, "m", A_I1 A_O1 S1, separator, A_I2 A_O2 S2 -> y, z ; y^-1 ; z^(2^k)
template <class Curve>
__global__ void __launch_bounds__(64) k_pd_tr1(uint32_t B, uint32_t m, StrobeHD* __restrict__ tr, const uint8_t* __restrict__ proofs, uint32_t plen,
                                               const uint8_t* __restrict__ comms, typename Curve::Fr* __restrict__ yy,
                                               typename Curve::Fr* __restrict__ ztab) {
  using Fr = typename Curve::Fr;
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  StrobeHD t = tr[b];
  Fr y, z;
  tr_stage_commitments<Curve>(t, proofs + (size_t)b * plen, comms + (size_t)b * m * 2 * Curve::MODBYTES, m, &y, &z);
  tr[b] = t;
  yy[(size_t)b * 2] = y;
  yy[(size_t)b * 2 + 1] = y.inv();                      // 0 -> 0, as FieldElement::inverse
  hd_square_table(z, ztab + (size_t)b * 32);
}

// flattened_constraints (prover.rs:142-184) per proof: wL | wR | wO from the shared circuit matrix and the proof's z
template <class Fr>
__global__ void __launch_bounds__(128) k_pd_flatten(const __grid_constant__ CircuitDev c, uint32_t B, const Fr* __restrict__ ztab, Fr* __restrict__ wts) {
  const uint32_t rows = 3 * c.n;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)B * rows) return;
  const uint32_t b = (uint32_t)(t / rows), row = (uint32_t)(t - (size_t)b * rows);
  store_vec(wts + t, csr_row_eval(c, row, ztab + (size_t)b * 32));
}

// t_1, t_3, t_4, t_5, t_6 with their blindings (draws m+3+2n .. +4; prover.rs:490-500) as rows over (g, h)
template <class Curve>
__global__ void __launch_bounds__(64) k_pd_tr2(uint32_t B, uint32_t m, uint32_t n, const uint8_t* __restrict__ keys, uint32_t klen,
                                               const typename Curve::Fr* __restrict__ tout, typename Curve::Fr* __restrict__ trows,
                                               typename Curve::Fr* __restrict__ tbl) {
  using Fr = typename Curve::Fr;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * 5) return;
  const uint32_t b = t / 5, k = t - b * 5;
  const int idx[5] = {1, 3, 4, 5, 6};
  const Fr bl = fr_stream_draw<Curve>(keys + (size_t)b * 64, klen, (uint64_t)m + 3 + 2 * (uint64_t)n + k);
  trows[(size_t)t * 2] = tout[(size_t)b * 6 + idx[k] - 1].from_mont();      // canonical integers: k_commit_pair takes digits
  trows[(size_t)t * 2 + 1] = bl.from_mont();
  tbl[t] = bl;
}

// T_i -> u, x ; t_2_blinding = <wV, v_blinding> ; t_x, t_x_blinding, e_blinding (prover.rs:508-541) into the record ; -> w ;
// the inner-product argument's separator (ipp.rs:62)
template <class Curve>
__global__ void __launch_bounds__(64) k_pd_tr3(const __grid_constant__ CircuitDev c, uint32_t B, uint32_t N, StrobeHD* __restrict__ tr,
                                               uint8_t* __restrict__ proofs, uint32_t plen, const typename Curve::Fr* __restrict__ tout,
                                               const typename Curve::Fr* __restrict__ tbl, const typename Curve::Fr* __restrict__ vblind,
                                               const typename Curve::Fr* __restrict__ blind3, const typename Curve::Fr* __restrict__ ztab,
                                               typename Curve::Fr* __restrict__ params) {
  using Fr = typename Curve::Fr;
  using PL = ProofLayout<Curve>;
  constexpr int MB = Curve::MODBYTES;
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  uint8_t* rec = proofs + (size_t)b * plen;
  StrobeHD t = tr[b];
  Fr u, x, w;
  tr_stage_t<Curve>(t, rec, &u, &x);
  Fr tb2 = Fr::zero();                                  // prover.rs:513
  for (uint32_t j = 0; j < c.m; j++) tb2 = tb2 + csr_row_eval(c, 3 * c.n + j, ztab + (size_t)b * 32) * vblind[(size_t)b * c.m + j];
  const Fr* tt = tout + (size_t)b * 6;
  const Fr* bl = tbl + (size_t)b * 5;
  // Poly6::eval (vector_poly.rs:115-119): x * (c1 + x * (c2 + ... + x * c6))
  const Fr t_x = x * (tt[0] + x * (tt[1] + x * (tt[2] + x * (tt[3] + x * (tt[4] + x * tt[5])))));
  const Fr t_xb = x * (bl[0] + x * (tb2 + x * (bl[1] + x * (bl[2] + x * (bl[3] + x * bl[4])))));
  const Fr* b3 = blind3 + (size_t)b * 3;
  const Fr e_b = x * (b3[0] + x * (b3[1] + x * b3[2]));   // prover.rs:537-541 with the second-phase blindings zero
  const Fr sc[3] = {t_x.from_mont(), t_xb.from_mont(), e_b.from_mont()};
  for (int k = 0; k < 3; k++) hd_limbs_to_be<8>(sc[k].v, MB, rec + PL::scalar(k));
  tr_stage_scalars<Curve>(t, rec, N, &w);
  tr[b] = t;
  Fr* p = params + (size_t)b * 4;
  p[0] = x; p[1] = u; p[2] = w; p[3] = Fr::zero();
}

// L_k, R_k -> u_k and its inverse for the next fold (ipp.rs:106-113 / 172-179)
template <class Curve>
__global__ void __launch_bounds__(64) k_pd_round(uint32_t B, uint32_t lg, uint32_t k, StrobeHD* __restrict__ tr, const uint8_t* __restrict__ proofs,
                                                 uint32_t plen, typename Curve::Fr* __restrict__ uv) {
  using Fr = typename Curve::Fr;
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  StrobeHD t = tr[b];
  const Fr u = tr_round<Curve>(t, proofs + (size_t)b * plen, lg, k);
  tr[b] = t;
  uv[(size_t)b * 2] = u;
  uv[(size_t)b * 2 + 1] = u.inv();
}

// a, b of every proof into its record (ipp.rs:196-201)
template <class Curve>
__global__ void __launch_bounds__(64) k_pd_final(uint32_t B, uint32_t lg, const typename Curve::Fr* __restrict__ ab, size_t ab_stride, size_t b_off,
                                                 uint8_t* __restrict__ proofs, uint32_t plen) {
  using Fr = typename Curve::Fr;
  using PL = ProofLayout<Curve>;
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const Fr a = load_vec(ab + (size_t)b * ab_stride).from_mont(), bb = load_vec(ab + (size_t)b * ab_stride + b_off).from_mont();
  uint8_t* rec = proofs + (size_t)b * plen + PL::a(lg);
  hd_limbs_to_be<8>(a.v, Curve::MODBYTES, rec);
  hd_limbs_to_be<8>(bb.v, Curve::MODBYTES, rec + Curve::MODBYTES);
}

template <class Curve>
static int pd_prove_t(bpgpu_pbatch* pb, const CircuitDev& c, const uint64_t* values, uint32_t m, uint32_t bits, const uint8_t* state0,
                      const uint8_t* keys, uint32_t klen, uint8_t* proofs, size_t stride, uint8_t* comms_xy) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  using PL = ProofLayout<Curve>;
  constexpr uint32_t MB = Curve::MODBYTES;
  bpgpu_ctx* ctx = pb->ctx;
  cudaStream_t st = ctx->stream;
  const uint32_t B = (uint32_t)pb->B, n = (uint32_t)pb->n, N = (uint32_t)pb->N;
  uint32_t lg = 0;
  while ((1u << lg) < N) lg++;
  const uint32_t plen = PL::len(lg);
  int rc;
  // ---- device-mode buffers (one allocation, kept with the pbatch)
  if (!pb->dmem || pb->dm != m || pb->dplen != plen) {
    if (pb->dmem) dev_free(ctx, pb->dmem);
    pb->dmem = nullptr;
    const size_t nsum = (size_t)B * (m > 5 ? m : 5);
    const size_t sizes[12] = {(size_t)B * sizeof(StrobeHD), (size_t)B * m * 2 * 32, (size_t)B * m * 32, (size_t)B * 5 * 2 * 32, (size_t)B * 5 * 32,
                              (size_t)B * 32 * 32, (size_t)B * m * 8, (size_t)B * 64, nsum * sizeof(XYZZ<Fq>), (size_t)B * plen,
                              (size_t)B * m * 2 * MB, 256};
    size_t total = 0;
    for (size_t s : sizes) total += up256(s);
    if (dev_alloc(ctx, &pb->dmem, total) != cudaSuccess) return BPGPU_E_CUDA;
    void** slots[12] = {&pb->tr, &pb->vrows, &pb->vblind, &pb->trows, &pb->tbl, &pb->ztab, &pb->values, &pb->dkeys, &pb->dsums, &pb->dproofs,
                        &pb->dcomms, &pb->dstate};
    uint8_t* p = (uint8_t*)pb->dmem;
    for (int k = 0; k < 12; k++) { *slots[k] = p; p += up256(sizes[k]); }
    pb->dm = m; pb->dplen = plen;
  }
  StrobeHD* tr = (StrobeHD*)pb->tr;
  uint8_t* d_proofs = (uint8_t*)pb->dproofs;
  uint8_t* d_comms = (uint8_t*)pb->dcomms;
  const uint8_t* d_keys = (const uint8_t*)pb->dkeys;
  // keys arrive packed by klen; the kernels want a stride of 64
  std::vector<uint8_t> kb((size_t)B * 64, 0);
  for (uint32_t b = 0; b < B; b++) memcpy(kb.data() + (size_t)b * 64, keys + (size_t)b * klen, klen);
  BP_CUDA_OK(cudaMemcpyAsync(pb->dkeys, kb.data(), kb.size(), cudaMemcpyHostToDevice, st));
  BP_CUDA_OK(cudaMemcpyAsync(pb->values, values, (size_t)B * m * 8, cudaMemcpyHostToDevice, st));
  BP_CUDA_OK(cudaMemcpyAsync(pb->dstate, state0, MERLIN_STATE_BYTES, cudaMemcpyHostToDevice, st));
  const void *tg = pb->runs2.table[2], *th = pb->runs1.table[2];       // window tables of the Pedersen pair (g, h) of V_j and T_i
  const unsigned gB = (B + 63) / 64;
  auto normalise = [&](uint32_t rows, const void* src, uint32_t per, uint8_t* dst, size_t dstride, uint32_t off, uint32_t pitch, int tag) {
    const uint32_t threads = (rows + PD_NORM - 1) / PD_NORM;
    k_pd_normalise<Curve><<<(threads + 63) / 64, 64, 0, st>>>(rows, (const XYZZ<Fq>*)src, per, dst, dstride, off, pitch, tag);
    ctx->launches++;
  };
  // ---- V commitments (prover.rs:119-129), witness, first-phase blindings
  k_pd_init<Curve><<<gB, 64, 0, st>>>(B, m, (const uint8_t*)pb->dstate, (const uint64_t*)pb->values, d_keys, klen, tr, (Fr*)pb->vrows,
                                      (Fr*)pb->vblind, (Fr*)pb->blind, (uint64_t*)pb->ctr0, d_proofs, plen);
  k_pd_witness<Fr><<<(unsigned)(((size_t)B * n + 127) / 128), 128, 0, st>>>(B, m, bits, (const uint64_t*)pb->values, (Fr*)pb->W);
  ctx->launches += 2;
  k_commit_pair<Curve><<<(B * m + 3) / 4, 128, 0, st>>>(B * m, tg, th, (const Fr*)pb->vrows, (XYZZ<Fq>*)pb->dsums);
  ctx->launches++;
  normalise(B * m, pb->dsums, m, d_comms, (size_t)m * 2 * MB, 0, 2 * MB, 0);
  // ---- A_I1, A_O1, S1 (prover.rs:340-362)
  if ((rc = fr_random_batch_run<Curve>(ctx, d_keys, klen, (const uint64_t*)pb->ctr0, B, 2 * (size_t)n, pb->S))) return rc;
  {
    const size_t total = (size_t)3 * B * (2 * n + 1);
    k_pb_pack3<Fr><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(B, n, (const Fr*)pb->W, (const Fr*)pb->S, (const Fr*)pb->blind, (Fr*)pb->rows1);
    ctx->launches++;
  }
  if ((rc = pb_sums_launch<Curve>(pb, pb->runs1, 2 * n + 1, (size_t)3 * B, pb->rows1, pb->sums))) return rc;
  normalise(3 * B, pb->sums, 3, d_proofs, plen, PL::point(0), PL::PB, 1);
  // ---- y, z; flattened constraints; l(x), r(x), t_1..t_6 (prover.rs:438-488)
  k_pd_tr1<Curve><<<gB, 64, 0, st>>>(B, m, tr, d_proofs, plen, d_comms, (Fr*)pb->uv, (Fr*)pb->ztab);
  k_pb_ytab<Fr><<<(2 * B + 127) / 128, 128, 0, st>>>(B, (const Fr*)pb->uv, (Fr*)pb->ytab);
  k_pd_flatten<Fr><<<(unsigned)(((size_t)B * 3 * n + 127) / 128), 128, 0, st>>>(c, B, (const Fr*)pb->ztab, (Fr*)pb->wts);
  k_pb_polys<Fr><<<dim3((n + 127) / 128, B), 128, 0, st>>>(n, (const Fr*)pb->ytab, (const Fr*)pb->W, (const Fr*)pb->S, (const Fr*)pb->wts,
                                                           (Fr*)pb->polys);
  k_pb_tpoly<Fr><<<B, 128, 0, st>>>(n, (const Fr*)pb->W, (const Fr*)pb->S, (const Fr*)pb->polys, (Fr*)pb->tout);
  // ---- T_1, T_3, T_4, T_5, T_6 (prover.rs:490-506)
  k_pd_tr2<Curve><<<(B * 5 + 63) / 64, 64, 0, st>>>(B, m, n, d_keys, klen, (const Fr*)pb->tout, (Fr*)pb->trows, (Fr*)pb->tbl);
  ctx->launches += 6;
  k_commit_pair<Curve><<<(B * 5 + 3) / 4, 128, 0, st>>>(B * 5, tg, th, (const Fr*)pb->trows, (XYZZ<Fq>*)pb->dsums);
  ctx->launches++;
  normalise(B * 5, pb->dsums, 5, d_proofs, plen, PL::point(6), PL::PB, 1);
  // ---- u, x, t_x, blindings, w; l_vec, r_vec, factors (prover.rs:508-563)
  k_pd_tr3<Curve><<<gB, 64, 0, st>>>(c, B, N, tr, d_proofs, plen, (const Fr*)pb->tout, (const Fr*)pb->tbl, (const Fr*)pb->vblind,
                                     (const Fr*)pb->blind, (const Fr*)pb->ztab, (Fr*)pb->params);
  k_pb_eval<Fr><<<dim3((N + 127) / 128, B), 128, 0, st>>>(n, N, (const Fr*)pb->ytab, (const Fr*)pb->params, (const Fr*)pb->W, (const Fr*)pb->S,
                                                          (const Fr*)pb->polys, (Fr*)pb->vecs);
  ctx->launches += 2;
  if ((rc = launch_check(ctx, "pd_prove stages"))) return rc;
  // ---- IPP rounds (ipp.rs:68-194)
  uint32_t n_dev = N;
  for (uint32_t k = 0; k < lg; k++) {
    k_pb_ipp_round<Fr><<<B, 256, 0, st>>>(N, n_dev, k ? 1 : 0, (const Fr*)pb->uv, (const Fr*)pb->params, (Fr*)pb->vecs, (Fr*)pb->rows2);
    ctx->launches++;
    if (k) n_dev >>= 1;
    if ((rc = pb_sums_launch<Curve>(pb, pb->runs2, N + 1, (size_t)2 * B, pb->rows2, pb->sums, n_dev >> 1))) return rc;
    // L_k and R_k of a proof sit lg points apart in its record
    normalise(2 * B, pb->sums, 2, d_proofs, plen, PL::L(k), lg * PL::PB, 1);
    k_pd_round<Curve><<<gB, 64, 0, st>>>(B, lg, k, tr, d_proofs, plen, (Fr*)pb->uv);
    ctx->launches++;
  }
  if (lg) {
    k_pb_last_fold<Fr><<<B, 128, 0, st>>>(N, 2u, (const Fr*)pb->uv, (Fr*)pb->vecs, (Fr*)pb->about);
    k_pd_final<Curve><<<gB, 64, 0, st>>>(B, lg, (const Fr*)pb->about, 2, 1, d_proofs, plen);
  } else {                                              // N = 1: a = l_vec[0], b = r_vec[0]
    k_pd_final<Curve><<<gB, 64, 0, st>>>(B, lg, (const Fr*)pb->vecs, 4, 1, d_proofs, plen);
  }
  ctx->launches += 2;
  if ((rc = launch_check(ctx, "pd_prove rounds"))) return rc;
  if (stride == plen) BP_CUDA_OK(cudaMemcpyAsync(proofs, d_proofs, (size_t)B * plen, cudaMemcpyDeviceToHost, st));
  else BP_CUDA_OK(cudaMemcpy2DAsync(proofs, stride, d_proofs, plen, plen, B, cudaMemcpyDeviceToHost, st));
  BP_CUDA_OK(cudaMemcpyAsync(comms_xy, d_comms, (size_t)B * m * 2 * MB, cudaMemcpyDeviceToHost, st));
  BP_CUDA_OK(stream_sync(ctx));                          // kb, proofs and comms are the caller's / locals
  pb->n_dev = 1;
  pb->started = false;
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
  pb->runs1.table16[0] = G->table16; pb->runs1.table16[1] = H->table16; pb->runs1.table16[2] = nullptr;      // wide tables when the caller built them
  pb->runs1.start[0] = 0; pb->runs1.start[1] = (uint32_t)n; pb->runs1.start[2] = (uint32_t)(2 * n); pb->runs1.start[3] = (uint32_t)(2 * n + 1);
  pb->runs2.nruns = 3; pb->runs2.table[0] = G->table; pb->runs2.table[1] = H->table; pb->runs2.table[2] = tg;
  pb->runs2.table16[0] = G->table16; pb->runs2.table16[1] = H->table16; pb->runs2.table16[2] = nullptr;
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
  if (pb->dmem) dev_free(pb->ctx, pb->dmem);
  if (pb->mem) dev_free(pb->ctx, pb->mem);
  delete pb;
}

#define PB_DISPATCH(pb, CALL) ((pb)->ctx->curve == BPGPU_BLS12_381 ? CALL(Bls) : CALL(Bn))

int bpgpu_pbatch_prove_range(bpgpu_pbatch* pb, const bpgpu_circuit* circuit, const uint64_t* values, size_t m, size_t bits,
                             const uint8_t* transcript_state, const uint8_t* keys, size_t key_len, uint8_t* proofs, size_t proof_stride,
                             uint8_t* comms_xy) {
  if (!pb || !circuit || !values || !transcript_state || !keys || !proofs || !comms_xy || !m || !bits || bits > 64 || key_len > 56) return BPGPU_E_ARG;
  if (m * bits != pb->n || circuit->dev.n != pb->n || circuit->dev.m != m) return BPGPU_E_LEN;
  if (circuit->ctx->device != pb->ctx->device || circuit->ctx->curve != pb->ctx->curve) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(pb->ctx->device));
  size_t lg = 0;
  while (((size_t)1 << lg) < pb->N) lg++;
  const size_t mb = (size_t)bpgpu_modbytes(pb->ctx->curve);
  if (proof_stride < 11 * (2 * mb + 1) + 3 * mb + 2 * lg * (2 * mb + 1) + 2 * mb) return BPGPU_E_ARG;
#define CALL(C) pd_prove_t<C>(pb, circuit->dev, values, (uint32_t)m, (uint32_t)bits, transcript_state, keys, (uint32_t)key_len, proofs, proof_stride, comms_xy)
  return PB_DISPATCH(pb, CALL);
#undef CALL
}

// the multiplier assignments of m positive_no gadgets (positive_no.rs:18-24) for ONE proof, built on the device from the values
int bpgpu_range_witness(bpgpu_ctx* ctx, const uint64_t* values, size_t m, size_t bits, bpgpu_scalars** out) {
  if (!ctx || !values || !out || !m || !bits || bits > 64 || m > ((size_t)1 << 22)) return BPGPU_E_ARG;
  *out = nullptr;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t n = m * bits;
  int rc = bpgpu_scalars_alloc(ctx, 3 * n, out);
  if (rc) return rc;
  void* dv = nullptr;
  if (dev_alloc(ctx, &dv, m * sizeof(uint64_t)) != cudaSuccess ||
      cudaMemcpyAsync(dv, values, m * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
    dev_free(ctx, dv); bpgpu_scalars_free(*out); *out = nullptr;
    return BPGPU_E_CUDA;
  }
  if (ctx->curve == BPGPU_BLS12_381)
    k_pd_witness<Bls::Fr><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(1, (uint32_t)m, (uint32_t)bits, (const uint64_t*)dv, (Bls::Fr*)(*out)->d);
  else
    k_pd_witness<Bn::Fr><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(1, (uint32_t)m, (uint32_t)bits, (const uint64_t*)dv, (Bn::Fr*)(*out)->d);
  ctx->launches++;
  dev_free(ctx, dv);
  rc = launch_check(ctx, "k_pd_witness");
  if (rc) { bpgpu_scalars_free(*out); *out = nullptr; }
  return rc;
}

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
