// Batched Fr vector kernels: the FieldElementVector algebra of the reference on device-resident
// vectors (Montgomery form, 8 x u32 per element).
//
// Replaces amcl_wrapper FieldElementVector::{inner_product, hadamard_product, new_vandermonde_vector,
// scaled_by, plus} and FieldElement::batch_invert at the call sites /root/reference/src/ipp.rs:77-82,
// 94-95,145-146,295; src/r1cs/prover.rs:463,472-485,513; src/r1cs/verifier.rs:342-352,416;
// src/utils/vector_poly.rs:79-106.  All results are canonical residues -> bit exact.
#include "common.cuh"
#include "host_fp.h"

namespace bp {

// x^e for e < 2^32 with a table of x^(2^k) (k < 32) in pw[]
template <class Fr>
__device__ __forceinline__ Fr pow_table(const Fr* __restrict__ pw, uint32_t e) {
  Fr acc = Fr::one();
  for (int k = 0; e; k++, e >>= 1)
    if (e & 1) acc = acc * pw[k];
  return acc;
}

template <class Fr>
__global__ void k_fr_vandermonde(const Fr* __restrict__ pw, size_t n, Fr* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) store_vec(out + i, pow_table(pw, (uint32_t)i));
}

// op 0: a*b  1: a+b  2: a-b  3: a*s (s = scalar in sc[0])  4: a*s + b
template <class Fr>
__global__ void k_fr_elementwise(int op, const Fr* __restrict__ a, const Fr* __restrict__ b, const Fr* __restrict__ sc, size_t n,
                                 Fr* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr x = load_vec(a + i), r;
  if (op == 0) r = x * load_vec(b + i);
  else if (op == 1) r = x + load_vec(b + i);
  else if (op == 2) r = x - load_vec(b + i);
  else if (op == 3) r = x * sc[0];
  else r = x * sc[0] + load_vec(b + i);
  store_vec(out + i, r);
}

template <class Fr>
__global__ void k_fr_invert(const Fr* __restrict__ a, size_t n, Fr* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) store_vec(out + i, load_vec(a + i).inv());
}

// block-wide sum of one Fr per thread (blockDim.x == 256); result valid on thread 0
template <class Fr>
__device__ __forceinline__ Fr block_sum_fr(Fr v, Fr* sm) {
  store_vec(sm + threadIdx.x, v);
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) store_vec(sm + threadIdx.x, load_vec(sm + threadIdx.x) + load_vec(sm + threadIdx.x + o));
    __syncthreads();
  }
  Fr r = load_vec(sm);
  __syncthreads();
  return r;
}

// NP inner products <a_k, b_k> of length n in one launch: grid = (blocks, NP); partial sums per block,
// finished by k_fr_dot_finish.  (vector_poly.rs:79-97 needs 9 of them over the same six vectors.)
struct DotPtrs { const void* a[9]; const void* b[9]; };
template <class Fr>
__global__ void __launch_bounds__(256) k_fr_dot_partial(DotPtrs p, size_t n, Fr* __restrict__ partial) {
  __shared__ __align__(16) unsigned char smraw[256 * sizeof(Fr)];
  Fr* sm = reinterpret_cast<Fr*>(smraw);
  const Fr* a = (const Fr*)p.a[blockIdx.y];
  const Fr* b = (const Fr*)p.b[blockIdx.y];
  Fr acc = Fr::zero();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc = acc + load_vec(a + i) * load_vec(b + i);
  Fr tot = block_sum_fr(acc, sm);
  if (threadIdx.x == 0) store_vec(partial + (size_t)blockIdx.y * gridDim.x + blockIdx.x, tot);
}
template <class Fr>
__global__ void __launch_bounds__(256) k_fr_dot_finish(const Fr* __restrict__ partial, int nblocks, Fr* __restrict__ out) {
  __shared__ __align__(16) unsigned char smraw[256 * sizeof(Fr)];
  Fr* sm = reinterpret_cast<Fr*>(smraw);
  Fr acc = Fr::zero();
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) acc = acc + load_vec(partial + (size_t)blockIdx.x * nblocks + i);
  Fr tot = block_sum_fr(acc, sm);
  if (threadIdx.x == 0) store_vec(out + blockIdx.x, tot);
}

// ---------------------------------------------------------------------------------------------
// host helpers
template <class Curve>
static int fr_dots(bpgpu_ctx* ctx, const DotPtrs& p, int np, size_t n, typename Curve::Fr* d_out /* np results, Montgomery */) {
  using Fr = typename Curve::Fr;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 64) blocks = 64;
  if (blocks < 1) blocks = 1;
  int rc = ctx->fr_tmp.reserve((size_t)np * blocks * sizeof(Fr));
  if (rc) return rc;
  k_fr_dot_partial<Fr><<<dim3(blocks, np), 256, 0, ctx->stream>>>(p, n, (Fr*)ctx->fr_tmp.p);
  k_fr_dot_finish<Fr><<<np, 256, 0, ctx->stream>>>((const Fr*)ctx->fr_tmp.p, blocks, d_out);
  ctx->launches += 2;
  return launch_check(ctx, "fr_dots");
}

template <class Curve>
int fr_dots_to_host(bpgpu_ctx* ctx, const DotPtrs& p, int np, size_t n, uint8_t* out_be) {
  using Fr = typename Curve::Fr;
  int rc = ctx->fr_out.reserve(16 * sizeof(Fr));
  if (rc) return rc;
  if ((rc = fr_dots<Curve>(ctx, p, np, n, (Fr*)ctx->fr_out.p))) return rc;
  BP_CUDA_OK(cudaMemcpyAsync(ctx->pinned, ctx->fr_out.p, np * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA_OK(stream_sync(ctx));
  using HF = host::HFp<typename std::conditional<Curve::ID == BPGPU_BLS12_381, BlsFr, BnFr>::type>;
  for (int k = 0; k < np; k++) reinterpret_cast<const HF*>(ctx->pinned)[k].to_be(out_be + (size_t)k * Curve::MODBYTES, Curve::MODBYTES);
  return BPGPU_OK;
}
template int fr_dots_to_host<Bls>(bpgpu_ctx*, const DotPtrs&, int, size_t, uint8_t*);
template int fr_dots_to_host<Bn>(bpgpu_ctx*, const DotPtrs&, int, size_t, uint8_t*);

// upload `cnt` host scalars (big endian) as Montgomery Fr into ctx->fr_args (kernel argument vectors)
template <class Curve>
int fr_args_upload(bpgpu_ctx* ctx, const uint8_t* be, int cnt, typename Curve::Fr** d_out) {
  using Fr = typename Curve::Fr;
  using HF = host::HFp<typename std::conditional<Curve::ID == BPGPU_BLS12_381, BlsFr, BnFr>::type>;
  int rc = ctx->fr_args.reserve((size_t)cnt * sizeof(Fr) + 64);
  if (rc) return rc;
  if ((size_t)cnt * sizeof(Fr) > ctx->pinned_cap / 2) return BPGPU_E_ARG;
  // staging lives in the upper half of the pinned buffer; results use the lower half
  uint8_t* stage = ctx->pinned + ctx->pinned_cap / 2;
  BP_CUDA_OK(stream_sync(ctx));     // previous use of the staging area is complete
  for (int k = 0; k < cnt; k++) reinterpret_cast<HF*>(stage)[k] = HF::from_be(be + (size_t)k * Curve::MODBYTES, Curve::MODBYTES);
  BP_CUDA_OK(cudaMemcpyAsync(ctx->fr_args.p, stage, (size_t)cnt * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  *d_out = (Fr*)ctx->fr_args.p;
  return BPGPU_OK;
}
template int fr_args_upload<Bls>(bpgpu_ctx*, const uint8_t*, int, Bls::Fr**);
template int fr_args_upload<Bn>(bpgpu_ctx*, const uint8_t*, int, Bn::Fr**);

// x^(2^k), k < 32, computed on the host and uploaded (vandermonde / power tables)
template <class Curve>
int fr_pow_table_upload(bpgpu_ctx* ctx, const uint8_t* x_be, Scratch& dst, typename Curve::Fr** d_out) {
  using Fr = typename Curve::Fr;
  using HF = host::HFp<typename std::conditional<Curve::ID == BPGPU_BLS12_381, BlsFr, BnFr>::type>;
  int rc = dst.reserve(32 * sizeof(Fr));
  if (rc) return rc;
  BP_CUDA_OK(stream_sync(ctx));
  HF* stage = reinterpret_cast<HF*>(ctx->pinned + ctx->pinned_cap / 2);
  HF cur = HF::from_be(x_be, Curve::MODBYTES);
  for (int k = 0; k < 32; k++) { stage[k] = cur; cur = cur.sqr(); }
  BP_CUDA_OK(cudaMemcpyAsync(dst.p, stage, 32 * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  BP_CUDA_OK(stream_sync(ctx));
  *d_out = (Fr*)dst.p;
  return BPGPU_OK;
}
template int fr_pow_table_upload<Bls>(bpgpu_ctx*, const uint8_t*, Scratch&, Bls::Fr**);
template int fr_pow_table_upload<Bn>(bpgpu_ctx*, const uint8_t*, Scratch&, Bn::Fr**);

template <class Curve>
static int vandermonde_t(bpgpu_ctx* ctx, const uint8_t* x_be, size_t n, void* d_out) {
  typename Curve::Fr* pw;
  int rc = fr_pow_table_upload<Curve>(ctx, x_be, ctx->fr_pow, &pw);
  if (rc) return rc;
  if (n) k_fr_vandermonde<typename Curve::Fr><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(pw, n, (typename Curve::Fr*)d_out);
  ctx->launches++;
  return launch_check(ctx, "k_fr_vandermonde");
}

template <class Curve>
static int elementwise_t(bpgpu_ctx* ctx, int op, const void* a, const void* b, const uint8_t* s_be, size_t n, void* out) {
  using Fr = typename Curve::Fr;
  Fr* sc = nullptr;
  if (s_be) { int rc = fr_args_upload<Curve>(ctx, s_be, 1, &sc); if (rc) return rc; }
  if (n) k_fr_elementwise<Fr><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(op, (const Fr*)a, (const Fr*)b, sc, n, (Fr*)out);
  ctx->launches++;
  return launch_check(ctx, "k_fr_elementwise");
}

}  // namespace bp

using namespace bp;

static bool range_ok(const bpgpu_scalars* s, size_t off, size_t n) { return s && off <= s->n && n <= s->n - off; }
static inline void* at(const bpgpu_scalars* s, size_t off) { return (uint8_t*)s->d + off * 32; }

extern "C" {

int bpgpu_scalars_alloc(bpgpu_ctx* ctx, size_t n, bpgpu_scalars** out) {
  if (!ctx || !out) return BPGPU_E_ARG;
  *out = nullptr;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  bpgpu_scalars* s = new (std::nothrow) bpgpu_scalars();
  if (!s) return BPGPU_E_CUDA;
  s->ctx = ctx; s->n = n; s->d = nullptr;
  if (dev_alloc(ctx, &s->d, n * 32) != cudaSuccess) { delete s; return BPGPU_E_CUDA; }
  if (cudaMemsetAsync(s->d, 0, n ? n * 32 : 16, ctx->stream) != cudaSuccess) { dev_free(ctx, s->d); delete s; return BPGPU_E_CUDA; }
  *out = s;
  return BPGPU_OK;
}

// k zeroed vectors of n elements carved out of ONE allocation (one pool request, one memset, one release with the last
// handle) -- the four outputs of the R1CS polynomial kernels
int scalars_alloc_many(bpgpu_ctx* ctx, size_t n, int k, bpgpu_scalars** const* outs) {
  for (int i = 0; i < k; i++) *outs[i] = nullptr;
  void* base = nullptr;
  const size_t bytes = (size_t)k * n * 32;
  if (dev_alloc(ctx, &base, bytes) != cudaSuccess) return BPGPU_E_CUDA;
  if (cudaMemsetAsync(base, 0, bytes ? bytes : 16, ctx->stream) != cudaSuccess) { dev_free(ctx, base); return BPGPU_E_CUDA; }
  bpgpu_shared_block* blk = new (std::nothrow) bpgpu_shared_block{base, 0};
  if (!blk) { dev_free(ctx, base); return BPGPU_E_CUDA; }
  for (int i = 0; i < k; i++) {
    bpgpu_scalars* s = new (std::nothrow) bpgpu_scalars();
    if (!s) {
      for (int j = 0; j < i; j++) { delete *outs[j]; *outs[j] = nullptr; }
      dev_free(ctx, base);
      delete blk;
      return BPGPU_E_CUDA;
    }
    s->ctx = ctx; s->d = (uint8_t*)base + (size_t)i * n * 32; s->n = n; s->blk = blk;
    blk->refs++;
    *outs[i] = s;
  }
  return BPGPU_OK;
}

int bpgpu_fr_vandermonde(bpgpu_ctx* ctx, const uint8_t* x_be, size_t n, bpgpu_scalars** out) {
  if (!ctx || !x_be || !out) return BPGPU_E_ARG;
  if (n >= (1ull << 32)) return BPGPU_E_ARG;
  int rc = bpgpu_scalars_alloc(ctx, n, out);
  if (rc) return rc;
  rc = ctx->curve == BPGPU_BLS12_381 ? vandermonde_t<Bls>(ctx, x_be, n, (*out)->d) : vandermonde_t<Bn>(ctx, x_be, n, (*out)->d);
  if (rc) { bpgpu_scalars_free(*out); *out = nullptr; }
  return rc;
}

static int ew(bpgpu_ctx* ctx, int op, const bpgpu_scalars* a, size_t aoff, const bpgpu_scalars* b, size_t boff, const uint8_t* s_be,
              size_t n, bpgpu_scalars* out, size_t ooff) {
  if (!ctx || !a || !out) return BPGPU_E_ARG;
  if (!range_ok(a, aoff, n) || !range_ok(out, ooff, n) || (b && !range_ok(b, boff, n))) return BPGPU_E_LEN;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  const void* bp_ = b ? at(b, boff) : nullptr;
  return ctx->curve == BPGPU_BLS12_381 ? elementwise_t<Bls>(ctx, op, at(a, aoff), bp_, s_be, n, at(out, ooff))
                                      : elementwise_t<Bn>(ctx, op, at(a, aoff), bp_, s_be, n, at(out, ooff));
}

int bpgpu_fr_hadamard(bpgpu_ctx* ctx, const bpgpu_scalars* a, size_t aoff, const bpgpu_scalars* b, size_t boff, size_t n,
                      bpgpu_scalars* out, size_t ooff) {
  if (!b) return BPGPU_E_ARG;
  return ew(ctx, 0, a, aoff, b, boff, nullptr, n, out, ooff);
}
int bpgpu_fr_add(bpgpu_ctx* ctx, const bpgpu_scalars* a, size_t aoff, const bpgpu_scalars* b, size_t boff, size_t n,
                 bpgpu_scalars* out, size_t ooff) {
  if (!b) return BPGPU_E_ARG;
  return ew(ctx, 1, a, aoff, b, boff, nullptr, n, out, ooff);
}
int bpgpu_fr_sub(bpgpu_ctx* ctx, const bpgpu_scalars* a, size_t aoff, const bpgpu_scalars* b, size_t boff, size_t n,
                 bpgpu_scalars* out, size_t ooff) {
  if (!b) return BPGPU_E_ARG;
  return ew(ctx, 2, a, aoff, b, boff, nullptr, n, out, ooff);
}
int bpgpu_fr_scale(bpgpu_ctx* ctx, const bpgpu_scalars* a, size_t aoff, size_t n, const uint8_t* s_be, bpgpu_scalars* out, size_t ooff) {
  if (!s_be) return BPGPU_E_ARG;
  return ew(ctx, 3, a, aoff, nullptr, 0, s_be, n, out, ooff);
}

int bpgpu_fr_inner_product(bpgpu_ctx* ctx, const bpgpu_scalars* a, size_t aoff, const bpgpu_scalars* b, size_t boff, size_t n,
                           uint8_t* out_be) {
  if (!ctx || !a || !b || !out_be) return BPGPU_E_ARG;
  if (!range_ok(a, aoff, n) || !range_ok(b, boff, n)) return BPGPU_E_LEN;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  DotPtrs p;
  for (int k = 0; k < 9; k++) { p.a[k] = at(a, aoff); p.b[k] = at(b, boff); }
  return ctx->curve == BPGPU_BLS12_381 ? fr_dots_to_host<Bls>(ctx, p, 1, n, out_be) : fr_dots_to_host<Bn>(ctx, p, 1, n, out_be);
}

// VecPoly3::special_inner_product (vector_poly.rs:79-97): t1..t6 from l1,l2,l3,r0,r1,r3 in one launch pair
int bpgpu_fr_poly3_special_inner_product(bpgpu_ctx* ctx, const bpgpu_scalars* l1, const bpgpu_scalars* l2, const bpgpu_scalars* l3,
                                         const bpgpu_scalars* r0, const bpgpu_scalars* r1, const bpgpu_scalars* r3, size_t n,
                                         uint8_t* t_be /* 6 * MODBYTES: t1..t6 */) {
  if (!ctx || !l1 || !l2 || !l3 || !r0 || !r1 || !r3 || !t_be) return BPGPU_E_ARG;
  const bpgpu_scalars* all[6] = {l1, l2, l3, r0, r1, r3};
  for (auto v : all) if (!range_ok(v, 0, n)) return BPGPU_E_LEN;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  // <l1,r0>, <l1,r1>, <l2,r0>, <l2,r1>, <l3,r0>, <l1,r3>, <l3,r1>, <l2,r3>, <l3,r3>
  const bpgpu_scalars* A[9] = {l1, l1, l2, l2, l3, l1, l3, l2, l3};
  const bpgpu_scalars* B[9] = {r0, r1, r0, r1, r0, r3, r1, r3, r3};
  DotPtrs p;
  for (int k = 0; k < 9; k++) { p.a[k] = A[k]->d; p.b[k] = B[k]->d; }
  int mb = bpgpu_modbytes(ctx->curve);
  uint8_t d[9 * 48];
  int rc = ctx->curve == BPGPU_BLS12_381 ? fr_dots_to_host<Bls>(ctx, p, 9, n, d) : fr_dots_to_host<Bn>(ctx, p, 9, n, d);
  if (rc) return rc;
  // t1 = d0; t2 = d1+d2; t3 = d3+d4; t4 = d5+d6; t5 = d7; t6 = d8  (host adds on canonical values)
  auto add = [&](int i, int j, uint8_t* out) {
    if (ctx->curve == BPGPU_BLS12_381) { using H = host::HFp<BlsFr>; (H::from_be(d + i * mb, mb) + (j >= 0 ? H::from_be(d + j * mb, mb) : H::zero())).to_be(out, mb); }
    else { using H = host::HFp<BnFr>; (H::from_be(d + i * mb, mb) + (j >= 0 ? H::from_be(d + j * mb, mb) : H::zero())).to_be(out, mb); }
  };
  add(0, -1, t_be); add(1, 2, t_be + mb); add(3, 4, t_be + 2 * mb); add(5, 6, t_be + 3 * mb); add(7, -1, t_be + 4 * mb); add(8, -1, t_be + 5 * mb);
  return BPGPU_OK;
}

// FieldElement::batch_invert (ipp.rs:295): elementwise inverses (0 -> 0) and the product of all inverses
int bpgpu_fr_batch_invert(bpgpu_ctx* ctx, const bpgpu_scalars* a, size_t aoff, size_t n, bpgpu_scalars* out, size_t ooff,
                          uint8_t* prod_inv_be) {
  if (!ctx || !a || !out) return BPGPU_E_ARG;
  if (!range_ok(a, aoff, n) || !range_ok(out, ooff, n)) return BPGPU_E_LEN;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  if (n) {
    if (ctx->curve == BPGPU_BLS12_381) k_fr_invert<Bls::Fr><<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>((const Bls::Fr*)at(a, aoff), n, (Bls::Fr*)at(out, ooff));
    else k_fr_invert<Bn::Fr><<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>((const Bn::Fr*)at(a, aoff), n, (Bn::Fr*)at(out, ooff));
    ctx->launches++;
    int rc = launch_check(ctx, "k_fr_invert");
    if (rc) return rc;
  }
  if (prod_inv_be) {
    // product of the inverses: download and multiply on the host (n is the number of IPP rounds: <= 32)
    int mb = bpgpu_modbytes(ctx->curve);
    std::vector<uint8_t> tmp(n * mb + 1);
    int rc = bpgpu_scalars_download(ctx, out, ooff, n, tmp.data());
    if (rc) return rc;
    if (ctx->curve == BPGPU_BLS12_381) { using H = host::HFp<BlsFr>; H p = H::one(); for (size_t i = 0; i < n; i++) p = p * H::from_be(tmp.data() + i * mb, mb); p.to_be(prod_inv_be, mb); }
    else { using H = host::HFp<BnFr>; H p = H::one(); for (size_t i = 0; i < n; i++) p = p * H::from_be(tmp.data() + i * mb, mb); p.to_be(prod_inv_be, mb); }
  }
  return BPGPU_OK;
}

}  // extern "C"
