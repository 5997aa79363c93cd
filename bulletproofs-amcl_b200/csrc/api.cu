// C ABI of libbpgpu (include/bpgpu.h): context, residency, byte<->Montgomery conversion kernels,
// the MSM entry points, and the self-test / microbenchmark hooks.
#include <string.h>

#include <mutex>
#include <new>

#include "common.cuh"
#include "host_fp.h"

namespace bp {

int msm_window_bits(size_t n);

int launch_check(bpgpu_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "bpgpu: launch of %s failed: %s\n", what, cudaGetErrorString(e));
    return BPGPU_E_CUDA;
  }
  (void)ctx;
  return BPGPU_OK;
}

// ------------------------------------------------------------------ conversion kernels
// X || Y -> affine Montgomery form.  Validity is decided as AMCL's ECP::frombytes / new_bigs do (coordinates < p, on the curve,
// (0, 1) = identity); an invalid point becomes the identity on the device and raises the ctx's bad-input word, which fails
// the entry point that consumed it with BPGPU_E_FORMAT.
template <class Curve>
__global__ void k_points_from_be(const uint8_t* __restrict__ xy, size_t n, Affine<typename Curve::Fq>* __restrict__ out, uint32_t* bad) {
  using Fq = typename Curve::Fq;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<Fq> a;
  if (!g1_from_be_checked<Curve>(xy + i * 2 * Curve::MODBYTES, &a)) {
    a = Affine<Fq>::inf();
    *(volatile uint32_t*)bad = 1;
    __threadfence_system();
  }
  store_vec(out + i, a);
}

template <class Fq>
__device__ __forceinline__ void affine_to_be2(const Affine<Fq>& a, int modbytes, uint8_t* out) {
  if (a.is_inf()) {
    for (int i = 0; i < 2 * modbytes; i++) out[i] = 0;
    out[2 * modbytes - 1] = 1;
    return;
  }
  Fq x = a.x.from_mont(), y = a.y.from_mont();
  limbs_to_be<Fq::N>(x.v, modbytes, out);
  limbs_to_be<Fq::N>(y.v, modbytes, out + modbytes);
}

template <class Curve>
__global__ void k_points_to_be(const Affine<typename Curve::Fq>* __restrict__ in, size_t n, uint8_t* __restrict__ xy) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<typename Curve::Fq> a = load_vec(in + i);
  affine_to_be2(a, Curve::MODBYTES, xy + i * 2 * Curve::MODBYTES);
}

// big-endian MODBYTES -> Fr, reduced mod r as FieldElement::from(&[u8; MODBYTES]) does (all 48 bytes count on BLS12-381);
// mont = 0 leaves the canonical integer (MSM digit source)
template <class Curve>
__global__ void k_scalars_from_be(const uint8_t* __restrict__ be, size_t n, int mont, typename Curve::Fr* __restrict__ out) {
  using Fr = typename Curve::Fr;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* src = be + i * Curve::MODBYTES;
  bool wide = false;
  for (int k = 0; k < Curve::MODBYTES - 32; k++) wide = wide || src[k] != 0;
  Fr s;
  if (wide) {
    s = fr_from_be_wide<Curve>(src);
    if (!mont) s = s.from_mont();
  } else {
    be_to_limbs<8>(src, Curve::MODBYTES, s.v);
    canonicalise(s);
    if (mont) s = s.to_mont();
  }
  store_vec(out + i, s);
}

template <class Curve>
__global__ void k_scalars_to_be(const typename Curve::Fr* __restrict__ in, size_t n, uint8_t* __restrict__ be) {
  using Fr = typename Curve::Fr;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = load_vec(in + i).from_mont();
  limbs_to_be<8>(s.v, Curve::MODBYTES, be + i * Curve::MODBYTES);
}

// ------------------------------------------------------------------ self-test kernels
template <class F>
__global__ void k_field_op(int op, const uint8_t* a, const uint8_t* b, size_t n, int modbytes, uint8_t* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x, y, r;
  be_to_limbs<F::N>(a + i * modbytes, modbytes, x.v);
  be_to_limbs<F::N>(b + i * modbytes, modbytes, y.v);
  canonicalise(x); canonicalise(y);
  x = x.to_mont(); y = y.to_mont();
  switch (op) {
    case 0: r = x * y; break;
    case 1: r = x + y; break;
    case 2: r = x - y; break;
    case 3: r = x.inv(); break;
    default: r = x.sqr(); break;
  }
  r = r.from_mont();
  limbs_to_be<F::N>(r.v, modbytes, out + i * modbytes);
}

template <class Curve>
__global__ void k_group_op(int op, const Affine<typename Curve::Fq>* p, const Affine<typename Curve::Fq>* q,
                           const typename Curve::Fr* k, size_t n, Affine<typename Curve::Fq>* out) {
  using Fq = typename Curve::Fq;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<Fq> P = XYZZ<Fq>::from_affine(load_vec(p + i));
  if (op == 0) {
    P.madd(load_vec(q + i));
  } else if (op == 1) {
    P.dbl();
  } else if (op == 2) {
    typename Curve::Fr s = load_vec(k + i);
    P = mul_limbs(P, s.v, 8);
  } else {
    XYZZ<Fq> Q = XYZZ<Fq>::from_affine(load_vec(q + i));
    Q.dbl(); Q.add(P);                       // exercises the full add: 2Q + P
    P = Q;
  }
  store_vec(out + i, P.to_affine());
}

// ------------------------------------------------------------------ integer-pipe microbenchmarks
// kind 0: independent IMAD.WIDE.U32 accumulate chains, 8 per thread, no memory traffic
__global__ void __launch_bounds__(256) k_imad_wide(int iters, uint32_t seed, uint64_t* sink) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
  uint64_t acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = seed + k;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int k = 0; k < 8; k++) acc[k] = (uint64_t)a * (b + k) + acc[k];   // IMAD.WIDE.U32
      a += (uint32_t)acc[0];
    }
  }
  uint64_t x = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) x ^= acc[k];
  if (x == 0x123456789abcdefull) sink[0] = x;
}
// kind 2: plain 32-bit IMAD (mad.lo) chains; kind 3: 1 IMAD.WIDE : 2 IMAD mix
__global__ void __launch_bounds__(256) k_imad_lo(int iters, uint32_t seed, int mix, uint64_t* sink) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
  uint32_t acc[16];
  uint64_t wacc[4];
#pragma unroll
  for (int k = 0; k < 16; k++) acc[k] = seed + k;
#pragma unroll
  for (int k = 0; k < 4; k++) wacc[k] = seed + k;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int k = 0; k < 16; k++) acc[k] = a * (b + k) + acc[k];            // IMAD
      if (mix) {
#pragma unroll
        for (int k = 0; k < 4; k++) { wacc[k] = (uint64_t)a * (b + k) + wacc[k]; wacc[k] = (uint64_t)b * (a + k) + wacc[k]; }   // 8 IMAD.WIDE
      }
      a += acc[0];
    }
  }
  uint64_t x = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) x ^= acc[k];
#pragma unroll
  for (int k = 0; k < 4; k++) x ^= wacc[k];
  if (x == 0x123456789abcdefull) sink[0] = x;
}
// kind 1 / 10+SP: dependent chains of Fq Montgomery products per thread (2 interleaved chains)
template <class F, int SP>
__global__ void __launch_bounds__(256) k_fq_mul_chain(int iters, const F* in, F* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  F x = load_vec(in + (i & 255)), y = load_vec(in + ((i + 7) & 255));
  for (int it = 0; it < iters; it++) { x = F::template mul_t<SP>(x, y); y = F::template mul_t<SP>(y, x); }
  if (x.is_zero() && y.is_zero()) store_vec(out, x);
}

// kind 33: a chain of dependent XYZZ doublings on one warp (latency of the serial part of a scalar multiplication)
template <class F>
__global__ void k_dbl_chain(int iters, const XYZZ<F>* in, XYZZ<F>* out) {
  XYZZ<F> p = load_vec(in + threadIdx.x);
  p.zz = p.zz + F::one();
  for (int it = 0; it < iters; it++) p.dbl();
  if (p.x.is_zero() && p.y.is_zero() && p.zzz.is_zero()) store_vec(out, p);
}

template <class Curve>
static int points_from_host_on(bpgpu_ctx* ctx, cudaStream_t st, const uint8_t* xy, size_t n, void* dst) {
  if (n == 0) return BPGPU_OK;
  size_t bytes = n * 2 * Curve::MODBYTES;
  int rc = ctx->io_dev.reserve(bytes);
  if (rc) return rc;
  BP_CUDA_OK(cudaMemcpyAsync(ctx->io_dev.p, xy, bytes, cudaMemcpyHostToDevice, st));
  k_points_from_be<Curve><<<(unsigned)((n + 127) / 128), 128, 0, st>>>((const uint8_t*)ctx->io_dev.p, n, (Affine<typename Curve::Fq>*)dst, ctx->bad_input);
  ctx->launches++;
  return launch_check(ctx, "k_points_from_be");
}
template <class Curve>
int points_from_host(bpgpu_ctx* ctx, const uint8_t* xy, size_t n, void* dst) {
  return points_from_host_on<Curve>(ctx, ctx->stream, xy, n, dst);
}
template int points_from_host<Bls>(bpgpu_ctx*, const uint8_t*, size_t, void*);
template int points_from_host<Bn>(bpgpu_ctx*, const uint8_t*, size_t, void*);

template <class Curve>
static int scalars_upload_t(bpgpu_ctx* ctx, const uint8_t* be, size_t n, int mont, void* dst, Scratch& stage) {
  if (n == 0) return BPGPU_OK;
  size_t bytes = n * Curve::MODBYTES;
  int rc = stage.reserve(bytes);
  if (rc) return rc;
  BP_CUDA_OK(cudaMemcpyAsync(stage.p, be, bytes, cudaMemcpyHostToDevice, ctx->stream));
  k_scalars_from_be<Curve><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint8_t*)stage.p, n, mont,
                                                                               (typename Curve::Fr*)dst);
  ctx->launches++;
  return launch_check(ctx, "k_scalars_from_be");
}

template <class Curve>
int scalars_from_host(bpgpu_ctx* ctx, const uint8_t* be, size_t n, int mont, void* dst) {
  return scalars_upload_t<Curve>(ctx, be, n, mont, dst, ctx->io_dev2);
}
template int scalars_from_host<Bls>(bpgpu_ctx*, const uint8_t*, size_t, int, void*);
template int scalars_from_host<Bn>(bpgpu_ctx*, const uint8_t*, size_t, int, void*);

static int fetch_result(bpgpu_ctx* ctx, const void* d_src, size_t bytes, uint8_t* host_out) {
  BP_CUDA_OK(cudaMemcpyAsync(ctx->pinned, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA_OK(stream_sync(ctx));
  memcpy(host_out, ctx->pinned, bytes);
  return BPGPU_OK;
}

}  // namespace bp

namespace bp {
// Host finish of one MSM: Horner over the general path's window sums (if any) + the table path's sum (if any),
// then one affine normalisation.
// set / nsets: which scalar set of a multi-set run (MsmResult::nsets) the W windows belong to
template <class FqParams>
static void msm_finish_mixed(const uint8_t* winsum_bytes, int W, int c, int qshift, const uint8_t* table_sum, int modbytes, uint8_t* out_xy,
                             int set = 0, int nsets = 1) {
  using HP = host::HXYZZ<FqParams>;
  HP acc = HP::inf();
  const HP* wp = reinterpret_cast<const HP*>(winsum_bytes) + (size_t)set * W;            // P_w of this set
  const HP* wq = reinterpret_cast<const HP*>(winsum_bytes) + (size_t)(nsets + set) * W;  // Q_w of this set
  // acc * 2^c + P_w + 2^q * Q_w = ((acc * 2^(c-q)) + Q_w) * 2^q + P_w : the shift of Q_w rides the window's own c doublings
  // (q = 5 + lgL1 < c always: a level-1 segment is at most half a window)
  const int q = qshift < c ? qshift : 0;
  for (int w = W - 1; w >= 0; w--) {
    HP qw = wq[w];
    if (q == 0 && !qw.is_inf()) for (int k = 0; k < qshift; k++) qw.dbl();
    if (w != W - 1) for (int k = 0; k < c - q; k++) acc.dbl();
    acc.add(qw);
    for (int k = 0; k < q; k++) acc.dbl();
    acc.add(wp[w]);
  }
  if (table_sum) acc.add(*reinterpret_cast<const HP*>(table_sum));
  acc.to_xy_be(modbytes, out_xy);
}

// The MSM in two halves: `begin` enqueues the device pipeline and the copy of its per-window sums on the ctx stream and
// returns; `finish` waits for the stream, checks the inputs and does the host part (Horner + one inversion).  Between the
// two the host thread is free -- to finish the previous MSM of ANOTHER context, whose sort stages (HBM / latency bound)
// then run under this one's bucket accumulation (integer pipe bound): bpgpu_msm_*_begin / bpgpu_msm_finish.
int msm_mixed_begin(bpgpu_ctx* ctx, const TableSeg* segs, int nsegs, const void* d_pts, const void* d_scal, bool mont, size_t n) {
  const bool bls = ctx->curve == BPGPU_BLS12_381;
  const size_t psz = bls ? sizeof(XYZZ<Bls::Fq>) : sizeof(XYZZ<Bn::Fq>);
  if (ctx->pending.active) return BPGPU_E_ARG;                    // one MSM in flight per ctx
  MsmResult res;
  res.W = 0; res.c = 0; res.qshift = 0; res.d_winsum = nullptr;
  int rc;
  size_t tn = 0;
  for (int k = 0; k < nsegs; k++) tn += segs[k].n;
  int hp = 0;
  if (tn) {
    rc = bls ? table_sum_run<Bls>(ctx, segs, nsegs, 1, &hp) : table_sum_run<Bn>(ctx, segs, nsegs, 1, &hp);
    if (rc) return rc;
  }
  if (n) {
    rc = bls ? msm_run<Bls>(ctx, (const Affine<Bls::Fq>*)d_pts, d_scal, mont, n, &res) : msm_run<Bn>(ctx, (const Affine<Bn::Fq>*)d_pts, d_scal, mont, n, &res);
    if (rc) return rc;
  }
  uint8_t* tsum = ctx->pinned + 2 * 128 * psz;      // after the window sums (W <= 86 at c = 3)
  if (res.W) BP_CUDA_OK(cudaMemcpyAsync(ctx->pinned, res.d_winsum, 2 * res.W * psz, cudaMemcpyDeviceToHost, ctx->stream));
  if (tn) {
    if (hp) BP_CUDA_OK(cudaMemcpyAsync(tsum, (const uint8_t*)ctx->tbl_part.p + TBL_MAX_GROUPS * psz, (size_t)hp * psz, cudaMemcpyDeviceToHost,
                                       ctx->stream));
    else BP_CUDA_OK(cudaMemcpyAsync(tsum, ctx->tbl_part.p, psz, cudaMemcpyDeviceToHost, ctx->stream));
  }
  ctx->pending.active = true;
  ctx->pending.W = res.W; ctx->pending.c = res.c; ctx->pending.qshift = res.qshift;
  ctx->pending.tn = tn; ctx->pending.hp = hp;
  return BPGPU_OK;
}

int msm_mixed_finish(bpgpu_ctx* ctx, uint8_t* out_xy) {
  if (!ctx->pending.active) return BPGPU_E_ARG;
  const int mb = bpgpu_modbytes(ctx->curve);
  const bool bls = ctx->curve == BPGPU_BLS12_381;
  const size_t psz = bls ? sizeof(XYZZ<Bls::Fq>) : sizeof(XYZZ<Bn::Fq>);
  const int W = ctx->pending.W, c = ctx->pending.c, qshift = ctx->pending.qshift, hp = ctx->pending.hp;
  const size_t tn = ctx->pending.tn;
  ctx->pending.active = false;
  uint8_t* tsum = ctx->pinned + 2 * 128 * psz;
  if (W || tn) BP_CUDA_OK(stream_sync(ctx));
  if (ctx->wait_points) {                  // the pipeline did not get as far as the wait (error path): drain the copy queue
    ctx->wait_points = false;
    cudaStreamSynchronize(ctx->copy_stream);
  }
  int rc;
  if ((rc = inputs_ok(ctx))) return rc;
  if (hp) { if (bls) host_sum_partials<BlsFq>(tsum, 1, hp); else host_sum_partials<BnFq>(tsum, 1, hp); }
  if (bls) msm_finish_mixed<BlsFq>(ctx->pinned, W, c, qshift, tn ? tsum : nullptr, mb, out_xy);
  else msm_finish_mixed<BnFq>(ctx->pinned, W, c, qshift, tn ? tsum : nullptr, mb, out_xy);
  return BPGPU_OK;
}

int msm_mixed_to_host(bpgpu_ctx* ctx, const TableSeg* segs, int nsegs, const void* d_pts, const void* d_scal, bool mont, size_t n,
                      uint8_t* out_xy) {
  int rc = msm_mixed_begin(ctx, segs, nsegs, d_pts, d_scal, mont, n);
  if (rc) return rc;
  return msm_mixed_finish(ctx, out_xy);
}

int msm_pair_to_host(bpgpu_ctx* ctx, const void* d_pts, const void* d_scal_a, const void* d_scal_b, bool mont, size_t n, uint8_t* out_a_xy,
                     uint8_t* out_b_xy) {
  const int mb = bpgpu_modbytes(ctx->curve);
  const bool bls = ctx->curve == BPGPU_BLS12_381;
  const size_t psz = bls ? sizeof(XYZZ<Bls::Fq>) : sizeof(XYZZ<Bn::Fq>);
  MsmResult res;
  int rc = bls ? msm_run<Bls>(ctx, (const Affine<Bls::Fq>*)d_pts, d_scal_a, mont, n, &res, d_scal_b)
               : msm_run<Bn>(ctx, (const Affine<Bn::Fq>*)d_pts, d_scal_a, mont, n, &res, d_scal_b);
  if (rc) return rc;
  if (res.W) {
    if ((size_t)4 * res.W * psz > ctx->pinned_cap / 2) return BPGPU_E_ARG;
    BP_CUDA_OK(cudaMemcpyAsync(ctx->pinned, res.d_winsum, (size_t)4 * res.W * psz, cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA_OK(stream_sync(ctx));
  }
  if ((rc = inputs_ok(ctx))) return rc;
  uint8_t* outs[2] = {out_a_xy, out_b_xy};
  for (int s = 0; s < 2; s++) {
    if (bls) msm_finish_mixed<BlsFq>(ctx->pinned, res.W, res.c, res.qshift, nullptr, mb, outs[s], s, 2);
    else msm_finish_mixed<BnFq>(ctx->pinned, res.W, res.c, res.qshift, nullptr, mb, outs[s], s, 2);
  }
  return BPGPU_OK;
}

int msm_to_host(bpgpu_ctx* ctx, const void* d_pts, const void* d_scal, bool mont, size_t n, uint8_t* out_xy) {
  return msm_mixed_to_host(ctx, nullptr, 0, d_pts, d_scal, mont, n, out_xy);
}
}  // namespace bp

using namespace bp;

#define DISPATCH(ctx, CALL)                          \
  ((ctx)->curve == BPGPU_BLS12_381 ? CALL(Bls) : CALL(Bn))

// one stream-ordered pool per device, created on first use and never destroyed (contexts come and go, e.g. the second
// driver of a batch call: a fresh pool per context would pay the physical allocation of its scratch again every call)
static cudaMemPool_t library_pool(int device) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {nullptr};
  if (device < 0 || device >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!pools[device]) {
    cudaMemPoolProps pp;
    memset(&pp, 0, sizeof pp);
    pp.allocType = cudaMemAllocationTypePinned;
    pp.handleTypes = cudaMemHandleTypeNone;
    pp.location.type = cudaMemLocationTypeDevice;
    pp.location.id = device;
    cudaMemPool_t p = nullptr;
    if (cudaMemPoolCreate(&p, &pp) != cudaSuccess) return nullptr;
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(p, cudaMemPoolAttrReleaseThreshold, &keep);
    pools[device] = p;
  }
  return pools[device];
}

extern "C" {

const char* bpgpu_strerror(int code) {
  switch (code) {
    case BPGPU_OK: return "ok";
    case BPGPU_E_LEN: return "unequal vector lengths (ValueError::UnequalSizeVectors)";
    case BPGPU_E_NOT_POW2: return "vector length is not a power of two";
    case BPGPU_E_GENS_LEN: return "not enough generators (R1CSError::InvalidGeneratorsLength)";
    case BPGPU_E_VERIFY: return "verification failed (R1CSError::VerificationError)";
    case BPGPU_E_FORMAT: return "malformed input (R1CSError::FormatError)";
    case BPGPU_E_CUDA: return "CUDA failure or no usable device";
    case BPGPU_E_ARG: return "bad argument";
    default: return "unknown error";
  }
}

int bpgpu_modbytes(int curve) { return curve == BPGPU_BLS12_381 ? 48 : (curve == BPGPU_BN254 ? 32 : BPGPU_E_ARG); }

int bpgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int bpgpu_ctx_create(int curve, int device, bpgpu_ctx** out) {
  if (!out || (curve != BPGPU_BLS12_381 && curve != BPGPU_BN254)) return BPGPU_E_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    fprintf(stderr, "bpgpu: no CUDA device (this library has no CPU fallback)\n");
    return BPGPU_E_CUDA;
  }
  if (device < 0 || device >= n) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(device));
  bpgpu_ctx* c = new (std::nothrow) bpgpu_ctx();
  if (!c) return BPGPU_E_CUDA;
  c->curve = curve;
  c->device = device;
  // any failure below releases what was created so far (bpgpu_ctx_destroy copes with a partially built ctx)
#define CTX_OK(expr)                                                                           \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      fprintf(stderr, "bpgpu: CUDA error %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__); \
      bpgpu_ctx_destroy(c);                                                                    \
      return BPGPU_E_CUDA;                                                                     \
    }                                                                                          \
  } while (0)
  cudaDeviceProp prop;
  CTX_OK(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  CTX_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CTX_OK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CTX_OK(cudaEventCreateWithFlags(&c->points_ready, cudaEventDisableTiming));
  CTX_OK(cudaEventCreateWithFlags(&c->sync_event, cudaEventDisableTiming | cudaEventBlockingSync));
  { const char* e = getenv("BPGPU_BLOCKING_SYNC"); c->blocking_sync = e && atoi(e) != 0; }
  // handle storage comes from a memory pool the LIBRARY owns (one per device, shared by its contexts, kept for the life of
  // the process): freed blocks stay in it for the next proof and the next context (release threshold raised), and the
  // device's default pool -- which other cudaMallocAsync users of the process share -- is left alone
  c->pool = library_pool(device);
  if (!c->pool) { bpgpu_ctx_destroy(c); return BPGPU_E_CUDA; }
  c->pinned_cap = 1 << 18;
  CTX_OK(cudaHostAlloc((void**)&c->pinned, c->pinned_cap, cudaHostAllocDefault));
  CTX_OK(cudaHostAlloc((void**)&c->bad_input, 64, cudaHostAllocMapped));
  *c->bad_input = 0;
#undef CTX_OK
  *out = c;
  return BPGPU_OK;
}

int bpgpu_ctx_aux(bpgpu_ctx* c, bpgpu_ctx** out) {
  if (!c || !out) return BPGPU_E_ARG;
  if (!c->aux) {
    int rc = bpgpu_ctx_create(c->curve, c->device, &c->aux);
    if (rc) return rc;
  }
  *out = c->aux;
  return BPGPU_OK;
}

void bpgpu_ctx_destroy(bpgpu_ctx* c) {
  if (!c) return;
  if (c->aux) { bpgpu_ctx_destroy(c->aux); c->aux = nullptr; }
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (bpgpu_fixed_bases* fb : c->fb_cache) bpgpu_fixed_bases_free(fb);
  c->fb_cache.clear();
  for (auto& kv : c->circuit_cache) bpgpu_circuit_free(kv.second);
  c->circuit_cache.clear();
  c->msm_a.release(); c->msm_b.release(); c->msm_c.release(); c->msm_d.release(); c->msm_e.release();
  c->io_dev.release(); c->io_dev2.release();
  c->ipp_pts.release(); c->ipp_scl.release(); c->parts_pts.release(); c->parts_scl.release(); c->tbl_part.release();
  c->fr_tmp.release(); c->fr_out.release(); c->fr_args.release(); c->fr_pow.release(); c->fr_pow2.release();
  c->vb.release();
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->bad_input) cudaFreeHost(c->bad_input);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->points_ready) cudaEventDestroy(c->points_ready);
  if (c->sync_event) cudaEventDestroy(c->sync_event);
  delete c;
}

void* bpgpu_ctx_stream(bpgpu_ctx* c) { return c ? (void*)c->stream : nullptr; }
int bpgpu_ctx_sync(bpgpu_ctx* c) {
  if (!c) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaStreamSynchronize(c->stream));
  return BPGPU_OK;
}
int bpgpu_ctx_curve(const bpgpu_ctx* c) { return c ? c->curve : BPGPU_E_ARG; }
int bpgpu_ctx_device(const bpgpu_ctx* c) { return c ? c->device : -1; }
uint64_t bpgpu_ctx_launches(const bpgpu_ctx* c) { return c ? c->launches : 0; }

int bpgpu_ctx_set_fixed_schedule(bpgpu_ctx* c, int on) {
  if (!c) return BPGPU_E_ARG;
  c->fixed_schedule = on != 0;
  return BPGPU_OK;
}

int bpgpu_ctx_set_blocking_sync(bpgpu_ctx* c, int on) {
  if (!c) return BPGPU_E_ARG;
  c->blocking_sync = on != 0;
  return BPGPU_OK;
}

int bpgpu_ctx_set_profile(bpgpu_ctx* c, int on) {
  if (!c) return BPGPU_E_ARG;
  c->profile = on;
  for (int i = 0; i < 8; i++) c->stage_ms_sum[i] = 0;
  c->stage_runs = 0;
  return BPGPU_OK;
}
int bpgpu_msm_stage_ms(const bpgpu_ctx* c, double* avg_ms, int cap) {
  if (!c || !avg_ms) return BPGPU_E_ARG;
  for (int i = 0; i < cap && i < 8; i++) avg_ms[i] = c->stage_runs ? c->stage_ms_sum[i] / (double)c->stage_runs : 0.0;
  return (int)c->stage_runs;
}

void* bpgpu_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void bpgpu_host_free(void* p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------ residency
int bpgpu_points_upload(bpgpu_ctx* ctx, const uint8_t* xy, size_t n, bpgpu_points** out) {
  if (!ctx || !out || (!xy && n)) return BPGPU_E_ARG;
  *out = nullptr;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  bpgpu_points* p = new (std::nothrow) bpgpu_points();
  if (!p) return BPGPU_E_CUDA;
  p->ctx = ctx; p->n = n; p->d = nullptr;
  size_t psz = (ctx->curve == BPGPU_BLS12_381) ? sizeof(Affine<Bls::Fq>) : sizeof(Affine<Bn::Fq>);
  if (dev_alloc(ctx, &p->d, n * psz) != cudaSuccess) { delete p; return BPGPU_E_CUDA; }
#define CALL(C) points_from_host<C>(ctx, xy, n, p->d)
  int rc = DISPATCH(ctx, CALL);
#undef CALL
  if (rc == BPGPU_OK && stream_sync(ctx) != cudaSuccess) rc = BPGPU_E_CUDA;
  if (rc == BPGPU_OK) rc = inputs_ok(ctx);
  if (rc) { dev_free(ctx, p->d); delete p; return rc; }
  *out = p;
  return BPGPU_OK;
}

int bpgpu_points_download(bpgpu_ctx* ctx, const bpgpu_points* p, size_t off, size_t n, uint8_t* xy) {
  if (!ctx || !p || (!xy && n)) return BPGPU_E_ARG;
  if (off > p->n || n > p->n - off) return BPGPU_E_ARG;
  if (n == 0) return BPGPU_OK;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  size_t bytes = n * 2 * bpgpu_modbytes(ctx->curve);
  int rc = ctx->io_dev.reserve(bytes);
  if (rc) return rc;
  if (ctx->curve == BPGPU_BLS12_381)
    k_points_to_be<Bls><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const Affine<Bls::Fq>*)p->d + off, n, (uint8_t*)ctx->io_dev.p);
  else
    k_points_to_be<Bn><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const Affine<Bn::Fq>*)p->d + off, n, (uint8_t*)ctx->io_dev.p);
  ctx->launches++;
  if ((rc = launch_check(ctx, "k_points_to_be"))) return rc;
  BP_CUDA_OK(cudaMemcpyAsync(xy, ctx->io_dev.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA_OK(stream_sync(ctx));
  return BPGPU_OK;
}

size_t bpgpu_points_len(const bpgpu_points* p) { return p ? p->n : 0; }

int bpgpu_points_precompute(bpgpu_ctx* ctx, bpgpu_points* p) {
  if (!ctx || !p) return BPGPU_E_ARG;
  if (p->table || p->n == 0) return BPGPU_OK;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  return ctx->curve == BPGPU_BLS12_381 ? build_tables<Bls>(ctx, p->d, p->n, &p->table) : build_tables<Bn>(ctx, p->d, p->n, &p->table);
}
int bpgpu_points_precompute_wide(bpgpu_ctx* ctx, bpgpu_points* p) {
  if (!ctx || !p) return BPGPU_E_ARG;
  if (p->table16 || p->n == 0) return BPGPU_OK;
  int rc = bpgpu_points_precompute(ctx, p);
  if (rc) return rc;
  return ctx->curve == BPGPU_BLS12_381 ? build_tables16<Bls>(ctx, p->table, p->n, &p->table16) : build_tables16<Bn>(ctx, p->table, p->n, &p->table16);
}
int bpgpu_points_has_tables(const bpgpu_points* p) { return p && p->table ? (p->table16 ? 2 : 1) : 0; }
void bpgpu_points_free(bpgpu_points* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  dev_free(p->ctx, p->d);
  if (p->table) cudaFree(p->table);
  if (p->table16) cudaFree(p->table16);
  delete p;
}

int bpgpu_scalars_upload(bpgpu_ctx* ctx, const uint8_t* be, size_t n, bpgpu_scalars** out) {
  if (!ctx || !out || (!be && n)) return BPGPU_E_ARG;
  *out = nullptr;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  bpgpu_scalars* s = new (std::nothrow) bpgpu_scalars();
  if (!s) return BPGPU_E_CUDA;
  s->ctx = ctx; s->n = n; s->d = nullptr;
  if (dev_alloc(ctx, &s->d, n * 32) != cudaSuccess) { delete s; return BPGPU_E_CUDA; }
#define CALL(C) scalars_upload_t<C>(ctx, be, n, 1, s->d, ctx->io_dev)
  int rc = DISPATCH(ctx, CALL);
#undef CALL
  if (rc == BPGPU_OK && stream_sync(ctx) != cudaSuccess) rc = BPGPU_E_CUDA;
  if (rc) { dev_free(ctx, s->d); delete s; return rc; }
  *out = s;
  return BPGPU_OK;
}

int bpgpu_scalars_download(bpgpu_ctx* ctx, const bpgpu_scalars* s, size_t off, size_t n, uint8_t* be) {
  if (!ctx || !s || (!be && n)) return BPGPU_E_ARG;
  if (off > s->n || n > s->n - off) return BPGPU_E_ARG;
  if (n == 0) return BPGPU_OK;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  size_t bytes = n * bpgpu_modbytes(ctx->curve);
  int rc = ctx->io_dev.reserve(bytes);
  if (rc) return rc;
  if (ctx->curve == BPGPU_BLS12_381)
    k_scalars_to_be<Bls><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const Bls::Fr*)s->d + off, n, (uint8_t*)ctx->io_dev.p);
  else
    k_scalars_to_be<Bn><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const Bn::Fr*)s->d + off, n, (uint8_t*)ctx->io_dev.p);
  ctx->launches++;
  if ((rc = launch_check(ctx, "k_scalars_to_be"))) return rc;
  BP_CUDA_OK(cudaMemcpyAsync(be, ctx->io_dev.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA_OK(stream_sync(ctx));
  return BPGPU_OK;
}

size_t bpgpu_scalars_len(const bpgpu_scalars* s) { return s ? s->n : 0; }
void bpgpu_scalars_free(bpgpu_scalars* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  if (s->blk) {
    if (--s->blk->refs == 0) { dev_free(s->ctx, s->blk->base); delete s->blk; }
  } else {
    dev_free(s->ctx, s->d);
  }
  delete s;
}

int bpgpu_scalars_view(bpgpu_scalars* s, size_t off, size_t n, bpgpu_scalars** out) {
  if (!s || !out) return BPGPU_E_ARG;
  *out = nullptr;
  if (off > s->n || n > s->n - off) return BPGPU_E_LEN;
  if (!s->blk) {                                   // first view: the allocation becomes shared, s holds one reference
    s->blk = new (std::nothrow) bpgpu_shared_block{s->d, 1};
    if (!s->blk) return BPGPU_E_CUDA;
  }
  bpgpu_scalars* v = new (std::nothrow) bpgpu_scalars();
  if (!v) return BPGPU_E_CUDA;
  v->ctx = s->ctx; v->d = (uint8_t*)s->d + off * 32; v->n = n; v->blk = s->blk;
  s->blk->refs++;
  *out = v;
  return BPGPU_OK;
}

// ------------------------------------------------------------------ MSM
int bpgpu_msm_window_bits(size_t n) { return msm_window_bits(n); }


// ---- begin halves (see msm_mixed_begin): host buffers passed here must stay valid until bpgpu_msm_finish
int bpgpu_msm_begin(bpgpu_ctx* ctx, const bpgpu_points* p, size_t off, size_t n, const uint8_t* scalars_be) {
  if (!ctx || !p || (!scalars_be && n)) return BPGPU_E_ARG;
  if (off > p->n || n > p->n - off) return BPGPU_E_LEN;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  int rc = ctx->msm_c.reserve(n * 32 + 32);
  if (rc) return rc;
#define CALL(C) scalars_upload_t<C>(ctx, scalars_be, n, 0, ctx->msm_c.p, ctx->io_dev2)
  rc = DISPATCH(ctx, CALL);
#undef CALL
  if (rc) return rc;
  size_t psz = (ctx->curve == BPGPU_BLS12_381) ? sizeof(Affine<Bls::Fq>) : sizeof(Affine<Bn::Fq>);
  if (p->table && n) {
    TableSeg seg{(const uint8_t*)p->table + off * TBL_ENTRIES * psz, ctx->msm_c.p, (uint32_t)n, 0};
    return msm_mixed_begin(ctx, &seg, 1, nullptr, nullptr, false, 0);
  }
  return msm_mixed_begin(ctx, nullptr, 0, (const uint8_t*)p->d + off * psz, ctx->msm_c.p, false, n);
}

// resident bases, scalars as 32-byte LITTLE-endian canonical integers (< r): exactly the device's limb layout, so the
// upload is one copy of 32 bytes per term and no conversion launch (48-byte big-endian scalars cost 50 % more PCIe
// traffic on BLS12-381 plus a pass over them)
int bpgpu_msm_le32_begin(bpgpu_ctx* ctx, const bpgpu_points* p, size_t off, size_t n, const uint8_t* scalars_le32) {
  if (!ctx || !p || (!scalars_le32 && n)) return BPGPU_E_ARG;
  if (off > p->n || n > p->n - off) return BPGPU_E_LEN;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  int rc = ctx->msm_c.reserve(n * 32 + 32);
  if (rc) return rc;
  if (n) BP_CUDA_OK(cudaMemcpyAsync(ctx->msm_c.p, scalars_le32, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  size_t psz = (ctx->curve == BPGPU_BLS12_381) ? sizeof(Affine<Bls::Fq>) : sizeof(Affine<Bn::Fq>);
  if (p->table && n) {
    TableSeg seg{(const uint8_t*)p->table + off * TBL_ENTRIES * psz, ctx->msm_c.p, (uint32_t)n, 0};
    return msm_mixed_begin(ctx, &seg, 1, nullptr, nullptr, false, 0);
  }
  return msm_mixed_begin(ctx, nullptr, 0, (const uint8_t*)p->d + off * psz, ctx->msm_c.p, false, n);
}

int bpgpu_msm_device_begin(bpgpu_ctx* ctx, const bpgpu_points* p, size_t poff, size_t n, const bpgpu_scalars* s, size_t soff) {
  if (!ctx || !p || !s) return BPGPU_E_ARG;
  if (poff > p->n || n > p->n - poff || soff > s->n || n > s->n - soff) return BPGPU_E_LEN;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  size_t psz = (ctx->curve == BPGPU_BLS12_381) ? sizeof(Affine<Bls::Fq>) : sizeof(Affine<Bn::Fq>);
  if (p->table && n) {
    TableSeg seg{(const uint8_t*)p->table + poff * TBL_ENTRIES * psz, (const uint8_t*)s->d + soff * 32, (uint32_t)n, 1};
    return msm_mixed_begin(ctx, &seg, 1, nullptr, nullptr, false, 0);
  }
  return msm_mixed_begin(ctx, nullptr, 0, (const uint8_t*)p->d + poff * psz, (const uint8_t*)s->d + soff * 32, true, n);
}

int bpgpu_msm_refs_begin(bpgpu_ctx* ctx, const uint8_t* points_xy, const uint8_t* scalars_be, size_t n) {
  if (!ctx || ((!points_xy || !scalars_be) && n)) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  size_t psz = (ctx->curve == BPGPU_BLS12_381) ? sizeof(Affine<Bls::Fq>) : sizeof(Affine<Bn::Fq>);
  int rc = ctx->msm_d.reserve(n * psz + 32);
  if (rc) return rc;
  if ((rc = ctx->msm_c.reserve(n * 32 + 32))) return rc;
  // Scalars first: the digit / sort stages need only them.  For large n the points follow on the copy queue, so their
  // transfer and conversion run under k_digits / k_scan / k_scatter; k_chunk_acc waits for `points_ready`.
  // (Every entry point returns with the ctx idle or with ONE pending MSM whose finish precedes the next begin, so the
  // staging buffers are free when we get here.)
  const bool overlap = n >= ((size_t)1 << 14);
#define CALL(C) scalars_upload_t<C>(ctx, scalars_be, n, 0, ctx->msm_c.p, ctx->io_dev2)
  rc = DISPATCH(ctx, CALL);
#undef CALL
  if (rc) return rc;
#define CALL(C) points_from_host_on<C>(ctx, overlap ? ctx->copy_stream : ctx->stream, points_xy, n, ctx->msm_d.p)
  rc = DISPATCH(ctx, CALL);
#undef CALL
  if (rc) { if (overlap) cudaStreamSynchronize(ctx->copy_stream); return rc; }
  if (overlap) {
    BP_CUDA_OK(cudaEventRecord(ctx->points_ready, ctx->copy_stream));
    ctx->wait_points = true;
  }
  rc = msm_mixed_begin(ctx, nullptr, 0, ctx->msm_d.p, ctx->msm_c.p, false, n);
  if (rc && ctx->wait_points) {            // msm_run did not get as far as the wait (error path): drain the copy queue
    ctx->wait_points = false;
    cudaStreamSynchronize(ctx->copy_stream);
  }
  return rc;
}

int bpgpu_msm_finish(bpgpu_ctx* ctx, uint8_t* out_xy) {
  if (!ctx || !out_xy) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  return msm_mixed_finish(ctx, out_xy);
}

int bpgpu_msm(bpgpu_ctx* ctx, const bpgpu_points* p, size_t off, size_t n, const uint8_t* scalars_be, uint8_t* out_xy) {
  if (!out_xy) return BPGPU_E_ARG;
  int rc = bpgpu_msm_begin(ctx, p, off, n, scalars_be);
  return rc ? rc : bpgpu_msm_finish(ctx, out_xy);
}

int bpgpu_msm_le32(bpgpu_ctx* ctx, const bpgpu_points* p, size_t off, size_t n, const uint8_t* scalars_le32, uint8_t* out_xy) {
  if (!out_xy) return BPGPU_E_ARG;
  int rc = bpgpu_msm_le32_begin(ctx, p, off, n, scalars_le32);
  return rc ? rc : bpgpu_msm_finish(ctx, out_xy);
}

int bpgpu_msm_device(bpgpu_ctx* ctx, const bpgpu_points* p, size_t poff, size_t n, const bpgpu_scalars* s, size_t soff,
                     uint8_t* out_xy) {
  if (!out_xy) return BPGPU_E_ARG;
  int rc = bpgpu_msm_device_begin(ctx, p, poff, n, s, soff);
  return rc ? rc : bpgpu_msm_finish(ctx, out_xy);
}

int bpgpu_msm_refs(bpgpu_ctx* ctx, const uint8_t* points_xy, const uint8_t* scalars_be, size_t n, uint8_t* out_xy) {
  if (!out_xy) return BPGPU_E_ARG;
  int rc = bpgpu_msm_refs_begin(ctx, points_xy, scalars_be, n);
  return rc ? rc : bpgpu_msm_finish(ctx, out_xy);
}

// ------------------------------------------------------------------ self tests
int bpgpu_selftest_field(bpgpu_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  if (!ctx || !a || !b || !out) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  int mb = bpgpu_modbytes(ctx->curve);
  size_t bytes = n * mb;
  int rc = ctx->io_dev.reserve(bytes * 3);
  if (rc) return rc;
  uint8_t* da = (uint8_t*)ctx->io_dev.p; uint8_t* db = da + bytes; uint8_t* dout = db + bytes;
  BP_CUDA_OK(cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
  BP_CUDA_OK(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
  unsigned blocks = (unsigned)((n + 63) / 64);
  if (ctx->curve == BPGPU_BLS12_381) {
    if (field == 0) k_field_op<Bls::Fq><<<blocks, 64, 0, ctx->stream>>>(op, da, db, n, mb, dout);
    else k_field_op<Bls::Fr><<<blocks, 64, 0, ctx->stream>>>(op, da, db, n, mb, dout);
  } else {
    if (field == 0) k_field_op<Bn::Fq><<<blocks, 64, 0, ctx->stream>>>(op, da, db, n, mb, dout);
    else k_field_op<Bn::Fr><<<blocks, 64, 0, ctx->stream>>>(op, da, db, n, mb, dout);
  }
  ctx->launches++;
  if ((rc = launch_check(ctx, "k_field_op"))) return rc;
  BP_CUDA_OK(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA_OK(stream_sync(ctx));
  return BPGPU_OK;
}

int bpgpu_selftest_group(bpgpu_ctx* ctx, int op, const uint8_t* p_xy, const uint8_t* q_xy, const uint8_t* scalars_be, size_t n,
                         uint8_t* out_xy) {
  if (!ctx || !p_xy || !q_xy || !scalars_be || !out_xy) return BPGPU_E_ARG;
  bpgpu_points *P = nullptr, *Q = nullptr, *O = nullptr;
  int rc = bpgpu_points_upload(ctx, p_xy, n, &P);
  if (!rc) rc = bpgpu_points_upload(ctx, q_xy, n, &Q);
  if (!rc) rc = bpgpu_points_upload(ctx, p_xy, n, &O);
  if (!rc) rc = ctx->msm_c.reserve(n * 32 + 32);
  if (!rc) {
#define CALL(C) scalars_upload_t<C>(ctx, scalars_be, n, 0, ctx->msm_c.p, ctx->io_dev2)
    rc = DISPATCH(ctx, CALL);
#undef CALL
  }
  if (!rc) {
    unsigned blocks = (unsigned)((n + 63) / 64);
    if (ctx->curve == BPGPU_BLS12_381)
      k_group_op<Bls><<<blocks, 64, 0, ctx->stream>>>(op, (const Affine<Bls::Fq>*)P->d, (const Affine<Bls::Fq>*)Q->d,
                                                      (const Bls::Fr*)ctx->msm_c.p, n, (Affine<Bls::Fq>*)O->d);
    else
      k_group_op<Bn><<<blocks, 64, 0, ctx->stream>>>(op, (const Affine<Bn::Fq>*)P->d, (const Affine<Bn::Fq>*)Q->d,
                                                     (const Bn::Fr*)ctx->msm_c.p, n, (Affine<Bn::Fq>*)O->d);
    ctx->launches++;
    rc = launch_check(ctx, "k_group_op");
  }
  if (!rc) rc = bpgpu_points_download(ctx, O, 0, n, out_xy);
  bpgpu_points_free(P); bpgpu_points_free(Q); bpgpu_points_free(O);
  return rc;
}

int bpgpu_int_pipe_bench(bpgpu_ctx* ctx, int kind, int iters, double* ops_per_s, double* ms_out) {
  if (!ctx || !ops_per_s) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  int rc = ctx->io_dev.reserve(1 << 16);
  if (rc) return rc;
  cudaEvent_t e0, e1;
  BP_CUDA_OK(cudaEventCreate(&e0));
  BP_CUDA_OK(cudaEventCreate(&e1));
  int blocks = ctx->sm_count * 8, threads = 256;
  if (kind == 31 || kind == 32 || kind == 33) { blocks = 1; threads = 32; }     // single-warp latency probes
  double ops = 0;
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    BP_CUDA_OK(cudaEventRecord(e0, ctx->stream));
    if (kind == 0) {
      k_imad_wide<<<blocks, threads, 0, ctx->stream>>>(iters, 12345u + rep, (uint64_t*)ctx->io_dev.p);
      ops = (double)blocks * threads * (double)iters * 64.0;
    } else if (kind == 2 || kind == 3) {
      k_imad_lo<<<blocks, threads, 0, ctx->stream>>>(iters, 12345u + rep, kind == 3, (uint64_t*)ctx->io_dev.p);
      ops = (double)blocks * threads * (double)iters * 4.0 * (kind == 3 ? 24.0 : 16.0);   // instructions
    } else {
      BP_CUDA_OK(cudaMemsetAsync(ctx->io_dev.p, 0x5a, 1 << 16, ctx->stream));
      const Bls::Fq* in = (const Bls::Fq*)ctx->io_dev.p;
      Bls::Fq* o = (Bls::Fq*)ctx->io_dev.p;
      switch (kind) {
        case 11: k_fq_mul_chain<Bls::Fq, 1><<<blocks, threads, 0, ctx->stream>>>(iters, in, o); break;
        case 12: k_fq_mul_chain<Bls::Fq, 2><<<blocks, threads, 0, ctx->stream>>>(iters, in, o); break;
        case 13: k_fq_mul_chain<Bls::Fq, 3><<<blocks, threads, 0, ctx->stream>>>(iters, in, o); break;
        case 14: k_fq_mul_chain<Bls::Fq, 4><<<blocks, threads, 0, ctx->stream>>>(iters, in, o); break;
        case 16: k_fq_mul_chain<Bls::Fq, 6><<<blocks, threads, 0, ctx->stream>>>(iters, in, o); break;
        case 31: k_fq_mul_chain<Bls::Fq, 0><<<blocks, threads, 0, ctx->stream>>>(iters, in, o); break;
        case 32: k_fq_mul_chain<Bn::Fq, 0><<<blocks, threads, 0, ctx->stream>>>(iters, (const Bn::Fq*)ctx->io_dev.p, (Bn::Fq*)ctx->io_dev.p); break;
        case 33: k_dbl_chain<Bls::Fq><<<blocks, threads, 0, ctx->stream>>>(iters, (const XYZZ<Bls::Fq>*)ctx->io_dev.p, (XYZZ<Bls::Fq>*)ctx->io_dev.p); break;
        case 20: k_fq_mul_chain<Bn::Fq, 0><<<blocks, threads, 0, ctx->stream>>>(iters, (const Bn::Fq*)ctx->io_dev.p, (Bn::Fq*)ctx->io_dev.p); break;
        case 22: k_fq_mul_chain<Bn::Fq, 2><<<blocks, threads, 0, ctx->stream>>>(iters, (const Bn::Fq*)ctx->io_dev.p, (Bn::Fq*)ctx->io_dev.p); break;
        default: k_fq_mul_chain<Bls::Fq, 0><<<blocks, threads, 0, ctx->stream>>>(iters, in, o); break;
      }
      ops = (double)blocks * threads * (double)iters * 2.0;
    }
    ctx->launches++;
    BP_CUDA_OK(cudaEventRecord(e1, ctx->stream));
    BP_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    BP_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if ((rc = launch_check(ctx, "int_pipe_bench"))) return rc;
  *ops_per_s = ops / (best * 1e-3);
  if (ms_out) *ms_out = best;
  return BPGPU_OK;
}

}  // extern "C"
