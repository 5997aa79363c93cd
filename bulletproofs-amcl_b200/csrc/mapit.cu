// Batched hash-to-curve on the device: the map step of G1::from_msg_hash, i.e. AMCL's ECP::mapit applied to
// SHAKE256(msg)[..MODBYTES], for whole generator tables at once.
//
// Replaces the per-generator host loop of utils::get_generators (/root/reference/src/utils/mod.rs:16-23), which the
// reference's own comments call out as slow (gadgets/randomizer.rs:431).  One thread per generator:
//   x = be(h) mod p ; repeat { y^2 = x^3 + b ; y = (y^2)^((p+1)/4) ; x += 1 } until y is a root ;
//   take the root whose canonical value is even ; clear the cofactor ; retry if that gave the identity ;
// then normalise to affine.  The SHAKE256 hash of the message stays on the host (hash_msg).
#include "common.cuh"

namespace bp {

template <class Curve>
struct MapConsts;
template <>
struct MapConsts<Bls> {
  static constexpr int B = 4;
  static constexpr int COF_LIMBS = 4;
  __device__ static uint32_t cof(int i) { constexpr uint32_t v[4] = {BLS_COF[0], BLS_COF[1], BLS_COF[2], BLS_COF[3]}; return v[i]; }
};
template <>
struct MapConsts<Bn> {
  static constexpr int B = 2;
  static constexpr int COF_LIMBS = 0;      // cofactor 1
  __device__ static uint32_t cof(int) { return 1; }
};

template <class Curve>
__global__ void __launch_bounds__(64) k_mapit(const uint8_t* __restrict__ hashes, size_t n, Affine<typename Curve::Fq>* __restrict__ out) {
  using Fq = typename Curve::Fq;
  using FqP = typename std::conditional<Curve::ID == BPGPU_BLS12_381, BlsFq, BnFq>::type;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x;
  be_to_limbs<Fq::N>(hashes + i * Curve::MODBYTES, Curve::MODBYTES, x.v);
  // x.rmod(p): 2^(32N)/p < 10 for both primes
#pragma unroll 1
  for (int k = 0; k < 10; k++) Fq::reduce_once(x.v);
  x = x.to_mont();
  Fq bcoef = Fq::zero();
  for (int k = 0; k < MapConsts<Curve>::B; k++) bcoef = bcoef + Fq::one();
  XYZZ<Fq> P = XYZZ<Fq>::inf();
#pragma unroll 1
  for (int tries = 0; tries < 4096; tries++) {
    Fq rhs = x.sqr() * x + bcoef;
    Fq y = rhs.pow_limbs([](int k) { return FqP::SQRT_E(k); }, Fq::N);
    bool is_root = (y.sqr() == rhs);
    Fq xc = x;
    x = x + Fq::one();                                   // x.inc(1) happens whether or not a point was found
    if (!is_root) continue;
    if (y.from_mont().v[0] & 1) y = y.neg();             // ECP::new_bigint(x, 0): the root with even canonical value
    Affine<Fq> A; A.x = xc; A.y = y;
    XYZZ<Fq> Q = XYZZ<Fq>::from_affine(A);
    if (MapConsts<Curve>::COF_LIMBS > 0) {               // P.cfp()
      uint32_t c[4];
      for (int k = 0; k < 4; k++) c[k] = MapConsts<Curve>::cof(k);
      Q = mul_limbs(Q, c, MapConsts<Curve>::COF_LIMBS);
    }
    if (!Q.is_inf()) { P = Q; break; }
  }
  store_vec(out + i, P.to_affine());
}

template <class Curve>
static int mapit_t(bpgpu_ctx* ctx, const uint8_t* hashes, size_t n, void* dst) {
  if (n == 0) return BPGPU_OK;
  size_t bytes = n * Curve::MODBYTES;
  int rc = ctx->io_dev.reserve(bytes);
  if (rc) return rc;
  BP_CUDA_OK(cudaMemcpyAsync(ctx->io_dev.p, hashes, bytes, cudaMemcpyHostToDevice, ctx->stream));
  k_mapit<Curve><<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>((const uint8_t*)ctx->io_dev.p, n, (Affine<typename Curve::Fq>*)dst);
  ctx->launches++;
  return launch_check(ctx, "k_mapit");
}

}  // namespace bp

using namespace bp;

extern "C" int bpgpu_points_from_hashes(bpgpu_ctx* ctx, const uint8_t* hashes, size_t n, bpgpu_points** out) {
  if (!ctx || !out || (!hashes && n)) return BPGPU_E_ARG;
  *out = nullptr;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  bpgpu_points* p = new (std::nothrow) bpgpu_points();
  if (!p) return BPGPU_E_CUDA;
  p->ctx = ctx; p->n = n; p->d = nullptr;
  size_t psz = (ctx->curve == BPGPU_BLS12_381) ? sizeof(Affine<Bls::Fq>) : sizeof(Affine<Bn::Fq>);
  if (dev_alloc(ctx, &p->d, n * psz) != cudaSuccess) { delete p; return BPGPU_E_CUDA; }
  int rc = ctx->curve == BPGPU_BLS12_381 ? mapit_t<Bls>(ctx, hashes, n, p->d) : mapit_t<Bn>(ctx, hashes, n, p->d);
  if (rc == BPGPU_OK && stream_sync(ctx) != cudaSuccess) rc = BPGPU_E_CUDA;
  if (rc) { dev_free(ctx, p->d); delete p; return rc; }
  *out = p;
  return BPGPU_OK;
}
