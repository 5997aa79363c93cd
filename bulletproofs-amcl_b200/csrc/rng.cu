// FieldElementVector::random on the device: the prover's blinding vectors s_L, s_R
// (/root/reference/src/r1cs/prover.rs:340-341, 401-402: n1 + n1 (+ n2 + n2) calls of FieldElement::random()).
//
// The host layer's Rng is a counter-mode stream, scalar_i = be_int(SHAKE256(key || le64(i))[..MODBYTES]) mod r
// (host/curve.hpp; key = seed_le64 || "blind" for the deterministic streams the oracle and the tests use, 32 bytes of OS
// entropy otherwise).  Counter mode means element i needs nobody else's state: one thread per element runs one
// Keccak-f[1600] (the message is far shorter than the 136-byte rate), reduces the 384- or 256-bit output mod r and
// leaves the Montgomery form in HBM -- the vector is born where the commitment MSM and the polynomial kernels read it,
// and the host neither generates, packs nor uploads 2n scalars.
#include "common.cuh"

namespace bp {

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

__device__ void keccak_f1600_dev(uint64_t* a) {
  const uint64_t RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
                           0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
                           0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
                           0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
                           0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
                           0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
#pragma unroll 1
  for (int r = 0; r < 24; r++) {
    uint64_t c[5], b[25];
#pragma unroll
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma unroll
    for (int x = 0; x < 5; x++) {
      const uint64_t d = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
#pragma unroll
      for (int y = 0; y < 5; y++) a[x + 5 * y] ^= d;
    }
#pragma unroll
    for (int x = 0; x < 5; x++)
#pragma unroll
      for (int y = 0; y < 5; y++) {
        const int i = x + 5 * y;
        b[y + 5 * ((2 * x + 3 * y) % 5)] = ROT[i] ? rotl64(a[i], ROT[i]) : a[i];
      }
#pragma unroll
    for (int y = 0; y < 5; y++)
#pragma unroll
      for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= RC[r];
  }
}

struct RngKey { uint8_t b[64]; uint32_t len; };

// be_int(SHAKE256(key || le64(ctr))[..MODBYTES]) mod r, Montgomery form
template <class Curve>
__device__ typename Curve::Fr fr_from_shake(const uint8_t* key, uint32_t klen, uint64_t ctr) {
  using Fr = typename Curve::Fr;
  // SHAKE256 of one short block: msg || 0x1f .. 0x80 at the end of the 136-byte rate
  uint8_t m[136];
#pragma unroll 1
  for (int k = 0; k < 136; k++) m[k] = 0;
  uint32_t len = klen;
#pragma unroll 1
  for (uint32_t k = 0; k < len; k++) m[k] = key[k];
#pragma unroll 1
  for (int k = 0; k < 8; k++) m[len + k] = (uint8_t)(ctr >> (8 * k));
  len += 8;
  m[len] ^= 0x1f;
  m[135] ^= 0x80;
  uint64_t st[25];
#pragma unroll 1
  for (int l = 0; l < 25; l++) {
    uint64_t v = 0;
    if (l < 17) for (int k = 7; k >= 0; k--) v = (v << 8) | m[8 * l + k];
    st[l] = v;
  }
  keccak_f1600_dev(st);
  // first MODBYTES output bytes as a big-endian integer
  constexpr int MB = Curve::MODBYTES;
  uint8_t ob[MB];
#pragma unroll 1
  for (int k = 0; k < MB; k++) ob[k] = (uint8_t)(st[k >> 3] >> (8 * (k & 7)));
  uint32_t limbs[MB / 4];
  be_to_limbs<MB / 4>(ob, MB, limbs);
  Fr lo;
#pragma unroll
  for (int k = 0; k < 8; k++) lo.v[k] = limbs[k];
  canonicalise(lo);                              // < 2^256 -> < r
  Fr res = lo * Fr::r2();                        // Montgomery form
  if (MB > 32) {                                 // + hi * 2^256: in Montgomery form hi * R * R = (hi * R2 / R) * R2 / R with R = 2^256
    Fr hi = Fr::zero();
#pragma unroll
    for (int k = 8; k < MB / 4; k++) hi.v[k - 8] = limbs[k];
    res = res + (hi * Fr::r2()) * Fr::r2();
  }
  return res;
}

template <class Curve>
__global__ void __launch_bounds__(128) k_fr_random(RngKey key, uint64_t ctr0, size_t n, typename Curve::Fr* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  store_vec(out + i, fr_from_shake<Curve>(key.b, key.len, ctr0 + i));
}

// one stream per proof of a batch (blockIdx.y): keys = batch x 64 bytes (device), out = batch x n
template <class Curve>
__global__ void __launch_bounds__(128) k_fr_random_batch(const uint8_t* __restrict__ keys, uint32_t klen, const uint64_t* __restrict__ ctr0,
                                                         size_t n, typename Curve::Fr* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t b = blockIdx.y;
  store_vec(out + b * n + i, fr_from_shake<Curve>(keys + b * 64, klen, ctr0[b] + i));
}

template <class Curve>
int fr_random_batch_run(bpgpu_ctx* ctx, const uint8_t* d_keys, uint32_t klen, const uint64_t* d_ctr0, size_t batch, size_t n, void* d_out) {
  if (!batch || !n) return BPGPU_OK;
  k_fr_random_batch<Curve><<<dim3((unsigned)((n + 127) / 128), (unsigned)batch), 128, 0, ctx->stream>>>(d_keys, klen, d_ctr0, n,
                                                                                                        (typename Curve::Fr*)d_out);
  ctx->launches++;
  return launch_check(ctx, "k_fr_random_batch");
}
template int fr_random_batch_run<Bls>(bpgpu_ctx*, const uint8_t*, uint32_t, const uint64_t*, size_t, size_t, void*);
template int fr_random_batch_run<Bn>(bpgpu_ctx*, const uint8_t*, uint32_t, const uint64_t*, size_t, size_t, void*);

}  // namespace bp

using namespace bp;

extern "C" int bpgpu_fr_random(bpgpu_ctx* ctx, const uint8_t* key, size_t key_len, uint64_t ctr0, size_t n, bpgpu_scalars** out) {
  if (!ctx || !out || (!key && key_len) || key_len > 64) return BPGPU_E_ARG;
  if (n >= (1ull << 32)) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  int rc = bpgpu_scalars_alloc(ctx, n, out);
  if (rc || n == 0) return rc;
  RngKey k;
  memset(k.b, 0, sizeof k.b);
  if (key_len) memcpy(k.b, key, key_len);
  k.len = (uint32_t)key_len;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  if (ctx->curve == BPGPU_BLS12_381) k_fr_random<Bls><<<blocks, 128, 0, ctx->stream>>>(k, ctr0, n, (Bls::Fr*)(*out)->d);
  else k_fr_random<Bn><<<blocks, 128, 0, ctx->stream>>>(k, ctr0, n, (Bn::Fr*)(*out)->d);
  ctx->launches++;
  rc = launch_check(ctx, "k_fr_random");
  if (rc) { bpgpu_scalars_free(*out); *out = nullptr; }
  return rc;
}
