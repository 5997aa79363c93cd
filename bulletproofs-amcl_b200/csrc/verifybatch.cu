// Batched R1CS verification with everything but the upload on the device.
//
// Verifier::verify (/root/reference/src/r1cs/verifier.rs:267-457) for `count` independent proofs of ONE one-phase circuit:
// the reference (and round 1 of this library) builds, per proof, O(N) scalars on the host -- flattened_constraints
// (:149-193), y^-i, y_inv_wR, delta, g/h scalars (:341-390), the s vector (ipp.rs:295-312) -- which left eight GPUs
// waiting on four host cores each.  Here the host uploads the proof and commitment BYTES; per slab of <= 4096 proofs
//   k_vb_transcript : one thread per proof: Merlin replay (csrc/merlin.cuh) -> y, z, u, x, w, u_k; proof scalars;
//                     one inversion per proof for all u_k^-1 and y^-1  (or: challenges taken from the caller's transcripts)
//   k_vb_points     : one thread per proof point: decode, range and on-curve checks (ECP::frombytes), Montgomery form
//   k_vb_scalars    : one block per proof: the circuit's CSR times the powers of z, s, y^-i, g/h scalars, delta, head
//   batch.cu        : the 2N+2 fixed terms from window tables, the proof's own points by Straus, verdict byte
// and one verdict per proof returns.  Nothing is merged across proofs (the reference has no batch API).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "batchsum.cuh"
#include "host_fp.h"
#include "verify_core.cuh"

namespace bp {

struct VbKey { uint8_t b[64]; uint32_t len; };

template <class Curve>
__global__ void __launch_bounds__(64) k_vb_transcript(uint32_t cnt, const uint8_t* __restrict__ state0, const uint8_t* __restrict__ chal,
                                                      const uint8_t* __restrict__ proofs, uint32_t plen, const uint8_t* __restrict__ comms,
                                                      uint32_t m, uint32_t lg, const __grid_constant__ VbKey key, uint64_t ctr0,
                                                      typename Curve::Fr* __restrict__ hdr, int32_t* __restrict__ status) {
  using Fr = typename Curve::Fr;
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= cnt) return;
  const uint32_t hl = vb_hdr_len(lg);
  Fr* h = hdr + (size_t)b * hl;
  const uint8_t* proof = proofs + (size_t)b * plen;
  const Fr r = fr_stream_draw<Curve>(key.b, key.len, ctr0 + b);
  int st;
  if (state0) st = vb_replay<Curve>(state0, proof, comms + (size_t)b * m * 2 * Curve::MODBYTES, m, lg, (uint64_t)1 << lg, r, h);
  else st = vb_from_challenges<Curve>(chal + (size_t)b * (VB_CH_FIXED + lg) * Curve::MODBYTES, proof, lg, r, h);
  if (st)
    for (uint32_t k = 0; k < hl; k++) h[k] = Fr::zero();      // the verdict is decided; keep the later stages well defined
  status[b] = st;
}

template <class Curve>
__global__ void __launch_bounds__(128) k_vb_points(uint32_t cnt, uint32_t vn, uint32_t m, uint32_t lg, const uint8_t* __restrict__ proofs,
                                                   uint32_t plen, const uint8_t* __restrict__ comms,
                                                   Affine<typename Curve::Fq>* __restrict__ vp, int32_t* __restrict__ status) {
  using Fq = typename Curve::Fq;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)cnt * vn) return;
  const uint32_t b = (uint32_t)(t / vn), k = (uint32_t)(t - (size_t)b * vn);
  const uint8_t* src = vb_var_point_bytes<Curve>(proofs + (size_t)b * plen, comms + (size_t)b * m * 2 * Curve::MODBYTES, m, lg, k);
  Affine<Fq> a;
  if (!g1_from_be_checked<Curve>(src, &a)) {
    a = Affine<Fq>::inf();
    atomicCAS(&status[b], 0, BPGPU_E_FORMAT);
  }
  store_vec(vp + t, a);
}

template <class Curve>
__global__ void __launch_bounds__(128) k_vb_scalars(const __grid_constant__ CircuitDev c, uint32_t N, uint32_t lg, uint32_t F, uint32_t vn,
                                                    const typename Curve::Fr* __restrict__ hdr, typename Curve::Fr* __restrict__ fs,
                                                    typename Curve::Fr* __restrict__ vs) {
  using Fr = typename Curve::Fr;
  __shared__ __align__(16) unsigned char sm_hdr_raw[(VB_HDR + 64) * sizeof(Fr)];
  __shared__ __align__(16) unsigned char sm_tab_raw[64 * sizeof(Fr)];
  __shared__ __align__(16) unsigned char sm_red_raw[128 * sizeof(Fr)];
  Fr* sh = reinterpret_cast<Fr*>(sm_hdr_raw);
  Fr* yitab = reinterpret_cast<Fr*>(sm_tab_raw);
  Fr* ztab = yitab + 32;
  Fr* red = reinterpret_cast<Fr*>(sm_red_raw);
  const uint32_t b = blockIdx.x;
  const uint32_t hl = vb_hdr_len(lg);
  for (uint32_t k = threadIdx.x; k < hl; k += blockDim.x) store_vec(sh + k, load_vec(hdr + (size_t)b * hl + k));
  __syncthreads();
  if (threadIdx.x == 0) hd_square_table(sh[VB_YINV], yitab);
  if (threadIdx.x == blockDim.x - 1) hd_square_table(sh[VB_Z], ztab);
  __syncthreads();
  Fr* f = fs + (size_t)b * F;
  Fr* v = vs + (size_t)b * vn;
  Fr dsum = Fr::zero();
  for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
    Fr g, h, d;
    vb_gh_element(c, i, N, lg, sh, yitab, ztab, &g, &h, &d);
    store_vec(f + i, g.from_mont());
    store_vec(f + N + i, h.from_mont());
    dsum = dsum + d;
  }
  // the constants row has an entry per constraint with a constant term (n of them for range gadgets): split over the block
  const uint32_t wrow = 3 * c.n + c.m;
  const uint32_t wlo = c.row_start[wrow], whi = c.row_start[wrow + 1], wper = (whi - wlo + blockDim.x - 1) / blockDim.x;
  const uint32_t wa = wlo + threadIdx.x * wper;
  const Fr wpart = wa < whi ? csr_chunk_eval(c, wa, wa + wper < whi ? wa + wper : whi, ztab) : Fr::zero();
  auto block_sum = [&](const Fr& mine) {
    __syncthreads();
    store_vec(red + threadIdx.x, mine);
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) store_vec(red + threadIdx.x, load_vec(red + threadIdx.x) + load_vec(red + threadIdx.x + o));
      __syncthreads();
    }
    return load_vec(red);
  };
  const Fr delta = block_sum(dsum);
  const Fr wc = block_sum(wpart);
  for (uint32_t j = threadIdx.x; j < c.m; j += blockDim.x) store_vec(v + 6 + j, vb_var_wv(c, j, sh, ztab));
  if (threadIdx.x == blockDim.x - 1) {
    Fr fg, fh;
    vb_head(c.m, lg, sh, delta, wc, &fg, &fh, v);
    store_vec(f + 2 * N, fg);
    store_vec(f + 2 * N + 1, fh);
  }
}

// flattened_constraints for ONE challenge z: out = [wL | wR | wO | wV | wc] (3n + m + 1 values, Montgomery)
template <class Fr> struct ZTab { Fr t[32]; };
template <class Fr>
__global__ void __launch_bounds__(128) k_circuit_flatten(const __grid_constant__ CircuitDev c, const __grid_constant__ ZTab<Fr> ztab,
                                                         Fr* __restrict__ out) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= 3 * c.n + c.m) return;                      // the constants row: k_circuit_wc
  store_vec(out + row, csr_row_eval(c, row, ztab.t));
}
template <class Fr>
__global__ void __launch_bounds__(512) k_circuit_wc(const __grid_constant__ CircuitDev c, const __grid_constant__ ZTab<Fr> ztab, Fr* __restrict__ out) {
  __shared__ __align__(16) unsigned char red_raw[512 * sizeof(Fr)];
  Fr* red = reinterpret_cast<Fr*>(red_raw);
  const uint32_t wrow = 3 * c.n + c.m;
  const uint32_t lo = c.row_start[wrow], hi = c.row_start[wrow + 1];
  const uint32_t per = (hi - lo + blockDim.x - 1) / blockDim.x;
  const uint32_t a = lo + threadIdx.x * per;
  store_vec(red + threadIdx.x, a < hi ? csr_chunk_eval(c, a, a + per < hi ? a + per : hi, ztab.t) : Fr::zero());
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) store_vec(red + threadIdx.x, load_vec(red + threadIdx.x) + load_vec(red + threadIdx.x + o));
    __syncthreads();
  }
  if (threadIdx.x == 0) store_vec(out + wrow, load_vec(red));
}

// proofs per group of launches (bounds the scratch: 15 XYZZ multiples per proof point); BPGPU_VB_SLAB overrides (tuning runs)
static size_t verify_slab() {
  static const size_t v = getenv("BPGPU_VB_SLAB") && atol(getenv("BPGPU_VB_SLAB")) > 0 ? (size_t)atol(getenv("BPGPU_VB_SLAB")) : 4096;
  return v;
}

template <class Curve>
static int verify_batch_t(bpgpu_ctx* ctx, const bpgpu_circuit* circ, const FixedRuns& runs, size_t count, const uint8_t* proofs, size_t stride,
                          const uint8_t* comms_xy, const uint8_t* state0, const uint8_t* chal_be, const uint8_t* key, size_t klen,
                          int32_t* verdicts, uint8_t* dbg_fixed_be, uint8_t* dbg_var_be) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  using PL = ProofLayout<Curve>;
  constexpr size_t MB = Curve::MODBYTES;
  const CircuitDev& c = circ->dev;
  uint32_t N = 1, lg = 0;
  while (N < c.n) { N <<= 1; lg++; }
  const uint32_t F = 2 * N + 2, vn = 6 + c.m + 5 + 2 * lg, hl = vb_hdr_len(lg);
  const uint32_t plen = PL::len(lg), nch = VB_CH_FIXED + lg;
  const size_t SLAB = verify_slab();
  const size_t slab = count < SLAB ? count : SLAB;
  int rc;
  const BatchScratch L = batch_scratch_layout<Curve>(slab, F, vn);
  if ((rc = ctx->msm_b.reserve(L.total))) return rc;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_proofs = 0, o_comms = o_proofs + up(slab * plen), o_chal = o_comms + up(slab * c.m * 2 * MB),
               o_hdr = o_chal + up(chal_be ? slab * nch * MB : 0), o_status = o_hdr + up(slab * hl * sizeof(Fr)),
               o_state = o_status + up(slab * 4), o_end = o_state + 256;
  if ((rc = ctx->vb.reserve(o_end))) return rc;
  uint8_t* base = (uint8_t*)ctx->msm_b.p;
  uint8_t* vb = (uint8_t*)ctx->vb.p;
  uint8_t *d_proofs = vb + o_proofs, *d_comms = vb + o_comms, *d_chal = chal_be ? vb + o_chal : nullptr, *d_state = state0 ? vb + o_state : nullptr;
  Fr* d_hdr = (Fr*)(vb + o_hdr);
  int32_t* d_status = (int32_t*)(vb + o_status);
  VbKey k;
  memset(&k, 0, sizeof k);
  memcpy(k.b, key, klen);
  k.len = (uint32_t)klen;
  if (state0) BP_CUDA_OK(cudaMemcpyAsync(d_state, state0, MERLIN_STATE_BYTES, cudaMemcpyHostToDevice, ctx->stream));
  std::vector<uint8_t> ident(slab);
  std::vector<int32_t> status(slab);
  for (size_t lo = 0; lo < count; lo += slab) {
    const size_t cnt = count - lo < slab ? count - lo : slab;
    if (stride == plen) BP_CUDA_OK(cudaMemcpyAsync(d_proofs, proofs + lo * stride, cnt * plen, cudaMemcpyHostToDevice, ctx->stream));
    else BP_CUDA_OK(cudaMemcpy2DAsync(d_proofs, plen, proofs + lo * stride, stride, plen, cnt, cudaMemcpyHostToDevice, ctx->stream));
    if (c.m) BP_CUDA_OK(cudaMemcpyAsync(d_comms, comms_xy + lo * c.m * 2 * MB, cnt * c.m * 2 * MB, cudaMemcpyHostToDevice, ctx->stream));
    if (chal_be) BP_CUDA_OK(cudaMemcpyAsync(d_chal, chal_be + lo * nch * MB, cnt * nch * MB, cudaMemcpyHostToDevice, ctx->stream));
    k_vb_transcript<Curve><<<(unsigned)((cnt + 63) / 64), 64, 0, ctx->stream>>>((uint32_t)cnt, d_state, d_chal, d_proofs, plen, d_comms, c.m, lg, k,
                                                                                (uint64_t)lo, d_hdr, d_status);
    const size_t np = cnt * vn;
    k_vb_points<Curve><<<(unsigned)((np + 127) / 128), 128, 0, ctx->stream>>>((uint32_t)cnt, vn, c.m, lg, d_proofs, plen, d_comms,
                                                                              (Affine<Fq>*)(base + L.vp), d_status);
    const unsigned sbs = N >= 128 ? 128u : (N >= 32 ? N : 32u);      // a power of two; no idle half-block for the 64-multiplier circuits
    k_vb_scalars<Curve><<<(unsigned)cnt, sbs, 0, ctx->stream>>>(c, N, lg, F, vn, d_hdr, (Fr*)(base + L.fs), (Fr*)(base + L.vs));
    ctx->launches += 3;
    if ((rc = launch_check(ctx, "verify_batch"))) return rc;
    std::vector<uint32_t> dbg;
    if (dbg_fixed_be && lo == 0) {                 // test hook: the scalars of the first slab as the device built them (canonical integers),
      dbg.resize(cnt * (size_t)(F + vn) * 8);      // read before the MSM stage splits the variable ones in place (GLV)
      BP_CUDA_OK(cudaMemcpyAsync(dbg.data(), base + L.fs, cnt * (size_t)F * 32, cudaMemcpyDeviceToHost, ctx->stream));
      BP_CUDA_OK(cudaMemcpyAsync(dbg.data() + cnt * (size_t)F * 8, base + L.vs, cnt * (size_t)vn * 32, cudaMemcpyDeviceToHost, ctx->stream));
      BP_CUDA_OK(stream_sync(ctx));
      for (size_t i = 0; i < cnt * (size_t)F; i++) hd_limbs_to_be<8>(dbg.data() + i * 8, (int)MB, dbg_fixed_be + i * MB);
      for (size_t i = 0; i < cnt * (size_t)vn; i++) hd_limbs_to_be<8>(dbg.data() + (cnt * (size_t)F + i) * 8, (int)MB, dbg_var_be + i * MB);
    }
    if ((rc = batch_identity_launch<Curve>(ctx, runs, F, cnt, base + L.fs, base + L.vp, base + L.vs, vn, base + L.sum, base + L.m, base + L.w,
                                           base + L.v)))
      return rc;
    BP_CUDA_OK(cudaMemcpyAsync(ident.data(), base + L.v, cnt, cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA_OK(cudaMemcpyAsync(status.data(), d_status, cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA_OK(stream_sync(ctx));
    for (size_t i = 0; i < cnt; i++) verdicts[lo + i] = status[i] ? status[i] : (ident[i] ? BPGPU_OK : BPGPU_E_VERIFY);
  }
  return BPGPU_OK;
}

template <class Curve>
static int circuit_flatten_t(bpgpu_ctx* ctx, const bpgpu_circuit* circ, const uint8_t* z_be, void* d_out) {
  using Fr = typename Curve::Fr;
  using HF = host::HFp<typename Curve::FrParams>;
  static_assert(sizeof(HF) == sizeof(Fr), "same layout");
  ZTab<Fr> tab;
  HF* t = reinterpret_cast<HF*>(tab.t);
  HF cur = HF::from_be(z_be, Curve::MODBYTES);
  for (int k = 0; k < 32; k++) { t[k] = cur; cur = cur.sqr(); }
  const uint32_t rows = 3 * circ->dev.n + circ->dev.m + 1;
  k_circuit_flatten<Fr><<<(rows + 127) / 128, 128, 0, ctx->stream>>>(circ->dev, tab, (Fr*)d_out);
  k_circuit_wc<Fr><<<1, 512, 0, ctx->stream>>>(circ->dev, tab, (Fr*)d_out);
  ctx->launches += 2;
  return launch_check(ctx, "k_circuit_flatten");
}

}  // namespace bp

using namespace bp;

extern "C" {

int bpgpu_circuit_create(bpgpu_ctx* ctx, size_t n, size_t m, size_t q, const uint32_t* row_start, const uint32_t* ent_q, const uint8_t* ent_coeff_be,
                         bpgpu_circuit** out) {
  if (!ctx || !out || !row_start) return BPGPU_E_ARG;
  *out = nullptr;
  if (n >= (1u << 28) || m >= (1u << 28) || q >= (1u << 30)) return BPGPU_E_ARG;
  const size_t rows = 3 * n + m + 1;
  if (row_start[0] != 0) return BPGPU_E_ARG;
  for (size_t r = 0; r < rows; r++) if (row_start[r + 1] < row_start[r]) return BPGPU_E_ARG;
  const size_t nnz = row_start[rows];
  if (nnz && (!ent_q || !ent_coeff_be)) return BPGPU_E_ARG;
  for (size_t e = 0; e < nnz; e++) if ((ent_q[e] & CSR_QMASK) >= q) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  bpgpu_circuit* c = new (std::nothrow) bpgpu_circuit();
  if (!c) return BPGPU_E_CUDA;
  c->ctx = ctx;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_q = up((rows + 1) * 4), o_c = o_q + up(nnz * 4), total = o_c + up(nnz * 32) + 256;
  if (dev_alloc(ctx, &c->mem, total) != cudaSuccess) { delete c; return BPGPU_E_CUDA; }
  uint8_t* base = (uint8_t*)c->mem;
  int rc = BPGPU_OK;
  if (cudaMemcpyAsync(base, row_start, (rows + 1) * 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = BPGPU_E_CUDA;
  if (!rc && nnz && cudaMemcpyAsync(base + o_q, ent_q, nnz * 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = BPGPU_E_CUDA;
  if (!rc && nnz)
    rc = ctx->curve == BPGPU_BLS12_381 ? scalars_from_host<Bls>(ctx, ent_coeff_be, nnz, 1, base + o_c)
                                       : scalars_from_host<Bn>(ctx, ent_coeff_be, nnz, 1, base + o_c);
  if (!rc && stream_sync(ctx) != cudaSuccess) rc = BPGPU_E_CUDA;
  if (rc) { dev_free(ctx, c->mem); delete c; return rc; }
  c->dev.n = (uint32_t)n; c->dev.m = (uint32_t)m; c->dev.q = (uint32_t)q; c->dev.nnz = (uint32_t)nnz;
  c->dev.row_start = (const uint32_t*)base;
  c->dev.ent_q = (const uint32_t*)(base + o_q);
  c->dev.ent_c = base + o_c;
  *out = c;
  return BPGPU_OK;
}

void bpgpu_circuit_free(bpgpu_circuit* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  dev_free(c->ctx, c->mem);
  delete c;
}

int bpgpu_ctx_circuit_put(bpgpu_ctx* ctx, uint64_t key, bpgpu_circuit* c) {
  if (!ctx || !c || c->ctx != ctx) return BPGPU_E_ARG;
  for (auto& kv : ctx->circuit_cache) if (kv.first == key) return BPGPU_E_ARG;
  ctx->circuit_cache.emplace_back(key, c);
  return BPGPU_OK;
}
const bpgpu_circuit* bpgpu_ctx_circuit_get(const bpgpu_ctx* ctx, uint64_t key) {
  if (!ctx) return nullptr;
  for (auto& kv : ctx->circuit_cache) if (kv.first == key) return kv.second;
  return nullptr;
}

size_t bpgpu_circuit_multipliers(const bpgpu_circuit* c) { return c ? c->dev.n : 0; }
size_t bpgpu_circuit_commitments(const bpgpu_circuit* c) { return c ? c->dev.m : 0; }

int bpgpu_circuit_flatten(bpgpu_ctx* ctx, const bpgpu_circuit* c, const uint8_t* z_be, bpgpu_scalars** out) {
  if (!ctx || !c || !z_be || !out || c->ctx->device != ctx->device) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  int rc = bpgpu_scalars_alloc(ctx, 3 * (size_t)c->dev.n + c->dev.m + 1, out);
  if (rc) return rc;
  rc = ctx->curve == BPGPU_BLS12_381 ? circuit_flatten_t<Bls>(ctx, c, z_be, (*out)->d) : circuit_flatten_t<Bn>(ctx, c, z_be, (*out)->d);
  if (rc) { bpgpu_scalars_free(*out); *out = nullptr; }
  return rc;
}

static int verify_batch_entry(bpgpu_ctx* ctx, const bpgpu_circuit* circuit, bpgpu_points* G, bpgpu_points* H, const uint8_t* g_xy,
                              const uint8_t* h_xy, size_t count, const uint8_t* proofs, size_t proof_stride, const uint8_t* comms_xy,
                              const uint8_t* transcript_state, const uint8_t* challenges_be, const uint8_t* rnd_key, size_t rnd_key_len,
                              int32_t* verdicts, uint8_t* dbg_fixed_be, uint8_t* dbg_var_be) {
  if (!ctx || !circuit || !G || !H || !g_xy || !h_xy || rnd_key_len > 56 || (!rnd_key && rnd_key_len)) return BPGPU_E_ARG;
  if (count && (!proofs || !verdicts || (!comms_xy && circuit->dev.m))) return BPGPU_E_ARG;
  if ((transcript_state == nullptr) == (challenges_be == nullptr)) return BPGPU_E_ARG;      // exactly one source of challenges
  if (circuit->ctx->device != ctx->device || circuit->ctx->curve != ctx->curve) return BPGPU_E_ARG;
  if (count == 0) return BPGPU_OK;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  size_t N = 1, lg = 0;
  while (N < circuit->dev.n) { N <<= 1; lg++; }
  const size_t mb = (size_t)bpgpu_modbytes(ctx->curve);
  const size_t plen = 11 * (2 * mb + 1) + 3 * mb + 2 * lg * (2 * mb + 1) + 2 * mb;
  if (proof_stride < plen) return BPGPU_E_ARG;
  if (lg >= 32) { for (size_t i = 0; i < count; i++) verdicts[i] = BPGPU_E_VERIFY; return BPGPU_OK; }      // ipp.rs:269-273
  if (G->n < N || H->n < N) { for (size_t i = 0; i < count; i++) verdicts[i] = BPGPU_E_GENS_LEN; return BPGPU_OK; }   // verifier.rs:297-299
  int rc;
  if ((rc = bpgpu_points_precompute(ctx, G)) || (rc = bpgpu_points_precompute(ctx, H))) return rc;
  bpgpu_fixed_run runs[4] = {{G, 0, N, nullptr}, {H, 0, N, nullptr}, {nullptr, 0, 1, g_xy}, {nullptr, 0, 1, h_xy}};
  FixedRuns fr;
  uint32_t F = 0;
  if ((rc = resolve_fixed_runs(ctx, runs, 4, &fr, &F))) return rc;
  return ctx->curve == BPGPU_BLS12_381
             ? verify_batch_t<Bls>(ctx, circuit, fr, count, proofs, proof_stride, comms_xy, transcript_state, challenges_be, rnd_key, rnd_key_len,
                                   verdicts, dbg_fixed_be, dbg_var_be)
             : verify_batch_t<Bn>(ctx, circuit, fr, count, proofs, proof_stride, comms_xy, transcript_state, challenges_be, rnd_key, rnd_key_len,
                                  verdicts, dbg_fixed_be, dbg_var_be);
}

int bpgpu_r1cs_verify_batch(bpgpu_ctx* ctx, const bpgpu_circuit* circuit, bpgpu_points* G, bpgpu_points* H, const uint8_t* g_xy, const uint8_t* h_xy,
                            size_t count, const uint8_t* proofs, size_t proof_stride, const uint8_t* comms_xy, const uint8_t* transcript_state,
                            const uint8_t* challenges_be, const uint8_t* rnd_key, size_t rnd_key_len, int32_t* verdicts) {
  return verify_batch_entry(ctx, circuit, G, H, g_xy, h_xy, count, proofs, proof_stride, comms_xy, transcript_state, challenges_be, rnd_key,
                            rnd_key_len, verdicts, nullptr, nullptr);
}

int bpgpu_r1cs_verify_batch_terms(bpgpu_ctx* ctx, const bpgpu_circuit* circuit, bpgpu_points* G, bpgpu_points* H, const uint8_t* g_xy,
                                  const uint8_t* h_xy, size_t count, const uint8_t* proofs, size_t proof_stride, const uint8_t* comms_xy,
                                  const uint8_t* transcript_state, const uint8_t* challenges_be, const uint8_t* rnd_key, size_t rnd_key_len,
                                  int32_t* verdicts, uint8_t* fixed_scalars_be, uint8_t* var_scalars_be) {
  if (!fixed_scalars_be || !var_scalars_be || count > 4096 || verify_slab() < count) return BPGPU_E_ARG;
  return verify_batch_entry(ctx, circuit, G, H, g_xy, h_xy, count, proofs, proof_stride, comms_xy, transcript_state, challenges_be, rnd_key,
                            rnd_key_len, verdicts, fixed_scalars_be, var_scalars_be);
}

}  // extern "C"
