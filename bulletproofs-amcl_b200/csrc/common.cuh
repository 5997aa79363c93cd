// Shared internals of libbpgpu: curve traits, the context object, scratch arenas, vector I/O.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <new>
#include <type_traits>
#include <vector>

#include "../../include/bpgpu.h"
#include "curves.cuh"

namespace bp {

// A scalar as the MSM consumes it: canonical integer, 8 little-endian 32-bit limbs.
struct ScalarInt { uint32_t v[8]; };

#define BP_CUDA_OK(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      fprintf(stderr, "bpgpu: CUDA error %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return BPGPU_E_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

// Growable device scratch buffer owned by a context.
struct Scratch {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return BPGPU_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    BP_CUDA_OK(cudaMalloc(&p, want));
    cap = want;
    return BPGPU_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace bp

struct bpgpu_fixed_bases;

struct bpgpu_ctx {
  int curve = 0;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaMemPool_t pool = nullptr;   // the library's stream-ordered pool on this device (handle storage: G1Vector / FieldElementVector / IPP state)
  // second queue for the host->device copy of the points of a large bpgpu_msm_refs: the copy and the byte->limb
  // conversion overlap the scalar pipeline (digits, scan, scatter) on `stream`; k_chunk_acc waits on `points_ready`
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t points_ready = nullptr;
  bool wait_points = false;        // consumed by the next msm_run
  // waiting for the stream: spin (lowest latency, one core per context) or sleep on a blocking event (BPGPU_BLOCKING_SYNC=1 /
  // bpgpu_ctx_set_blocking_sync: lets more contexts than cores keep launches in flight)
  bool blocking_sync = false;
  cudaEvent_t sync_event = nullptr;
  uint64_t launches = 0;
  struct { bool active = false; int W = 0, c = 0, qshift = 0, hp = 0; size_t tn = 0; } pending;   // an MSM begun, not finished (api.cu)
  bool fixed_schedule = false;    // table sums with a fixed trip count per term (secret scalars): bpgpu_ctx_set_fixed_schedule
  // MSM scratch
  bp::Scratch msm_a, msm_b, msm_c, msm_d, msm_e, io_dev, io_dev2;
  bp::Scratch fr_tmp, fr_out, fr_args, fr_pow, fr_pow2, ipp_pts, ipp_scl, parts_pts, parts_scl, tbl_part;
  bp::Scratch vb;                 // batched verifier / prover: proof bytes, per-proof headers, transcript states
  uint8_t* pinned = nullptr;      // small pinned staging (results, challenges)
  // host-mapped word the decoding kernels raise when a host-supplied point is not a point (coordinate >= p, off the curve);
  // read after the next synchronisation by the entry point that consumed the input (inputs_ok)
  uint32_t* bad_input = nullptr;
  size_t pinned_cap = 0;
  // per-stage CUDA-event timing of the MSM pipeline (bpgpu_ctx_set_profile)
  int profile = 0;
  double stage_ms_sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint64_t stage_runs = 0;
  std::vector<bpgpu_fixed_bases*> fb_cache;   // ctx-owned fixed-base tables (fixedbase.cu)
  bpgpu_ctx* aux = nullptr;       // a second context on the same device, owned by this one (bpgpu_ctx_aux): the other driver of a batch call
  std::vector<std::pair<uint64_t, bpgpu_circuit*>> circuit_cache;   // recorded circuits kept with the ctx (bpgpu_ctx_circuit_put)
};

struct bpgpu_points {
  bpgpu_ctx* ctx;
  void* d;        // Affine<Fq>[n]
  size_t n;
  void* table = nullptr;   // optional window tables Affine<Fq>[n][64][15] (bpgpu_points_precompute, fixedbase.cu)
  void* table16 = nullptr; // optional wide tables Affine<Fq>[n][16][65535] (bpgpu_points_precompute_wide)
};
// a device allocation shared by several bpgpu_scalars handles (bpgpu_scalars_view): released with the last of them
struct bpgpu_shared_block { void* base; int refs; };
struct bpgpu_scalars {
  bpgpu_ctx* ctx;
  void* d;        // Fr[n] Montgomery
  size_t n;
  bpgpu_shared_block* blk = nullptr;   // null: the handle owns d outright
};

namespace bp {

// Handle storage (G1Vector / FieldElementVector / IPP state) comes from the device's stream-ordered pool on the ctx
// stream: allocation and release are queue operations (microseconds), not driver calls that synchronise the device.
// The pool is the library's own, one per device (api.cu library_pool; release threshold raised so freed blocks are reused).
inline cudaError_t stream_sync(bpgpu_ctx* ctx) {
  if (!ctx->blocking_sync) return cudaStreamSynchronize(ctx->stream);
  cudaError_t e = cudaEventRecord(ctx->sync_event, ctx->stream);
  return e != cudaSuccess ? e : cudaEventSynchronize(ctx->sync_event);
}
// BPGPU_E_FORMAT if a point decoded since the last check was invalid (call after a stream synchronisation)
inline int inputs_ok(bpgpu_ctx* ctx) {
  if (ctx->bad_input && *(volatile uint32_t*)ctx->bad_input) { *(volatile uint32_t*)ctx->bad_input = 0; return BPGPU_E_FORMAT; }
  return BPGPU_OK;
}
inline cudaError_t dev_alloc(bpgpu_ctx* ctx, void** p, size_t bytes) { return cudaMallocFromPoolAsync(p, bytes ? bytes : 16, ctx->pool, ctx->stream); }
inline void dev_free(bpgpu_ctx* ctx, void* p) { if (p) cudaFreeAsync(p, ctx->stream); }

// ------------------------------------------------------------------ 128-bit vector I/O
template <class T>
__device__ __forceinline__ T load_vec(const T* src) {
  static_assert(sizeof(T) % 16 == 0, "16-byte multiple");
  T out;
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(&out);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
  return out;
}
template <class T>
__device__ __forceinline__ T load_vec_ro(const T* src) {   // read-only path (LDG.NC)
  T out;
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(&out);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = __ldg(s + i);
  return out;
}
template <class T>
__device__ __forceinline__ void store_vec(T* dst, const T& v) {
  static_assert(sizeof(T) % 16 == 0, "16-byte multiple");
  uint4* d = reinterpret_cast<uint4*>(dst);
  const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
}

// ------------------------------------------------------------------ byte <-> limb conversion
// big-endian MODBYTES -> N little-endian limbs (canonical integer); bytes beyond 4N are ignored
template <int N>
__device__ __forceinline__ void be_to_limbs(const uint8_t* be, int nbytes, uint32_t* out) {
#pragma unroll
  for (int i = 0; i < N; i++) {
    const uint8_t* p = be + nbytes - 4 * (i + 1);
    out[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
  }
}
template <int N>
__device__ __forceinline__ void limbs_to_be(const uint32_t* in, int nbytes, uint8_t* be) {
  for (int i = 0; i < nbytes - 4 * N; i++) be[i] = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint8_t* p = be + nbytes - 4 * (i + 1);
    p[0] = (uint8_t)(in[i] >> 24); p[1] = (uint8_t)(in[i] >> 16); p[2] = (uint8_t)(in[i] >> 8); p[3] = (uint8_t)in[i];
  }
}

// x mod p for an arbitrary N-limb integer x < 2^(32N): 2^(32N) / p < 10 for all four fields (BLS12-381 Fq: 9.85), so ten
// conditional subtractions always reach [0, p)
template <class F>
__device__ __forceinline__ void canonicalise(F& x) { hd_canonicalise(x); }

int launch_check(bpgpu_ctx* ctx, const char* what);
}  // namespace bp
extern "C" int scalars_alloc_many(bpgpu_ctx* ctx, size_t n, int k, bpgpu_scalars** const* outs);   // frvec.cu
namespace bp {

// per-window sums an MSM leaves on the device: result = sum_w 2^(c*w) * winsum[w]
// window w = P_w + 2^qshift * Q_w with P = d_winsum[0..W), Q = d_winsum[W..2W)
// with nsets scalar sets (msm_run d_scalars2) set s owns windows [s*W, (s+1)*W) of the nsets*W in all: P = d_winsum[0, nsets*W),
// Q = d_winsum[nsets*W, 2*nsets*W)
struct MsmResult { int W; int c; int qshift; const void* d_winsum; int nsets = 1; };

// ---- window tables of fixed points (fixedbase.cu): T[i][w][d-1] = d * 2^(8w) * P_i, w < 32, d = 1..255, affine
// 8-bit unsigned windows: a term costs 32 mixed additions (4-bit windows: 64) and a point's table 8160 affine entries
// (765 KB on BLS12-381, 510 KB on BN254) -- HBM is what a B200 has plenty of: 2048 generators (n = 1024) take 1.6 GB
static const int TBL_BITS = 8;
static const int TBL_WINDOWS = 256 / TBL_BITS;                // 32
static const int TBL_DIGITS = (1 << TBL_BITS) - 1;            // 255
static const int TBL_PER_LIMB = 32 / TBL_BITS;                // windows per 32-bit scalar limb
static const int TBL_ENTRIES = TBL_WINDOWS * TBL_DIGITS;      // per point
// WIDE tables (bpgpu_points_precompute_wide): 16-bit windows, T16[i][W][d-1] = d * 2^(16 W) * P_i, 16 x 65535 affine entries per
// point (100 MB on BLS12-381, 67 MB on BN254): HALF the additions per term, for generator sets of a few dozen points that are
// summed millions of times (the 64-multiplier statements of the batch verifier / prover)
static const int TBL16_WINDOWS = 16;
static const int TBL16_DIGITS = 65535;
static const size_t TBL16_ENTRIES = (size_t)TBL16_WINDOWS * TBL16_DIGITS;
static const int TBL_MAX_SEGS = 12;
static const int TBL_MAX_GROUPS = 4;
// one run of terms whose points have tables: table already offset to the first point, scalars = Fr[n];
// group = which of the (up to 4) independent sums of one launch the run belongs to (segments sorted by group)
// rows (optional): term t reads table row rows[t] instead of row t (compacted term lists of the IPP rounds)
struct TableSeg { const void* table; const void* scalars; uint32_t n; int mont; int group = 0; const uint32_t* rows = nullptr; };
template <class Curve> int build_tables(bpgpu_ctx* ctx, const void* d_affine, size_t n, void** table_out);
// per group g: sum over its segments of sum_i s_i * P_i, left as XYZZ points at ctx->tbl_part.p[0 .. ngroups)
// (no doublings, no buckets); one launch pair for all groups
// host_partials (optional): the caller finishes on the host; if the launch leaves per-block sums instead of the group
// totals, *host_partials = blocks per group and group g's sums are at tbl_part.p[TBL_MAX_GROUPS + g * blocks + b]
// (at most TBL_HOST_PARTIALS in all), else *host_partials = 0 and the totals are at tbl_part.p[g]
static const int TBL_HOST_PARTIALS = 160;
template <class Curve> int table_sum_run(bpgpu_ctx* ctx, const TableSeg* segs, int nsegs, int ngroups = 1, int* host_partials = nullptr);
template <class FqParams> void host_sum_partials(uint8_t* xyzz_bytes, int ngroups, int per);
// table-only MSMs, one per group: device sums, one D2H, one shared inversion for the affine results
int msm_tables_to_host(bpgpu_ctx* ctx, const TableSeg* segs, int nsegs, int ngroups, uint8_t* const* outs_xy);
// (fixed-base cache lookup) table of a point that is one of the bases of a cached bpgpu_fixed_bases, or nullptr
template <class Curve> int build_tables16(bpgpu_ctx* ctx, const void* table8, size_t n, void** table16_out);
const void* fixed_table_lookup(bpgpu_ctx* ctx, const uint8_t* xy);
// full MSM over table segments plus (optionally) a general run of device points/scalars; host finish
int msm_mixed_to_host(bpgpu_ctx* ctx, const TableSeg* segs, int nsegs, const void* d_pts, const void* d_scal, bool mont, size_t n,
                      uint8_t* out_xy);

// XYZZ points (device layout, read back to the host) -> affine X||Y big endian with one shared inversion (fixedbase.cu)
void normalise_points_host(int curve, const uint8_t* xyzz_bytes, size_t count, uint8_t* out_xy);
// one counter-mode stream per proof of a batch: out[b][i] = SHAKE256(keys[b] || le64(ctr0[b] + i)) mod r  (rng.cu)
template <class Curve> int fr_random_batch_run(bpgpu_ctx* ctx, const uint8_t* d_keys, uint32_t klen, const uint64_t* d_ctr0, size_t batch, size_t n,
                                               void* d_out);
// host X||Y big-endian points -> device affine Montgomery (api.cu)
template <class Curve> int points_from_host(bpgpu_ctx* ctx, const uint8_t* xy, size_t n, void* dst);
// host big-endian scalars -> device Fr (Montgomery if mont) (api.cu)
template <class Curve> int scalars_from_host(bpgpu_ctx* ctx, const uint8_t* be, size_t n, int mont, void* dst);
// full MSM: device pipeline + host finish; d_scal = Fr[n] (Montgomery if mont) (api.cu)
int msm_to_host(bpgpu_ctx* ctx, const void* d_pts, const void* d_scal, bool mont, size_t n, uint8_t* out_xy);
// host scalars -> device Montgomery argument block in ctx->fr_args (frvec.cu)
template <class Curve> int fr_args_upload(bpgpu_ctx* ctx, const uint8_t* be, int cnt, typename Curve::Fr** d_out);
template <class Curve> int fr_pow_table_upload(bpgpu_ctx* ctx, const uint8_t* x_be, Scratch& dst, typename Curve::Fr** d_out);

template <class Curve> int msm_run(bpgpu_ctx* ctx, const Affine<typename Curve::Fq>* d_points, const void* d_scalars,
                                   bool scalars_mont, size_t n, MsmResult* res, const void* d_scalars2 = nullptr);
// two MSMs over the same device points (scalars Fr[n] each) from ONE pipeline run and one synchronisation (api.cu)
int msm_pair_to_host(bpgpu_ctx* ctx, const void* d_pts, const void* d_scal_a, const void* d_scal_b, bool mont, size_t n, uint8_t* out_a_xy,
                     uint8_t* out_b_xy);

}  // namespace bp
