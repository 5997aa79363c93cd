// Inner-product-argument rounds with device-resident state (IPP::create_ipp,
// /root/reference/src/ipp.rs:35-202) and the verifier's s-vector (ipp.rs:262-315).
//
// The reference folds the generator vectors every round with one two-scalar multiplication per
// element (ipp.rs:119-129,185-187: ~255 doublings each, 85 % of its IPP time).  On a GPU that is a
// 255-deep dependent chain on a shrinking number of threads.  This implementation NEVER
// materialises folded generators.  After k rounds the folded generator is the fixed combination
//     G^(k)[j] = sum_{i = j mod n_k} sG_k(i) * G[i],   sG_k(i) = Gf[i] * prod_{t<k} (bit_t(i) ? u_t : u_t^-1)
// (bit_t(i) = bit lgN-1-t of i; for H the roles of u and u^-1 swap), hence
//     L_k = sum_{i: bit_k(i)=1} a_k[i mod n_k/2] * sG_k(i) * G[i] + sum_{i: bit_k(i)=0} b_k[(i mod n_k/2) + n_k/2] * sH_k(i) * H[i] + c_L * Q
// is ONE MSM over the ORIGINAL bases [G | H | Q] with per-index scalars built by a fused Fr kernel
// (and R_k likewise on the complementary index set).  L_k, R_k are the same group elements the
// reference computes, so the transcript, the challenges and the proof bytes are identical; only
// a, b and the coefficient vectors sG, sH are folded (Fr work), all resident across the lg N
// rounds.  Per round the host sees 2 points (D2H) and sends u, u^-1 (H2D), as in SURVEY.md 3.1.
#include "common.cuh"
#include "host_fp.h"

struct bpgpu_ipp {
  bpgpu_ctx* ctx;
  size_t N, n_cur;
  // table mode (G and H carry window tables and Q = q_scalar * fixed base): no point copies at all
  const void *tG, *tH, *tQ;
  void* wq;                // Fr[1]: q_scalar (Montgomery) -- the Q terms c_L*Q, c_R*Q become (c*q_scalar) * base
  void* rows;              // u32[N]: rows_lo[N/2] | rows_hi[N/2] of the current round (table mode)
  void* P;                 // Affine[2N+1] : G | H | Q   (general mode only)
  void *a, *b;             // Fr[N]
  void *sG, *sH;           // Fr[N]
  void *sclL, *sclR;       // Fr[2N+1]
  // bpgpu_ipp_fold only records the challenge: the fold runs at the head of the next round's kernel (or of finish), so a
  // round is fold + scalar build + cross products in ONE launch, with u, u^-1 as kernel arguments (no H2D copy)
  bool pending;            // a fold by (pu, pui) has been requested but not executed
  size_t n_dev;            // vector length on the device (n_cur once the pending fold has run)
  uint32_t pu[8], pui[8];  // canonical limbs of u, u^-1
};

namespace bp {

// scalars of the L and R MSMs of the current round (length 2N+1 each; slot 2N is filled by k_ipp_cross)
template <class Fr>
__global__ void __launch_bounds__(128) k_ipp_build(uint32_t N, uint32_t n_cur, const Fr* __restrict__ a, const Fr* __restrict__ b,
                                                   const Fr* __restrict__ sG, const Fr* __restrict__ sH,
                                                   Fr* __restrict__ sclL, Fr* __restrict__ sclR) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint32_t half = n_cur >> 1;
  const uint32_t p = i & (n_cur - 1);
  const Fr zero = Fr::zero();
  Fr g = load_vec(sG + i), h = load_vec(sH + i);
  if (p < half) {
    // i is in the left half: H_L takes b_R (L), G_L takes a_R (R)
    store_vec(sclL + i, zero);
    store_vec(sclL + N + i, load_vec(b + p + half) * h);
    store_vec(sclR + i, load_vec(a + p + half) * g);
    store_vec(sclR + N + i, zero);
  } else {
    // right half: G_R takes a_L (L), H_R takes b_L (R)
    store_vec(sclL + i, load_vec(a + p - half) * g);
    store_vec(sclL + N + i, zero);
    store_vec(sclR + i, zero);
    store_vec(sclR + N + i, load_vec(b + p - half) * h);
  }
}

// Table mode: only the N + 1 non-zero terms of each side are listed.  With lo = {i : (i mod n_cur) < n_cur/2} and hi its
// complement (N/2 indices each, written to rows_lo / rows_hi):
//   L = sum_{i in hi} a[p - half] sG[i] * G[i] + sum_{i in lo} b[p + half] sH[i] * H[i]     -> sclL = [G part | H part | c_L]
//   R = sum_{i in lo} a[p + half] sG[i] * G[i] + sum_{i in hi} b[p - half] sH[i] * H[i]     -> sclR = [G part | H part | c_R]
template <class Fr>
__global__ void __launch_bounds__(128) k_ipp_build_compact(uint32_t N, uint32_t n_cur, const Fr* __restrict__ a, const Fr* __restrict__ b,
                                                           const Fr* __restrict__ sG, const Fr* __restrict__ sH, Fr* __restrict__ sclL,
                                                           Fr* __restrict__ sclR, uint32_t* __restrict__ rows_lo,
                                                           uint32_t* __restrict__ rows_hi) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t Nh = N >> 1;
  if (t >= Nh) return;
  const uint32_t half = n_cur >> 1;
  const uint32_t q = t / half, r = t - q * half;
  const uint32_t ilo = q * n_cur + r, ihi = ilo + half;
  rows_lo[t] = ilo;
  rows_hi[t] = ihi;
  const Fr aL = load_vec(a + r), aR = load_vec(a + r + half), bL = load_vec(b + r), bR = load_vec(b + r + half);
  store_vec(sclL + t, aL * load_vec(sG + ihi));            // G_R takes a_L
  store_vec(sclL + Nh + t, bR * load_vec(sH + ilo));       // H_L takes b_R
  store_vec(sclR + t, aR * load_vec(sG + ilo));            // G_L takes a_R
  store_vec(sclR + Nh + t, bL * load_vec(sH + ihi));       // H_R takes b_L
}

// c_L = <a_L, b_R>, c_R = <a_R, b_L>  (ipp.rs:77-78,145-146); one block, results to outL / outR
template <class Fr>
__global__ void __launch_bounds__(256) k_ipp_cross(uint32_t n_cur, const Fr* __restrict__ a, const Fr* __restrict__ b,
                                                   const Fr* __restrict__ wq, Fr* __restrict__ outL, Fr* __restrict__ outR) {
  __shared__ __align__(16) unsigned char smraw[256 * sizeof(Fr)];
  Fr* sm = reinterpret_cast<Fr*>(smraw);
  const uint32_t half = n_cur >> 1;
  Fr cl = Fr::zero(), cr = Fr::zero();
  for (uint32_t j = threadIdx.x; j < half; j += blockDim.x) {
    Fr al = load_vec(a + j), ar = load_vec(a + j + half), bl = load_vec(b + j), br = load_vec(b + j + half);
    cl = cl + al * br;
    cr = cr + ar * bl;
  }
  for (int pass = 0; pass < 2; pass++) {
    store_vec(sm + threadIdx.x, pass == 0 ? cl : cr);
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) store_vec(sm + threadIdx.x, load_vec(sm + threadIdx.x) + load_vec(sm + threadIdx.x + o));
      __syncthreads();
    }
    if (threadIdx.x == 0) { Fr c = load_vec(sm); if (wq) c = c * wq[0]; store_vec(pass == 0 ? outL : outR, c); }
    __syncthreads();
  }
}

// fold a, b by (u, u^-1) and multiply the coefficient vectors (ipp.rs:115-130,181-188); uv = {u, u_inv}
struct FrArg { uint32_t v[8]; };             // a canonical scalar passed by value
template <class Fr>
__device__ __forceinline__ Fr arg_to_mont(const FrArg& x) {
  Fr t;
#pragma unroll
  for (int k = 0; k < 8; k++) t.v[k] = x.v[k];
  return t.to_mont();
}

template <class Fr>
__global__ void __launch_bounds__(128) k_ipp_init(uint32_t N, const Fr* __restrict__ a, const Fr* __restrict__ b, const Fr* __restrict__ Gf,
                                                  const Fr* __restrict__ Hf, FrArg q, Fr* __restrict__ da, Fr* __restrict__ db,
                                                  Fr* __restrict__ dG, Fr* __restrict__ dH, Fr* __restrict__ wq) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) store_vec(wq, arg_to_mont<Fr>(q));
  if (i >= N) return;
  store_vec(da + i, load_vec(a + i));
  store_vec(db + i, load_vec(b + i));
  store_vec(dG + i, load_vec(Gf + i));
  store_vec(dH + i, load_vec(Hf + i));
}

template <class Fr>
__device__ __forceinline__ void fold_element(uint32_t i, uint32_t n_cur, const Fr& u, const Fr& ui, Fr* __restrict__ a, Fr* __restrict__ b,
                                             Fr* __restrict__ sG, Fr* __restrict__ sH, Fr* __restrict__ ab_out) {
  const uint32_t half = n_cur >> 1;
  const uint32_t p = i & (n_cur - 1);
  Fr g = load_vec(sG + i), h = load_vec(sH + i);
  if (p < half) { g = g * ui; h = h * u; } else { g = g * u; h = h * ui; }
  store_vec(sG + i, g);
  store_vec(sH + i, h);
  if (i < half) {
    Fr al = load_vec(a + i), ar = load_vec(a + i + half), bl = load_vec(b + i), br = load_vec(b + i + half);
    const Fr na = al * u + ui * ar, nb = bl * ui + u * br;
    store_vec(a + i, na);
    store_vec(b + i, nb);
    if (half == 1 && ab_out) { store_vec(ab_out, na); store_vec(ab_out + 1, nb); }   // the proof's a, b side by side
  }
}

// fold a, b by (u, u^-1) and multiply the coefficient vectors (ipp.rs:115-130,181-188)
template <class Fr>
__global__ void __launch_bounds__(128) k_ipp_fold(uint32_t N, uint32_t n_cur, FrArg uc, FrArg uic, Fr* __restrict__ a, Fr* __restrict__ b,
                                                  Fr* __restrict__ sG, Fr* __restrict__ sH, Fr* __restrict__ ab_out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const Fr u = arg_to_mont<Fr>(uc), ui = arg_to_mont<Fr>(uic);
  fold_element(i, n_cur, u, ui, a, b, sG, sH, ab_out);
}

// One launch per round for N <= 4096 in table mode: [pending fold] -> compacted L / R scalar lists -> cross products.
// A single block, so __syncthreads orders the fold's writes before the build's reads.
template <class Fr>
__global__ void __launch_bounds__(512) k_ipp_round_fused(uint32_t N, uint32_t n_in, int do_fold, FrArg uc, FrArg uic, Fr* __restrict__ a,
                                                         Fr* __restrict__ b, Fr* __restrict__ sG, Fr* __restrict__ sH,
                                                         const Fr* __restrict__ wq, Fr* __restrict__ sclL, Fr* __restrict__ sclR,
                                                         uint32_t* __restrict__ rows_lo, uint32_t* __restrict__ rows_hi) {
  __shared__ __align__(16) unsigned char smraw[512 * sizeof(Fr)];
  Fr* sm = reinterpret_cast<Fr*>(smraw);
  uint32_t n_cur = n_in;
  if (do_fold) {
    const Fr u = arg_to_mont<Fr>(uc), ui = arg_to_mont<Fr>(uic);
    for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) fold_element(i, n_cur, u, ui, a, b, sG, sH, (Fr*)nullptr);
    n_cur >>= 1;
    __syncthreads();
  }
  const uint32_t Nh = N >> 1, half = n_cur >> 1;
  for (uint32_t t = threadIdx.x; t < Nh; t += blockDim.x) {             // k_ipp_build_compact
    const uint32_t q = t / half, r = t - q * half;
    const uint32_t ilo = q * n_cur + r, ihi = ilo + half;
    rows_lo[t] = ilo;
    rows_hi[t] = ihi;
    const Fr aL = load_vec(a + r), aR = load_vec(a + r + half), bL = load_vec(b + r), bR = load_vec(b + r + half);
    store_vec(sclL + t, aL * load_vec(sG + ihi));
    store_vec(sclL + Nh + t, bR * load_vec(sH + ilo));
    store_vec(sclR + t, aR * load_vec(sG + ilo));
    store_vec(sclR + Nh + t, bL * load_vec(sH + ihi));
  }
  Fr cl = Fr::zero(), cr = Fr::zero();                                  // k_ipp_cross
  for (uint32_t j = threadIdx.x; j < half; j += blockDim.x) {
    Fr al = load_vec(a + j), ar = load_vec(a + j + half), bl = load_vec(b + j), br = load_vec(b + j + half);
    cl = cl + al * br;
    cr = cr + ar * bl;
  }
  for (int pass = 0; pass < 2; pass++) {
    store_vec(sm + threadIdx.x, pass == 0 ? cl : cr);
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) store_vec(sm + threadIdx.x, load_vec(sm + threadIdx.x) + load_vec(sm + threadIdx.x + o));
      __syncthreads();
    }
    if (threadIdx.x == 0) { Fr c = load_vec(sm) * wq[0]; store_vec((pass == 0 ? sclL : sclR) + N, c); }
    __syncthreads();
  }
}

// s[i] = prod_t (bit_{lg-1-t}(i) ? u_t : u_t^-1)   (closed form of ipp.rs:303-312); uv = u[0..lg) | u_inv[0..lg)
template <class Fr>
__global__ void __launch_bounds__(128) k_ipp_s(uint32_t N, int lg, const Fr* __restrict__ uv, Fr* __restrict__ s) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  Fr acc = Fr::one();
  for (int t = 0; t < lg; t++) acc = acc * (((i >> (lg - 1 - t)) & 1) ? uv[t] : uv[lg + t]);
  store_vec(s + i, acc);
}

// scalars of verify_ipp's MSM (ipp.rs:220-242) for points [Q | G | H | L | R]:
//   [a*b | (a*s_i)*Gf_i | (b*s_{n-1-i})*Hf_i | -u_k^2 | -u_k^-2 ] ; args = {a, b} ; uv as in k_ipp_s
template <class Fr>
__global__ void __launch_bounds__(128) k_ipp_verify_scalars(uint32_t N, int lg, const Fr* __restrict__ ab, const Fr* __restrict__ uv,
                                                            const Fr* __restrict__ s, const Fr* __restrict__ Gf,
                                                            const Fr* __restrict__ Hf, Fr* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    store_vec(out + 1 + i, (ab[0] * load_vec(s + i)) * load_vec(Gf + i));
    store_vec(out + 1 + N + i, (ab[1] * load_vec(s + (N - 1 - i))) * load_vec(Hf + i));
  }
  if (i == 0) store_vec(out, ab[0] * ab[1]);
  if (i < (uint32_t)lg) {
    store_vec(out + 1 + 2 * N + i, uv[i].sqr().neg());
    store_vec(out + 1 + 2 * N + lg + i, uv[lg + i].sqr().neg());
  }
}

// ---- hybrid for large N in table mode: the folded generators ARE materialised, once ----
// While G and H are the original generators every round costs N + 1 table terms per side whatever n_cur is.  After k rounds
//     G'[j] = sum_{t < 2^k} sG[t * n_cur + j] * G[t * n_cur + j]      (H' likewise, Q = q_scalar * base)
// is 2 n_cur + 1 table sums of 2^k terms (32 N additions in all: one more round), and from then on L and R are MSMs over
// the 2 n_cur + 1 materialised points (bucket pipeline, coefficient vectors = 1).  Same group elements, same proof bytes.
template <class Fr>
__global__ void __launch_bounds__(128) k_ipp_mat_scalars(uint32_t N, uint32_t n_cur, Fr* __restrict__ sG, Fr* __restrict__ sH,
                                                         const Fr* __restrict__ wq, Fr* __restrict__ canon) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) store_vec(canon + 2 * (size_t)N, load_vec(wq).from_mont());
  if (i >= N) return;
  store_vec(canon + i, load_vec(sG + i).from_mont());
  store_vec(canon + N + i, load_vec(sH + i).from_mont());
  if (i < n_cur) { store_vec(sG + i, Fr::one()); store_vec(sH + i, Fr::one()); }
}

// MAT_LANES lanes per materialised point: lane q adds the entries of byte-windows q, q + MAT_LANES, ... of the point's 2^k
// terms (a chain of 32 / MAT_LANES * 2^k mixed additions), then log2(MAT_LANES) tree levels.  Few lanes per point keep the
// tree (dependent FULL additions on half, a quarter, ... of the lanes) a small share of the work.
static const int MAT_LANES = 8;
template <class Curve>
__global__ void __launch_bounds__(128) k_ipp_materialise(uint32_t N, uint32_t n_cur, const void* __restrict__ tG, const void* __restrict__ tH,
                                                         const void* __restrict__ tQ, const typename Curve::Fr* __restrict__ canon,
                                                         XYZZ<typename Curve::Fq>* __restrict__ out) {
  using Fq = typename Curve::Fq;
  __shared__ __align__(16) unsigned char smraw[128 * sizeof(XYZZ<Fq>)];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(smraw);
  const uint32_t rows = 2 * n_cur + 1;
  const uint32_t r = blockIdx.x * (128 / MAT_LANES) + threadIdx.x / MAT_LANES, q = threadIdx.x % MAT_LANES;
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  if (r < rows) {
    const bool isq = r == 2 * n_cur, ish = r >= n_cur;
    const Affine<Fq>* tb = (const Affine<Fq>*)(isq ? tQ : ish ? tH : tG);
    const uint32_t j = isq ? 0 : ish ? r - n_cur : r;
    const typename Curve::Fr* sc = canon + (isq ? 2 * (size_t)N : ish ? (size_t)N : 0);
    const uint32_t step = isq ? 1 : n_cur, terms = isq ? 1 : N / n_cur;
#pragma unroll 1
    for (uint32_t t = 0; t < terms; t++) {
      const uint32_t i = t * step + j;
#pragma unroll 1
      for (uint32_t w = q; w < (uint32_t)TBL_WINDOWS; w += MAT_LANES) {
        const uint32_t d = (sc[i].v[w / TBL_PER_LIMB] >> ((w % TBL_PER_LIMB) * TBL_BITS)) & (uint32_t)TBL_DIGITS;
        if (d) acc.madd(load_vec_ro(tb + ((size_t)i * TBL_WINDOWS + w) * TBL_DIGITS + (d - 1)));
      }
    }
  }
  store_vec(sm + threadIdx.x, acc);
  __syncthreads();
  for (int o = MAT_LANES / 2; o > 0; o >>= 1) {
    if ((int)q < o) {
      XYZZ<Fq> a = load_vec(sm + threadIdx.x), c = load_vec(sm + threadIdx.x + o);
      a.add(c);
      store_vec(sm + threadIdx.x, a);
    }
    __syncthreads();
  }
  if (q == 0 && r < rows) store_vec(out + r, load_vec(sm + threadIdx.x));
}

// XYZZ -> affine (Montgomery form, the identity as (0, 0)); MAT_NORM points per thread share one inversion
static const int MAT_NORM = 4;
template <class Fq>
__global__ void __launch_bounds__(64) k_ipp_mat_affine(uint32_t rows, const XYZZ<Fq>* __restrict__ src, Affine<Fq>* __restrict__ dst) {
  const uint32_t r0 = (blockIdx.x * blockDim.x + threadIdx.x) * MAT_NORM;
  if (r0 >= rows) return;
  const uint32_t cnt = min((uint32_t)MAT_NORM, rows - r0);
  Fq pre[MAT_NORM + 1];
  pre[0] = Fq::one();
  for (uint32_t i = 0; i < cnt; i++) {
    const Fq zzz = load_vec(&src[r0 + i].zzz);
    pre[i + 1] = zzz.is_zero() ? pre[i] : Fq::mulc(pre[i], zzz);
  }
  Fq inv = pre[cnt].inv();
  for (uint32_t i = cnt; i-- > 0;) {
    const XYZZ<Fq> p = load_vec(src + r0 + i);
    if (p.is_inf()) { store_vec(dst + r0 + i, Affine<Fq>::inf()); continue; }
    const Fq i3 = Fq::mulc(inv, pre[i]);                // 1 / zzz
    inv = Fq::mulc(inv, p.zzz);
    const Fq i1 = Fq::mulc(i3, p.zz);                   // 1 / z
    Affine<Fq> a;
    a.x = Fq::mulc(p.x, Fq::mulc(i1, i1));
    a.y = Fq::mulc(p.y, i3);
    store_vec(dst + r0 + i, a);
  }
}

// after how many table rounds the generators are materialised (0 = never); BPGPU_IPP_HYBRID / BPGPU_IPP_HYBRID_MIN override
static int hybrid_rounds(size_t N) {
  const char *ek = getenv("BPGPU_IPP_HYBRID"), *em = getenv("BPGPU_IPP_HYBRID_MIN");
  const int k = ek ? atoi(ek) : 4;
  const size_t nmin = em ? (size_t)atoll(em) : (size_t)8192;
  return N >= nmin && k > 0 && ((size_t)2 << k) <= N ? k : 0;      // at least one round is left after materialising
}

template <class Curve>
static int ipp_materialise_t(bpgpu_ipp* st) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  bpgpu_ctx* ctx = st->ctx;
  const uint32_t N = (uint32_t)st->N, n = (uint32_t)st->n_cur, rows = 2 * n + 1;
  cudaStream_t s = ctx->stream;
  const size_t pbytes = ((size_t)rows * sizeof(Affine<Fq>) + 255) & ~(size_t)255;
  BP_CUDA_OK(dev_alloc(ctx, &st->P, pbytes + (size_t)rows * sizeof(XYZZ<Fq>)));
  XYZZ<Fq>* sums = (XYZZ<Fq>*)((uint8_t*)st->P + pbytes);
  Fr* canon = (Fr*)st->sclL;                            // 2N + 1 slots, free between rounds
  k_ipp_mat_scalars<Fr><<<(N + 127) / 128, 128, 0, s>>>(N, n, (Fr*)st->sG, (Fr*)st->sH, (const Fr*)st->wq, canon);
  k_ipp_materialise<Curve><<<(rows * MAT_LANES + 127) / 128, 128, 0, s>>>(N, n, st->tG, st->tH, st->tQ, canon, sums);
  k_ipp_mat_affine<Fq><<<((rows + MAT_NORM - 1) / MAT_NORM + 63) / 64, 64, 0, s>>>(rows, sums, (Affine<Fq>*)st->P);
  ctx->launches += 3;
  st->N = n;                                            // general mode over [G' | H' | Q] from here on
  st->wq = nullptr;
  st->tG = st->tH = st->tQ = nullptr;
  return launch_check(ctx, "ipp_materialise");
}

template <class Curve>
static int ipp_begin_t(bpgpu_ipp* st, const bpgpu_points* Gp, size_t goff, const bpgpu_points* Hp, size_t hoff, const uint8_t* Q_xy,
                       const uint8_t* q_base_xy, const uint8_t* q_scalar_be, const void* Gf, const void* Hf, const void* a, const void* b) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  bpgpu_ctx* ctx = st->ctx;
  const size_t N = st->N;
  cudaStream_t s = ctx->stream;
  int rc;
  const bool tables = Gp->table && Hp->table && q_base_xy && q_scalar_be;
  void* frs = nullptr;
  const size_t fr_count = 4 * N + 2 * (2 * N + 1) + 1;
  BP_CUDA_OK(dev_alloc(ctx, &frs, fr_count * sizeof(Fr) + (tables ? (N + 1) * sizeof(uint32_t) : 0)));   // + the row lists (table mode)
  st->a = frs;
  st->b = (Fr*)frs + N;
  st->sG = (Fr*)frs + 2 * N;
  st->sH = (Fr*)frs + 3 * N;
  st->sclL = (Fr*)frs + 4 * N;
  st->sclR = (Fr*)frs + 4 * N + (2 * N + 1);
  st->wq = nullptr;
  if (tables) {
    bpgpu_fixed_bases* fb = nullptr;
    if ((rc = bpgpu_fixed_bases_get(ctx, q_base_xy, 1, &fb))) return rc;
    st->tQ = fixed_table_lookup(ctx, q_base_xy);
    st->tG = (const uint8_t*)Gp->table + goff * TBL_ENTRIES * sizeof(Affine<Fq>);
    st->tH = (const uint8_t*)Hp->table + hoff * TBL_ENTRIES * sizeof(Affine<Fq>);
    st->wq = (Fr*)frs + 4 * N + 2 * (2 * N + 1);
    st->rows = (Fr*)frs + fr_count;                       // same allocation: released with it
    // one launch: the four vector clones of ipp.rs:57-60 and q_scalar (a kernel argument) in Montgomery form
    FrArg qa;
    for (int k = 0; k < 8; k++) {
      const uint8_t* p = q_scalar_be + Curve::MODBYTES - 4 * (k + 1);
      qa.v[k] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    }
    k_ipp_init<Fr><<<(unsigned)((N + 127) / 128), 128, 0, s>>>((uint32_t)N, (const Fr*)a, (const Fr*)b, (const Fr*)Gf, (const Fr*)Hf, qa, (Fr*)st->a,
                                                               (Fr*)st->b, (Fr*)st->sG, (Fr*)st->sH, (Fr*)st->wq);
    ctx->launches++;
    return launch_check(ctx, "k_ipp_init");
  } else {
    const void* G = (const Affine<Fq>*)Gp->d + goff;
    const void* H = (const Affine<Fq>*)Hp->d + hoff;
    BP_CUDA_OK(dev_alloc(ctx, &st->P, (2 * N + 1) * sizeof(Affine<Fq>)));
    BP_CUDA_OK(cudaMemcpyAsync(st->P, G, N * sizeof(Affine<Fq>), cudaMemcpyDeviceToDevice, s));
    BP_CUDA_OK(cudaMemcpyAsync((Affine<Fq>*)st->P + N, H, N * sizeof(Affine<Fq>), cudaMemcpyDeviceToDevice, s));
    uint8_t qtmp[2 * 48];
    if (!Q_xy) {                                 // Q given only as q_scalar * q_base: evaluate it once
      bpgpu_fixed_bases* fb = nullptr;
      if ((rc = bpgpu_fixed_bases_get(ctx, q_base_xy, 1, &fb))) return rc;
      if ((rc = bpgpu_fixed_bases_commit(ctx, fb, q_scalar_be, 1, qtmp))) return rc;
      Q_xy = qtmp;
    }
    if ((rc = points_from_host<Curve>(ctx, Q_xy, 1, (Affine<Fq>*)st->P + 2 * N))) return rc;
    BP_CUDA_OK(cudaStreamSynchronize(s));        // qtmp is a stack buffer
    if ((rc = inputs_ok(ctx))) return rc;
  }
  BP_CUDA_OK(cudaMemcpyAsync(st->a, a, N * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
  BP_CUDA_OK(cudaMemcpyAsync(st->b, b, N * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
  BP_CUDA_OK(cudaMemcpyAsync(st->sG, Gf, N * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
  BP_CUDA_OK(cudaMemcpyAsync(st->sH, Hf, N * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
  return BPGPU_OK;
}

template <class Curve>
static int ipp_round_t(bpgpu_ipp* st, uint8_t* L_xy, uint8_t* R_xy) {
  using Fr = typename Curve::Fr;
  bpgpu_ctx* ctx = st->ctx;
  uint32_t N = (uint32_t)st->N;
  const uint32_t n = (uint32_t)st->n_cur;
  int rc;
  FrArg uc, uic;
  memcpy(uc.v, st->pu, sizeof uc.v);
  memcpy(uic.v, st->pui, sizeof uic.v);
  const int hk = st->wq ? hybrid_rounds(N) : 0;
  if (hk > 0 && n <= (N >> hk)) {                       // table rounds are over: materialise, then general mode (below)
    if (st->pending) {
      k_ipp_fold<Fr><<<(N + 127) / 128, 128, 0, ctx->stream>>>(N, (uint32_t)st->n_dev, uc, uic, (Fr*)st->a, (Fr*)st->b, (Fr*)st->sG, (Fr*)st->sH,
                                                               (Fr*)nullptr);
      ctx->launches++;
      st->pending = false;
      st->n_dev = n;
    }
    if ((rc = ipp_materialise_t<Curve>(st))) return rc;
    N = (uint32_t)st->N;
  }
  if (st->wq && N <= 4096) {
    // small table mode: [fold] + scalar lists + cross products in one launch, both sums in one more (two groups)
    const uint32_t Nh = N >> 1;
    uint32_t* rows_lo = (uint32_t*)st->rows;
    uint32_t* rows_hi = rows_lo + Nh;
    Fr *sl = (Fr*)st->sclL, *sr = (Fr*)st->sclR;
    k_ipp_round_fused<Fr><<<1, 512, 0, ctx->stream>>>(N, (uint32_t)st->n_dev, st->pending ? 1 : 0, uc, uic, (Fr*)st->a, (Fr*)st->b, (Fr*)st->sG,
                                                      (Fr*)st->sH, (const Fr*)st->wq, sl, sr, rows_lo, rows_hi);
    ctx->launches += 1;
    st->pending = false;
    st->n_dev = n;
    if ((rc = launch_check(ctx, "ipp_round"))) return rc;
    TableSeg segs[6] = {{st->tG, sl, Nh, 1, 0, rows_hi}, {st->tH, sl + Nh, Nh, 1, 0, rows_lo}, {st->tQ, sl + N, 1, 1, 0, nullptr},
                        {st->tG, sr, Nh, 1, 1, rows_lo}, {st->tH, sr + Nh, Nh, 1, 1, rows_hi}, {st->tQ, sr + N, 1, 1, 1, nullptr}};
    uint8_t* outs[2] = {L_xy, R_xy};
    return msm_tables_to_host(ctx, segs, 6, 2, outs);
  }
  if (st->pending) {
    k_ipp_fold<Fr><<<(N + 127) / 128, 128, 0, ctx->stream>>>(N, (uint32_t)st->n_dev, uc, uic, (Fr*)st->a, (Fr*)st->b, (Fr*)st->sG, (Fr*)st->sH,
                                                             (Fr*)nullptr);
    ctx->launches++;
    st->pending = false;
    st->n_dev = n;
  }
  if (st->wq) {
    // table mode: both sums in ONE launch (two groups) over the compacted term lists, one D2H, one shared inversion
    const uint32_t Nh = N >> 1;
    uint32_t* rows_lo = (uint32_t*)st->rows;
    uint32_t* rows_hi = rows_lo + Nh;
    Fr *sl = (Fr*)st->sclL, *sr = (Fr*)st->sclR;
    k_ipp_build_compact<Fr><<<(Nh + 127) / 128, 128, 0, ctx->stream>>>(N, n, (const Fr*)st->a, (const Fr*)st->b, (const Fr*)st->sG,
                                                                     (const Fr*)st->sH, sl, sr, rows_lo, rows_hi);
    k_ipp_cross<Fr><<<1, 256, 0, ctx->stream>>>(n, (const Fr*)st->a, (const Fr*)st->b, (const Fr*)st->wq, sl + N, sr + N);
    ctx->launches += 2;
    if ((rc = launch_check(ctx, "ipp_round"))) return rc;
    TableSeg segs[6] = {{st->tG, sl, Nh, 1, 0, rows_hi}, {st->tH, sl + Nh, Nh, 1, 0, rows_lo}, {st->tQ, sl + N, 1, 1, 0, nullptr},
                        {st->tG, sr, Nh, 1, 1, rows_lo}, {st->tH, sr + Nh, Nh, 1, 1, rows_hi}, {st->tQ, sr + N, 1, 1, 1, nullptr}};
    uint8_t* outs[2] = {L_xy, R_xy};
    return msm_tables_to_host(ctx, segs, 6, 2, outs);
  }
  k_ipp_build<Fr><<<(N + 127) / 128, 128, 0, ctx->stream>>>(N, n, (const Fr*)st->a, (const Fr*)st->b, (const Fr*)st->sG,
                                                           (const Fr*)st->sH, (Fr*)st->sclL, (Fr*)st->sclR);
  k_ipp_cross<Fr><<<1, 256, 0, ctx->stream>>>(n, (const Fr*)st->a, (const Fr*)st->b, (const Fr*)st->wq, (Fr*)st->sclL + 2 * N,
                                              (Fr*)st->sclR + 2 * N);
  ctx->launches += 2;
  rc = launch_check(ctx, "ipp_round");
  if (rc) return rc;
  // L and R share the points: one pipeline run with the R scalars as extra windows, one synchronisation
  return msm_pair_to_host(ctx, st->P, st->sclL, st->sclR, true, 2 * (size_t)N + 1, L_xy, R_xy);
}

// records the challenge; the fold itself runs at the head of the next round's launch (ipp_round_t) or of finish
template <class Curve>
static int ipp_fold_t(bpgpu_ipp* st, const uint8_t* u_be, const uint8_t* ui_be) {
  const int mb = Curve::MODBYTES;
  for (int k = 0; k < 8; k++) {
    const uint8_t *p = u_be + mb - 4 * (k + 1), *q = ui_be + mb - 4 * (k + 1);
    st->pu[k] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    st->pui[k] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
  }
  st->pending = true;
  st->n_cur >>= 1;
  return BPGPU_OK;
}

template <class Curve>
static int ipp_finish_t(bpgpu_ipp* st, uint8_t* a_be, uint8_t* b_be) {
  using Fr = typename Curve::Fr;
  bpgpu_ctx* ctx = st->ctx;
  const uint32_t N = (uint32_t)st->N;
  Fr* ab = (Fr*)st->sclL;                              // a, b side by side for ONE download
  if (st->pending) {
    FrArg uc, uic;
    memcpy(uc.v, st->pu, sizeof uc.v);
    memcpy(uic.v, st->pui, sizeof uic.v);
    k_ipp_fold<Fr><<<(N + 127) / 128, 128, 0, ctx->stream>>>(N, (uint32_t)st->n_dev, uc, uic, (Fr*)st->a, (Fr*)st->b, (Fr*)st->sG, (Fr*)st->sH,
                                                             st->n_dev == 2 ? ab : (Fr*)nullptr);
    ctx->launches++;
    st->pending = false;
    st->n_dev = st->n_cur;
    int rc = launch_check(ctx, "k_ipp_fold");
    if (rc) return rc;
  }
  if (N == 1) {                                        // no round ever ran: a, b are the inputs
    BP_CUDA_OK(cudaMemcpyAsync(ab, st->a, sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    BP_CUDA_OK(cudaMemcpyAsync(ab + 1, st->b, sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  uint8_t both[2 * 48];
  bpgpu_scalars v{ctx, ab, 2};
  int rc = bpgpu_scalars_download(ctx, &v, 0, 2, both);
  if (rc) return rc;
  memcpy(a_be, both, Curve::MODBYTES);
  memcpy(b_be, both + Curve::MODBYTES, Curve::MODBYTES);
  return BPGPU_OK;
}

// u[0..lg) | u_inv[0..lg) as device Montgomery values in ctx->fr_args
template <class Curve>
static int upload_challenges(bpgpu_ctx* ctx, const uint8_t* u_be, size_t lg, typename Curve::Fr** uv) {
  using HF = host::HFp<typename std::conditional<Curve::ID == BPGPU_BLS12_381, BlsFr, BnFr>::type>;
  std::vector<uint8_t> both(2 * lg * Curve::MODBYTES + 1);
  for (size_t k = 0; k < lg; k++) {
    memcpy(both.data() + k * Curve::MODBYTES, u_be + k * Curve::MODBYTES, Curve::MODBYTES);
    HF::from_be(u_be + k * Curve::MODBYTES, Curve::MODBYTES).inv().to_be(both.data() + (lg + k) * Curve::MODBYTES, Curve::MODBYTES);
  }
  return fr_args_upload<Curve>(ctx, both.data(), (int)(2 * lg), uv);
}

template <class Curve>
static int ipp_s_t(bpgpu_ctx* ctx, const uint8_t* u_be, size_t lg, void* d_s) {
  using Fr = typename Curve::Fr;
  Fr* uv;
  int rc = upload_challenges<Curve>(ctx, u_be, lg, &uv);
  if (rc) return rc;
  const uint32_t N = 1u << lg;
  k_ipp_s<Fr><<<(N + 127) / 128, 128, 0, ctx->stream>>>(N, (int)lg, uv, (Fr*)d_s);
  ctx->launches++;
  return launch_check(ctx, "k_ipp_s");
}

// expected_P of verify_ipp.  Always ONE general Pippenger run over [Q | G | H | L | R]: the proof-specific points need the
// general path anyway, and gathering G and H next to them is cheaper than an extra table launch pair (measured at n = 64:
// 1.05 ms vs 1.21 ms; at n = 2^14 the table entries per term still cost more than the bucket method's work).
template <class Curve>
static int ipp_verify_t(bpgpu_ctx* ctx, const void* G, const void* H, const uint8_t* Q_xy, const void* Gf, const void* Hf,
                        const uint8_t* a_be, const uint8_t* b_be, const uint8_t* u_be, const uint8_t* L_xy, const uint8_t* R_xy,
                        size_t lg, uint8_t* out_xy) {
  using Fq = typename Curve::Fq;
  using Fr = typename Curve::Fr;
  const size_t N = (size_t)1 << lg, total = 1 + 2 * N + 2 * lg;
  int rc;
  if ((rc = ctx->ipp_pts.reserve(total * sizeof(Affine<Fq>)))) return rc;
  if ((rc = ctx->ipp_scl.reserve((total + N) * sizeof(Fr)))) return rc;
  Affine<Fq>* P = (Affine<Fq>*)ctx->ipp_pts.p;
  Fr* scl = (Fr*)ctx->ipp_scl.p;
  Fr* s = scl + total;
  if ((rc = points_from_host<Curve>(ctx, Q_xy, 1, P))) return rc;
  BP_CUDA_OK(cudaMemcpyAsync(P + 1, G, N * sizeof(Affine<Fq>), cudaMemcpyDeviceToDevice, ctx->stream));
  BP_CUDA_OK(cudaMemcpyAsync(P + 1 + N, H, N * sizeof(Affine<Fq>), cudaMemcpyDeviceToDevice, ctx->stream));
  if (lg) {
    if ((rc = points_from_host<Curve>(ctx, L_xy, lg, P + 1 + 2 * N))) return rc;
    if ((rc = points_from_host<Curve>(ctx, R_xy, lg, P + 1 + 2 * N + lg))) return rc;
  }
  // argument block: u | u_inv | a | b
  using HF = host::HFp<typename std::conditional<Curve::ID == BPGPU_BLS12_381, BlsFr, BnFr>::type>;
  const int mb = Curve::MODBYTES;
  std::vector<uint8_t> args((2 * lg + 2) * mb);
  for (size_t k = 0; k < lg; k++) {
    memcpy(args.data() + k * mb, u_be + k * mb, mb);
    HF::from_be(u_be + k * mb, mb).inv().to_be(args.data() + (lg + k) * mb, mb);
  }
  memcpy(args.data() + 2 * lg * mb, a_be, mb);
  memcpy(args.data() + (2 * lg + 1) * mb, b_be, mb);
  Fr* dargs;
  if ((rc = fr_args_upload<Curve>(ctx, args.data(), (int)(2 * lg + 2), &dargs))) return rc;
  k_ipp_s<Fr><<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>((uint32_t)N, (int)lg, dargs, s);
  k_ipp_verify_scalars<Fr><<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>((uint32_t)N, (int)lg, dargs + 2 * lg, dargs, s,
                                                                            (const Fr*)Gf, (const Fr*)Hf, scl);
  ctx->launches += 2;
  if ((rc = launch_check(ctx, "ipp_verify"))) return rc;
  return msm_to_host(ctx, P, scl, true, total, out_xy);
}

}  // namespace bp

using namespace bp;

static bool prange(const bpgpu_points* p, size_t off, size_t n) { return p && off <= p->n && n <= p->n - off; }
static bool srange(const bpgpu_scalars* s, size_t off, size_t n) { return s && off <= s->n && n <= s->n - off; }
static size_t psize(const bpgpu_ctx* c) { return c->curve == BPGPU_BLS12_381 ? sizeof(Affine<Bls::Fq>) : sizeof(Affine<Bn::Fq>); }

extern "C" {

static int ipp_begin_common(bpgpu_ctx* ctx, const bpgpu_points* G, size_t goff, const bpgpu_points* H, size_t hoff, const uint8_t* Q_xy,
                            const uint8_t* q_base_xy, const uint8_t* q_scalar_be, const bpgpu_scalars* Gf, const bpgpu_scalars* Hf,
                            const bpgpu_scalars* a, const bpgpu_scalars* b, size_t n, bpgpu_ipp** out) {
  if (!ctx || !G || !H || !Gf || !Hf || !a || !b || !out) return BPGPU_E_ARG;
  if (!Q_xy && !(q_base_xy && q_scalar_be)) return BPGPU_E_ARG;
  *out = nullptr;
  if (n == 0 || (n & (n - 1))) return BPGPU_E_NOT_POW2;                           // ipp.rs:48
  if (!prange(G, goff, n) || !prange(H, hoff, n)) return BPGPU_E_LEN;             // ipp.rs:51
  if (Gf->n != n || Hf->n != n || a->n != n || b->n != n) return BPGPU_E_LEN;     // ipp.rs:52-55
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  bpgpu_ipp* st = new (std::nothrow) bpgpu_ipp();
  if (!st) return BPGPU_E_CUDA;
  memset(st, 0, sizeof *st);
  st->ctx = ctx; st->N = n; st->n_cur = n; st->n_dev = n; st->pending = false;
  int rc = ctx->curve == BPGPU_BLS12_381 ? ipp_begin_t<Bls>(st, G, goff, H, hoff, Q_xy, q_base_xy, q_scalar_be, Gf->d, Hf->d, a->d, b->d)
                                        : ipp_begin_t<Bn>(st, G, goff, H, hoff, Q_xy, q_base_xy, q_scalar_be, Gf->d, Hf->d, a->d, b->d);
  if (rc) { bpgpu_ipp_free(st); return rc; }
  *out = st;
  return BPGPU_OK;
}

int bpgpu_ipp_begin(bpgpu_ctx* ctx, const bpgpu_points* G, size_t goff, const bpgpu_points* H, size_t hoff, const uint8_t* Q_xy,
                    const bpgpu_scalars* Gf, const bpgpu_scalars* Hf, const bpgpu_scalars* a, const bpgpu_scalars* b, size_t n,
                    bpgpu_ipp** out) {
  if (!Q_xy) return BPGPU_E_ARG;
  return ipp_begin_common(ctx, G, goff, H, hoff, Q_xy, nullptr, nullptr, Gf, Hf, a, b, n, out);
}

int bpgpu_ipp_begin_fixed_q(bpgpu_ctx* ctx, const bpgpu_points* G, size_t goff, const bpgpu_points* H, size_t hoff,
                            const uint8_t* q_base_xy, const uint8_t* q_scalar_be, const bpgpu_scalars* Gf, const bpgpu_scalars* Hf,
                            const bpgpu_scalars* a, const bpgpu_scalars* b, size_t n, bpgpu_ipp** out) {
  if (!q_base_xy || !q_scalar_be) return BPGPU_E_ARG;
  return ipp_begin_common(ctx, G, goff, H, hoff, nullptr, q_base_xy, q_scalar_be, Gf, Hf, a, b, n, out);
}

size_t bpgpu_ipp_len(const bpgpu_ipp* st) { return st ? st->n_cur : 0; }

int bpgpu_ipp_round_LR(bpgpu_ipp* st, uint8_t* L_xy, uint8_t* R_xy) {
  if (!st || !L_xy || !R_xy) return BPGPU_E_ARG;
  if (st->n_cur < 2) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(st->ctx->device));
  return st->ctx->curve == BPGPU_BLS12_381 ? ipp_round_t<Bls>(st, L_xy, R_xy) : ipp_round_t<Bn>(st, L_xy, R_xy);
}

int bpgpu_ipp_fold(bpgpu_ipp* st, const uint8_t* u_be, const uint8_t* u_inv_be) {
  if (!st || !u_be || !u_inv_be) return BPGPU_E_ARG;
  if (st->n_cur < 2) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(st->ctx->device));
  return st->ctx->curve == BPGPU_BLS12_381 ? ipp_fold_t<Bls>(st, u_be, u_inv_be) : ipp_fold_t<Bn>(st, u_be, u_inv_be);
}

int bpgpu_ipp_finish(bpgpu_ipp* st, uint8_t* a_be, uint8_t* b_be) {
  if (!st || !a_be || !b_be) return BPGPU_E_ARG;
  if (st->n_cur != 1) return BPGPU_E_ARG;
  BP_CUDA_OK(cudaSetDevice(st->ctx->device));
  return st->ctx->curve == BPGPU_BLS12_381 ? ipp_finish_t<Bls>(st, a_be, b_be) : ipp_finish_t<Bn>(st, a_be, b_be);
}

void bpgpu_ipp_free(bpgpu_ipp* st) {
  if (!st) return;
  cudaSetDevice(st->ctx->device);
  dev_free(st->ctx, st->P);
  dev_free(st->ctx, st->a);                              // a, b, sG, sH, the scalar lists and the row lists: one allocation
  delete st;
}

int bpgpu_ipp_verification_scalars(bpgpu_ctx* ctx, const uint8_t* u_be, size_t lg, bpgpu_scalars** s_out) {
  if (!ctx || (!u_be && lg) || !s_out) return BPGPU_E_ARG;
  if (lg >= 32) return BPGPU_E_VERIFY;                                             // ipp.rs:269-273
  int rc = bpgpu_scalars_alloc(ctx, (size_t)1 << lg, s_out);
  if (rc) return rc;
  rc = ctx->curve == BPGPU_BLS12_381 ? ipp_s_t<Bls>(ctx, u_be, lg, (*s_out)->d) : ipp_s_t<Bn>(ctx, u_be, lg, (*s_out)->d);
  if (rc) { bpgpu_scalars_free(*s_out); *s_out = nullptr; }
  return rc;
}

int bpgpu_ipp_verify_msm(bpgpu_ctx* ctx, const bpgpu_points* G, size_t goff, const bpgpu_points* H, size_t hoff,
                         const uint8_t* Q_xy, const bpgpu_scalars* Gf, const bpgpu_scalars* Hf, const uint8_t* a_be,
                         const uint8_t* b_be, const uint8_t* u_be, const uint8_t* L_xy, const uint8_t* R_xy, size_t lg,
                         uint8_t* out_xy) {
  if (!ctx || !G || !H || !Q_xy || !Gf || !Hf || !a_be || !b_be || !out_xy) return BPGPU_E_ARG;
  if (lg && (!u_be || !L_xy || !R_xy)) return BPGPU_E_ARG;
  if (lg >= 32) return BPGPU_E_VERIFY;
  const size_t n = (size_t)1 << lg;
  if (!prange(G, goff, n) || !prange(H, hoff, n) || !srange(Gf, 0, n) || !srange(Hf, 0, n)) return BPGPU_E_LEN;
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  const void* g = (const uint8_t*)G->d + goff * psize(ctx);
  const void* h = (const uint8_t*)H->d + hoff * psize(ctx);
  return ctx->curve == BPGPU_BLS12_381 ? ipp_verify_t<Bls>(ctx, g, h, Q_xy, Gf->d, Hf->d, a_be, b_be, u_be, L_xy, R_xy, lg, out_xy)
                                      : ipp_verify_t<Bn>(ctx, g, h, Q_xy, Gf->d, Hf->d, a_be, b_be, u_be, L_xy, R_xy, lg, out_xy);
}

}  // extern "C"
