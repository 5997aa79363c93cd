// Witness generation for the hash gadgets, batched (the caller's step BEFORE the proving path):
// MiMC with its multiplier assignments, and the Poseidon permutation.
//
// /root/reference/src/r1cs/gadgets/helper_constraints/mimc.rs:10-29 `mimc(xl, xr, constants, rounds)`:
//     per round  xl, xr := xr + (xl + c_i)^3, xl ;  the image is xl after the last round
// and mimc.rs:53-77 `enforce_mimc_2_inputs`, whose two multipliers per round hold
//     (l, l, l^2)  and  (l^2, l, l^3)   with l = xl + c_i
// -- the a_L / a_R / a_O assignments a prover of a MiMC preimage statement needs.  One thread per instance: `rounds`
// dependent cubings in Fr (the constants are shared by all instances), witness vectors written where the commitment MSMs
// and the polynomial kernels read them (instance-major, 2 * rounds multipliers each).
#include "common.cuh"

namespace bp {

template <class Curve>
__global__ void __launch_bounds__(128) k_mimc_witness(uint32_t count, uint32_t rounds, const typename Curve::Fr* __restrict__ xl_in,
                                                      const typename Curve::Fr* __restrict__ xr_in, const typename Curve::Fr* __restrict__ consts,
                                                      typename Curve::Fr* __restrict__ image, typename Curve::Fr* __restrict__ aL,
                                                      typename Curve::Fr* __restrict__ aR, typename Curve::Fr* __restrict__ aO) {
  using Fr = typename Curve::Fr;
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= count) return;
  Fr xl = load_vec(xl_in + b), xr = load_vec(xr_in + b);
  const size_t base = (size_t)b * 2 * rounds;
  for (uint32_t i = 0; i < rounds; i++) {
    const Fr l = xl + load_vec(consts + i);           // mimc.rs:21
    const Fr sq = l.sqr();
    const Fr cube = sq * l;                            // :22
    if (aL) {                                          // mimc.rs:70-71: multiply(l, l) -> l^2 ; multiply(l^2, l) -> l^3
      store_vec(aL + base + 2 * i, l);      store_vec(aR + base + 2 * i, l);      store_vec(aO + base + 2 * i, sq);
      store_vec(aL + base + 2 * i + 1, sq); store_vec(aR + base + 2 * i + 1, l);  store_vec(aO + base + 2 * i + 1, cube);
    }
    const Fr nl = cube + xr;                           // :23-25
    xr = xl;
    xl = nl;
  }
  store_vec(image + b, xl);
}

// /root/reference/src/r1cs/gadgets/helper_constraints/poseidon.rs:202-293 `Poseidon_permutation`: full rounds (round key +
// S-box on every element), partial rounds (round keys on every element, S-box on the LAST one), full rounds; after each
// S-box layer the linear layer new[i] = sum_j state[j] * MDS[j][i].  S-box (poseidon.rs:122-138): 0 cube, 1 inverse
// (0 -> 0 as FieldElement::inverse), 2 quint.  Round keys and the MDS matrix are the caller's (the reference's tables are
// parameters, not code); one thread per instance, the state (width <= 9) in registers / local memory.
static const int POSEIDON_MAX_WIDTH = 9;
template <class Fr>
__device__ __forceinline__ Fr poseidon_sbox(const Fr& e, int sbox) {
  if (sbox == 1) return e.inv();
  const Fr sq = e.sqr();
  if (sbox == 0) return sq * e;
  return sq.sqr() * e;
}
template <class Curve>
__global__ void __launch_bounds__(128) k_poseidon(uint32_t count, uint32_t width, uint32_t full_begin, uint32_t partial, uint32_t full_end, int sbox,
                                                  const typename Curve::Fr* __restrict__ input, const typename Curve::Fr* __restrict__ keys,
                                                  const typename Curve::Fr* __restrict__ mds, typename Curve::Fr* __restrict__ out) {
  using Fr = typename Curve::Fr;
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= count) return;
  Fr st[POSEIDON_MAX_WIDTH], tmp[POSEIDON_MAX_WIDTH];
  for (uint32_t i = 0; i < width; i++) st[i] = load_vec(input + (size_t)b * width + i);
  uint32_t off = 0;
  const uint32_t total = full_begin + partial + full_end;
  for (uint32_t r = 0; r < total; r++) {
    const bool full = r < full_begin || r >= full_begin + partial;
    for (uint32_t i = 0; i < width; i++) {
      st[i] = st[i] + load_vec(keys + off + i);
      if (full || i == width - 1) st[i] = poseidon_sbox(st[i], sbox);
    }
    off += width;
    for (uint32_t i = 0; i < width; i++) {
      Fr acc = Fr::zero();
      for (uint32_t j = 0; j < width; j++) acc = acc + st[j] * load_vec(mds + j * width + i);
      tmp[i] = acc;
    }
    for (uint32_t i = 0; i < width; i++) st[i] = tmp[i];
  }
  for (uint32_t i = 0; i < width; i++) store_vec(out + (size_t)b * width + i, st[i]);
}

}  // namespace bp

using namespace bp;

extern "C" int bpgpu_poseidon_permutation(bpgpu_ctx* ctx, const bpgpu_scalars* input, size_t count, size_t width, size_t full_rounds_beginning,
                                          size_t partial_rounds, size_t full_rounds_end, int sbox, const bpgpu_scalars* round_keys,
                                          const bpgpu_scalars* mds, bpgpu_scalars** out) {
  if (!ctx || !input || !round_keys || !mds || !out || sbox < 0 || sbox > 2 || count >= (1ull << 28)) return BPGPU_E_ARG;
  if (width != 3 && width != 5 && width != 9) return BPGPU_E_ARG;                        // poseidon.rs:32-34
  const size_t total = full_rounds_beginning + partial_rounds + full_rounds_end;
  if (total >= (1u << 16)) return BPGPU_E_ARG;
  if (input->n < count * width || round_keys->n < total * width || mds->n < width * width) return BPGPU_E_LEN;   // poseidon.rs:58-64,92-96
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  int rc = bpgpu_scalars_alloc(ctx, count * width, out);
  if (rc || count == 0) return rc;
  const unsigned blocks = (unsigned)((count + 127) / 128);
  if (ctx->curve == BPGPU_BLS12_381)
    k_poseidon<Bls><<<blocks, 128, 0, ctx->stream>>>((uint32_t)count, (uint32_t)width, (uint32_t)full_rounds_beginning, (uint32_t)partial_rounds,
                                                      (uint32_t)full_rounds_end, sbox, (const Bls::Fr*)input->d, (const Bls::Fr*)round_keys->d,
                                                      (const Bls::Fr*)mds->d, (Bls::Fr*)(*out)->d);
  else
    k_poseidon<Bn><<<blocks, 128, 0, ctx->stream>>>((uint32_t)count, (uint32_t)width, (uint32_t)full_rounds_beginning, (uint32_t)partial_rounds,
                                                     (uint32_t)full_rounds_end, sbox, (const Bn::Fr*)input->d, (const Bn::Fr*)round_keys->d,
                                                     (const Bn::Fr*)mds->d, (Bn::Fr*)(*out)->d);
  ctx->launches++;
  rc = launch_check(ctx, "k_poseidon");
  if (rc) { bpgpu_scalars_free(*out); *out = nullptr; }
  return rc;
}

extern "C" int bpgpu_mimc_witness(bpgpu_ctx* ctx, const bpgpu_scalars* xl, const bpgpu_scalars* xr, size_t count, const bpgpu_scalars* constants,
                                  size_t rounds, bpgpu_scalars** image, bpgpu_scalars** a_L, bpgpu_scalars** a_R, bpgpu_scalars** a_O) {
  if (!ctx || !xl || !xr || !constants || !image || count >= (1ull << 31) || rounds >= (1u << 20)) return BPGPU_E_ARG;
  if ((a_L || a_R || a_O) && !(a_L && a_R && a_O)) return BPGPU_E_ARG;
  if (xl->n < count || xr->n < count || constants->n < rounds) return BPGPU_E_LEN;     // mimc.rs:18 assert_eq!(constants.len(), mimc_rounds)
  BP_CUDA_OK(cudaSetDevice(ctx->device));
  *image = nullptr;
  if (a_L) { *a_L = nullptr; *a_R = nullptr; *a_O = nullptr; }
  int rc = bpgpu_scalars_alloc(ctx, count, image);
  if (!rc && a_L) {
    bpgpu_scalars** outs[3] = {a_L, a_R, a_O};
    rc = scalars_alloc_many(ctx, count * 2 * rounds, 3, outs);
  }
  if (!rc && count) {
    const unsigned blocks = (unsigned)((count + 127) / 128);
    if (ctx->curve == BPGPU_BLS12_381)
      k_mimc_witness<Bls><<<blocks, 128, 0, ctx->stream>>>((uint32_t)count, (uint32_t)rounds, (const Bls::Fr*)xl->d, (const Bls::Fr*)xr->d,
                                                            (const Bls::Fr*)constants->d, (Bls::Fr*)(*image)->d, a_L ? (Bls::Fr*)(*a_L)->d : nullptr,
                                                            a_L ? (Bls::Fr*)(*a_R)->d : nullptr, a_L ? (Bls::Fr*)(*a_O)->d : nullptr);
    else
      k_mimc_witness<Bn><<<blocks, 128, 0, ctx->stream>>>((uint32_t)count, (uint32_t)rounds, (const Bn::Fr*)xl->d, (const Bn::Fr*)xr->d,
                                                           (const Bn::Fr*)constants->d, (Bn::Fr*)(*image)->d, a_L ? (Bn::Fr*)(*a_L)->d : nullptr,
                                                           a_L ? (Bn::Fr*)(*a_R)->d : nullptr, a_L ? (Bn::Fr*)(*a_O)->d : nullptr);
    ctx->launches++;
    rc = launch_check(ctx, "k_mimc_witness");
  }
  if (rc) {
    bpgpu_scalars_free(*image); *image = nullptr;
    if (a_L) { bpgpu_scalars_free(*a_L); bpgpu_scalars_free(*a_R); bpgpu_scalars_free(*a_O); *a_L = *a_R = *a_O = nullptr; }
  }
  return rc;
}
