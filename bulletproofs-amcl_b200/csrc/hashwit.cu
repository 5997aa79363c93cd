// Witness generation for the MiMC hash gadget, batched (the caller's step BEFORE the proving path).
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

}  // namespace bp

using namespace bp;

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
