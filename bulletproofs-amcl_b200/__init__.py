"""bulletproofs-amcl_b200 — B200-native G1 MSM / inner-product-argument hot path.

Host-side Python binding (ctypes) over the C ABI in include/bpgpu.h.  The compute lives in
`libbpgpu.so` (hand-written sm_100a CUDA, csrc/); this package only marshals bytes.
There is no CPU fallback: importing works anywhere (so the ABI can be inspected), but creating
a context without the built library or without a CUDA device raises.
"""
from .binding import (BLS12_381, BN254, BpgpuError, Circuit, Context, DevicePoints, DeviceScalars, bound_check_circuit_csr,
                      bound_check_verify_batch, build_library, g1_sum, lib, library_path, msm_sharded, r1cs_replay_challenges, r1cs_transcript_state,
                      range_circuit_csr, range_prove_batch, range_prove_many, range_verify_batch, range_verify_many)

__all__ = ["BLS12_381", "BN254", "BpgpuError", "Circuit", "Context", "DevicePoints", "DeviceScalars", "bound_check_circuit_csr",
           "bound_check_verify_batch", "build_library", "g1_sum", "lib", "library_path", "msm_sharded", "r1cs_replay_challenges", "r1cs_transcript_state",
           "range_circuit_csr", "range_prove_batch", "range_prove_many", "range_verify_batch", "range_verify_many"]
