/* bphost.h — C entry points of the HOST layer of libbpgpu: the C++ mirror of lovesh/bulletproofs-amcl's
 * prover / verifier API (the headers under bulletproofs-amcl_b200/host) flattened to plain bytes, so that it can be driven
 * from C, Python (ctypes) or a Rust test harness.  Paths below are relative to /root/reference/src.
 *
 * The Rust crate keeps its own host side (Merlin transcript, constraint bookkeeping) and binds include/bpgpu.h
 * directly; these entry points exist because no Rust toolchain is available where this library
 * is built and tested: they run the reference's host logic restated in C++ above the same device ABI, so that the
 * reference's own tests (ipp.rs:318-490, bound_check.rs:188-225) can be replayed end to end.
 *
 * Conventions: as include/bpgpu.h (scalars MODBYTES big endian, points X||Y).  A serialised IPP proof is
 *   L_1..L_lg | R_1..R_lg  (G1::to_bytes(): 0x04||X||Y each)  |  a | b  (FieldElement::to_bytes()).
 * A serialised R1CS proof is  A_I1 A_O1 S1 A_I2 A_O2 S2 T_1 T_3 T_4 T_5 T_6 | t_x t_x_blinding e_blinding | IPP proof
 * (proof.rs:26-58 field order).  Return values: BPGPU_OK, or a negative BPGPU_E_* / BPH_E_* code.
 *
 * Randomness: the reference draws blindings from OS entropy (FieldElement::random()).  Every prover entry point
 * takes (rng_mode, seed): mode 0 = OS entropy, mode 1 = the deterministic stream
 * SHAKE256(seed_le64 || "blind" || i_le64) mod r consumed in the reference's draw order, which is what
 * the parity tests use.  The verifier's batching scalar (verifier.rs:392) is an explicit argument (NULL = OS entropy).
 */
#ifndef BPHOST_H
#define BPHOST_H

#include "bpgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define BPH_E_MISSING_ASSIGNMENT (-8) /* R1CSError::MissingAssignment */
#define BPH_E_GADGET (-9)             /* R1CSError::GadgetError */
#define BPH_E_BUFFER (-10)            /* output buffer too small; *len holds the required size */
#define BPH_E_ENTROPY (-11)           /* the OS could not supply entropy for blindings (getrandom failed): nothing was proved */

/* merlin::Transcript known-answer hook: new(label); append_message(msg_label, msg); challenge_bytes(ch_label, out_len) */
int bph_merlin_kat(const char* label, const char* msg_label, const uint8_t* msg, size_t msg_len, const char* ch_label,
                   uint8_t* out, size_t out_len);
/* TranscriptProtocol::challenge_scalar after commit_point / commit_scalar (transcript.rs:47-60) on the ctx curve */
int bph_transcript_kat(int curve, const char* label, const uint8_t* point_xy, const uint8_t* scalar_be, uint8_t* challenge_be);

/* utils/mod.rs:16-23 get_generators(prefix, n): G_i = G1::from_msg_hash(prefix || decimal(i)), i = 1..n, as a
 * device-resident table.  SHAKE256 runs on the host, try-and-increment + sqrt + cofactor clearing on the device. */
int bph_get_generators(bpgpu_ctx* ctx, const char* prefix, size_t n, bpgpu_points** out);
/* G1::from_msg_hash(msg) */
int bph_g1_from_msg_hash(bpgpu_ctx* ctx, const uint8_t* msg, size_t msg_len, uint8_t* out_xy);

/* IPP::create_ipp (ipp.rs:35-202) with a fresh Transcript::new(transcript_label) */
int bph_ipp_create(bpgpu_ctx* ctx, const char* transcript_label, const bpgpu_points* G, const bpgpu_points* H, const uint8_t* Q_xy,
                   const uint8_t* G_factors_be, const uint8_t* H_factors_be, const uint8_t* a_be, const uint8_t* b_be, size_t n,
                   uint8_t* proof, size_t cap, size_t* len);
/* IPP::verify_ipp (ipp.rs:204-260): BPGPU_OK or BPGPU_E_VERIFY */
int bph_ipp_verify(bpgpu_ctx* ctx, const char* transcript_label, size_t n, const uint8_t* G_factors_be, const uint8_t* H_factors_be,
                   const uint8_t* P_xy, const uint8_t* Q_xy, const bpgpu_points* G, const bpgpu_points* H, const uint8_t* proof,
                   size_t len);

/* gen_proof_of_bounded_num / verify_proof_of_bounded_num (gadgets/bound_check.rs:133-178).
 * randomness_be: blinding of the commitment to val (NULL = drawn from the rng).  comms_xy receives 3 points. */
int bph_bound_check_prove(bpgpu_ctx* ctx, const char* transcript_label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G,
                          const bpgpu_points* H, uint64_t val, const uint8_t* randomness_be, uint64_t lower, uint64_t upper,
                          size_t max_bits_in_val, int rng_mode, uint64_t seed, uint8_t* proof, size_t cap, size_t* len,
                          uint8_t* comms_xy);
int bph_bound_check_verify(bpgpu_ctx* ctx, const char* transcript_label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G,
                           const bpgpu_points* H, uint64_t lower, uint64_t upper, size_t max_bits_in_val, const uint8_t* proof,
                           size_t len, const uint8_t* comms_xy, const uint8_t* verifier_r_be);

/* m values, each proven in [0, 2^bits) with positive_no_gadget (helper_constraints/positive_no.rs:8-40) in ONE
 * constraint system: n = m*bits multipliers (BASELINE.json configs 2, 3, 5).  comms_xy receives m points. */
int bph_range_prove(bpgpu_ctx* ctx, const char* transcript_label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G,
                    const bpgpu_points* H, const uint64_t* values, size_t m, size_t bits, int rng_mode, uint64_t seed, uint8_t* proof,
                    size_t cap, size_t* len, uint8_t* comms_xy);
int bph_range_verify(bpgpu_ctx* ctx, const char* transcript_label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G,
                     const bpgpu_points* H, size_t m, size_t bits, const uint8_t* proof, size_t len, const uint8_t* comms_xy,
                     const uint8_t* verifier_r_be);

/* ---- batches of independent proofs (BASELINE.json config 5; throughput runs of configs 1-3) -----------------
 * The reference verifies one proof per Verifier::verify call (verifier.rs:267) and has no batch API; per-proof verdicts are
 * kept (proofs are NOT merged with random weights).  `count` proofs are spread over `nctx` contexts of ONE device, one host
 * thread per context (a context = one CUDA stream), all sharing the generator tables G, H.  Proof i uses values
 * values[i*m .. i*m+m), blinding seed seed+i (rng_mode 1) and the transcript label given.  Proofs are fixed-size records of
 * `proof_stride` bytes (bph_range_proof_len), commitments m*2*MODBYTES bytes each.
 * verdicts[i] = 0 (Ok), or the negative error code of proof i (BPGPU_E_VERIFY = VerificationError). */
size_t bph_range_proof_len(int curve, size_t m, size_t bits);
int bph_range_prove_many(bpgpu_ctx* const* ctxs, size_t nctx, const char* transcript_label, const uint8_t* g_xy, const uint8_t* h_xy,
                         const bpgpu_points* G, const bpgpu_points* H, const uint64_t* values, size_t count, size_t m, size_t bits,
                         int rng_mode, uint64_t seed, uint8_t* proofs, size_t proof_stride, uint8_t* comms_xy);
int bph_range_verify_many(bpgpu_ctx* const* ctxs, size_t nctx, const char* transcript_label, const uint8_t* g_xy, const uint8_t* h_xy,
                          const bpgpu_points* G, const bpgpu_points* H, size_t count, size_t m, size_t bits, const uint8_t* proofs,
                          size_t proof_stride, const uint8_t* comms_xy, int32_t* verdicts);

/* The same verdicts as bph_range_verify_many from ONE batched device call per slab of <= 4096 proofs
 * (bpgpu_r1cs_verify_batch, csrc/verifybatch.cu): the circuit (m x positive_no_gadget) is recorded once as a sparse matrix
 * (Verifier::export_csr -> bpgpu_circuit), the host uploads proof and commitment bytes, and the device replays every proof's
 * transcript (one thread per proof), builds its O(N) verification scalars (one block per proof: flattened constraints from
 * z, the s vector, y^-i, delta, g/h scalars, head scalars) and evaluates its verification MSM -- one verdict per proof,
 * nothing merged.  G and H get window tables on first use (bpgpu_points_precompute).
 * bph_range_verify_batch_mode selects where the transcripts run:
 *   mode 0  on the device (default; what bph_range_verify_batch does; `nthreads` unused);
 *   mode 1  on `nthreads` host threads (0 = all cores) as the reference's design has it; only the 5 + lg challenges per
 *           proof are uploaded, everything else as mode 0;
 *   mode 2  round 1's path: transcripts AND all scalars on host threads (bpgpu_msm_batch_is_identity) -- kept for A/B runs. */
int bph_range_verify_batch(bpgpu_ctx* ctx, const char* transcript_label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G,
                           bpgpu_points* H, size_t count, size_t m, size_t bits, const uint8_t* proofs, size_t proof_stride,
                           const uint8_t* comms_xy, size_t nthreads, int32_t* verdicts);
int bph_range_verify_batch_mode(bpgpu_ctx* ctx, const char* transcript_label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G,
                                bpgpu_points* H, size_t count, size_t m, size_t bits, const uint8_t* proofs, size_t proof_stride,
                                const uint8_t* comms_xy, int mode, size_t nthreads, int32_t* verdicts);
/* verify_proof_of_bounded_num (gadgets/bound_check.rs:163-178) for `count` proofs over the same [lower, upper] and
 * max_bits_in_val: 3 commitments per proof (v, v - lower, upper - v), n = 2 * bits multipliers; modes 0 and 1 as above.
 * Proof records as bph_range_proof_len(curve, 2, bits). */
int bph_bound_check_verify_batch(bpgpu_ctx* ctx, const char* transcript_label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G,
                                 bpgpu_points* H, size_t count, uint64_t lower, uint64_t upper, size_t max_bits_in_val,
                                 const uint8_t* proofs, size_t proof_stride, const uint8_t* comms_xy, int mode, size_t nthreads,
                                 int32_t* verdicts);
/* The recorded circuits as CSR arrays (host only, no device needed; see bpgpu_circuit_create for the layout).  Call once
 * with the array pointers NULL to get n, m, q, nnz, then with buffers of 3n + m + 2, nnz and nnz * MODBYTES entries. */
int bph_range_circuit_csr(int curve, size_t m, size_t bits, size_t* n, size_t* m_out, size_t* q, size_t* nnz, uint32_t* row_start,
                          uint32_t* ent_q, uint8_t* ent_coeff_be);
int bph_bound_check_circuit_csr(int curve, uint64_t lower, uint64_t upper, size_t bits, size_t* n, size_t* m_out, size_t* q, size_t* nnz,
                                uint32_t* row_start, uint32_t* ent_q, uint8_t* ent_coeff_be);
/* process-wide: the prover's witness commitments A_I, A_O, S (prover.rs:347-362: inner_product_const_time in the
 * reference) use fixed-schedule table sums (bpgpu_ctx_set_fixed_schedule) when the generators carry window tables.  Off by
 * default (BPH_FIXED_SCHEDULE=1 in the environment turns it on); the proof bytes do not change. */
void bph_set_secret_fixed_schedule(int on);
/* exported Merlin state (203 bytes) after Transcript::new(label) + r1cs_domain_sep(): bpgpu_r1cs_verify_batch's transcript_state */
void bph_r1cs_transcript_state(const char* transcript_label, uint8_t* out203);
/* the challenges y, z, u, x, w, u_1..u_lg ((5 + lg) x MODBYTES, big endian) of one proof of a one-phase circuit with m
 * commitments from a HOST transcript (verifier.rs:279-323, ipp.rs:278-288): bpgpu_r1cs_verify_batch's challenges_be */
int bph_r1cs_replay_challenges(int curve, const char* transcript_label, const uint8_t* proof, const uint8_t* comms_xy, size_t m, size_t lg,
                               uint8_t* out_be);

/* `count` independent range proofs (m x `bits`-bit values each) proved in LOCK-STEP on one context (bpgpu_pbatch_*): every
 * prover stage and every IPP round is one device call for a whole slab of proofs.  Same arguments and same bytes as
 * bph_range_prove_many (proof i uses seed + i); G and H get window tables on first use.  bph_range_prove_batch_mode selects
 * where the transcripts run:
 *   mode 0  on the device (default; what bph_range_prove_batch does; `nthreads` unused): bpgpu_pbatch_prove_range -- values
 *           and blinding keys go up, finished proof records come down, nothing in between;
 *   mode 1  on `nthreads` host threads (0 = all cores) between the device stages, as the reference's design has it. */
int bph_range_prove_batch(bpgpu_ctx* ctx, const char* transcript_label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G,
                          bpgpu_points* H, const uint64_t* values, size_t count, size_t m, size_t bits, int rng_mode, uint64_t seed,
                          size_t nthreads, uint8_t* proofs, size_t proof_stride, uint8_t* comms_xy);
int bph_range_prove_batch_mode(bpgpu_ctx* ctx, const char* transcript_label, const uint8_t* g_xy, const uint8_t* h_xy, bpgpu_points* G,
                               bpgpu_points* H, const uint64_t* values, size_t count, size_t m, size_t bits, int rng_mode, uint64_t seed,
                               int mode, size_t nthreads, uint8_t* proofs, size_t proof_stride, uint8_t* comms_xy);

/* ---- a TWO-PHASE circuit: k-shuffle ({y} is a permutation of {x}) over 2k committed values, x[0] range-checked to `bits`
 * bits in the first phase (0 = none).  The gadget is the example of the reference's ConstraintSystem documentation
 * (constraint_system.rs:86-135); it exercises specify_randomized_constraints, the challenge drawn between the phases and
 * the second-phase commitments A_I2, A_O2, S2 (prover.rs:300-319,384-436; verifier.rs:245-264).  n = bits + 2(k-1)
 * multipliers; comms_xy = the 2k commitments, x first.  Proof layout as bph_range_proof_len(curve, 1, n) bytes. */
int bph_shuffle_prove(bpgpu_ctx* ctx, const char* transcript_label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G,
                      const bpgpu_points* H, const uint64_t* x, const uint64_t* y, size_t k, size_t bits, int rng_mode, uint64_t seed,
                      uint8_t* proof, size_t cap, size_t* len, uint8_t* comms_xy);
int bph_shuffle_verify(bpgpu_ctx* ctx, const char* transcript_label, const uint8_t* g_xy, const uint8_t* h_xy, const bpgpu_points* G,
                       const bpgpu_points* H, size_t k, size_t bits, const uint8_t* proof, size_t len, const uint8_t* comms_xy,
                       const uint8_t* verifier_r_be);

/* ---- MSM sharded by points over several contexts: one process, one context per GPU (or several per GPU) ----
 * shard k = `counts[k]` points resident on ctxs[k]'s device (bpgpu_points_upload there), its scalars are the next counts[k]
 * entries of scalars_be.  One host thread per context runs bpgpu_msm on its shard; the affine partial sums (2*MODBYTES
 * bytes each -- the only data that leaves a GPU) are added on the host.  The total is a group element, so the result
 * bytes do not depend on how the points were split.  (With one PROCESS per GPU the same partials travel through
 * ncclAllGather instead: bulletproofs-amcl_b200/sharding.py, bench.py.) */
int bph_msm_sharded(bpgpu_ctx* const* ctxs, size_t nctx, const bpgpu_points* const* shards, const size_t* counts, const uint8_t* scalars_be,
                    uint8_t* out_xy);
/* sum of `count` affine points given as X||Y (identity = (0, 1)), on the host: the combine step of the sharded MSM */
int bph_g1_sum(int curve, const uint8_t* points_xy, size_t count, uint8_t* out_xy);

#ifdef __cplusplus
}
#endif
#endif /* BPHOST_H */
