/* bpgpu.h — C ABI of the B200-native G1 MSM / inner-product-argument hot path.
 *
 * This is the drop-in boundary for lovesh/bulletproofs-amcl's data-parallel path.  The reference
 * has no FFI of its own: its seam is the set of `amcl_wrapper` vector methods it calls
 * (SURVEY.md section 8b).  Each entry point below names the reference call sites it replaces
 * (paths relative to /root/reference/src).  INTEGRATION.md shows the Rust `extern "C"` binding.
 *
 * Conventions
 *  - curve ids: BPGPU_BLS12_381 (Cargo feature `bls381`, default) and BPGPU_BN254 (feature
 *    `bn254` = AMCL's Nogami BN254).  MODBYTES = 48 / 32.
 *  - A scalar (Fr) on the ABI is exactly `FieldElement::to_bytes()`: MODBYTES bytes, big endian,
 *    value < r (transcript.rs:47-49).
 *  - A point (G1) on the ABI is `G1::to_bytes()` WITHOUT its leading 0x04 tag: X || Y, MODBYTES
 *    big endian each (transcript.rs:51-53).  The identity is AMCL's (X, Y) = (0, 1).
 *  - All pointers are plain host pointers unless the parameter is an opaque device handle.
 *    Host buffers may be pageable; pinned buffers (bpgpu_host_alloc) make copies asynchronous.
 *  - Every function returns BPGPU_OK (0) or a negative error; nothing throws, nothing calls back.
 *  - One bpgpu_ctx per host thread (the reference is single threaded: prover.rs:36 holds
 *    `&mut Transcript`).  A ctx owns one CUDA stream and its scratch memory.
 *  - There is NO CPU fallback: if no CUDA device is usable, bpgpu_ctx_create fails with
 *    BPGPU_E_CUDA.
 */
#ifndef BPGPU_H
#define BPGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPGPU_BLS12_381 0
#define BPGPU_BN254 1

#define BPGPU_OK 0
#define BPGPU_E_LEN (-1)      /* amcl_wrapper ValueError::UnequalSizeVectors (ipp.rs:77-104 unwraps) */
#define BPGPU_E_NOT_POW2 (-2) /* create_ipp's assert!(n.is_power_of_two()) (ipp.rs:48) */
#define BPGPU_E_GENS_LEN (-3) /* R1CSError::InvalidGeneratorsLength (prover.rs:332-334, verifier.rs:297-299) */
#define BPGPU_E_VERIFY (-4)   /* R1CSError::VerificationError (ipp.rs:258,272,275; verifier.rs:360,453) */
#define BPGPU_E_FORMAT (-5)   /* R1CSError::FormatError / bad encoding */
#define BPGPU_E_CUDA (-6)     /* CUDA runtime failure or no device */
#define BPGPU_E_ARG (-7)      /* null pointer / bad curve id / out-of-range offset */

typedef struct bpgpu_ctx bpgpu_ctx;
typedef struct bpgpu_points bpgpu_points;   /* device-resident G1Vector (affine, Montgomery form) */
typedef struct bpgpu_scalars bpgpu_scalars; /* device-resident FieldElementVector (Montgomery form) */
typedef struct bpgpu_ipp bpgpu_ipp;         /* device-resident state of one create_ipp run */
typedef struct bpgpu_fixed_bases bpgpu_fixed_bases; /* window tables of a few fixed bases, e.g. the Pedersen pair (g, h) */
typedef struct bpgpu_circuit bpgpu_circuit;         /* a one-phase constraint system as a device-resident sparse matrix */

const char* bpgpu_strerror(int code);
int bpgpu_modbytes(int curve);              /* amcl_wrapper::constants::MODBYTES */
int bpgpu_device_count(void);

int bpgpu_ctx_create(int curve, int device, bpgpu_ctx** out);
void bpgpu_ctx_destroy(bpgpu_ctx* ctx);
/* a second context (stream + scratch) on the same device, created on first use, owned by and destroyed with `ctx`: the other
 * driver thread of the batch calls, which keep two slabs in flight (a context costs ~20 ms to set up: not per call) */
int bpgpu_ctx_aux(bpgpu_ctx* ctx, bpgpu_ctx** out);
void* bpgpu_ctx_stream(bpgpu_ctx* ctx);     /* the cudaStream_t all work of this ctx is issued on */
int bpgpu_ctx_sync(bpgpu_ctx* ctx);
int bpgpu_ctx_curve(const bpgpu_ctx* ctx);
int bpgpu_ctx_device(const bpgpu_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
uint64_t bpgpu_ctx_launches(const bpgpu_ctx* ctx);

/* Secret scalars (the reference commits to the witness and its blindings with inner_product_const_time /
 * commit_to_field_element_vectors, prover.rs:347-362): while this switch is on, every MSM of the ctx that runs on window
 * tables (bpgpu_points_precompute) uses a FIXED schedule -- 32 table loads and 32 mixed additions per term whatever the
 * scalar's digits are, zero digits and zero scalars included (their sums are discarded by a lane-wise select).  It does
 * NOT make the fetched table addresses independent of the scalars (one entry per digit is read, not a masked scan of the
 * row), and MSMs that take the bucket path (no tables) are variable time regardless: see DESIGN.md "secret scalars". */
int bpgpu_ctx_set_fixed_schedule(bpgpu_ctx* ctx, int on);
/* How the host waits for this context's stream: 0 (default) spins -- lowest latency, one core per context; 1 sleeps on a
 * blocking event, so that MORE contexts than host cores can keep independent proofs in flight (one proof per context is
 * round-synchronous: with c cores, c spinning contexts leave the GPU idle between rounds).  Also BPGPU_BLOCKING_SYNC=1 at
 * creation.  Measured with 4 cores on config 3: 4 spinning contexts 307 proofs/s, 32 sleeping ones 416. */
int bpgpu_ctx_set_blocking_sync(bpgpu_ctx* ctx, int on);
/* per-stage CUDA-event timing of the MSM pipeline on the ctx stream (roofline evidence for bench.py).
 * stages: 0 digits, 1 scan, 2 scatter, 3 chunk_acc, 4 giant, 5 merge, 6 reduce_l1, 7 reduce_l2.
 * set_profile(min_n > 0) records every MSM of at least min_n terms (0 = off); bpgpu_msm_stage_ms writes the average ms
 * per stage since then and returns the run count. */
int bpgpu_ctx_set_profile(bpgpu_ctx* ctx, int min_n);
int bpgpu_msm_stage_ms(const bpgpu_ctx* ctx, double* avg_ms, int cap);

void* bpgpu_host_alloc(size_t bytes);       /* pinned host memory */
void bpgpu_host_free(void* p);

/* ---- G1Vector / FieldElementVector residency (G1Vector::from / FieldElementVector::from) ---- */
int bpgpu_points_upload(bpgpu_ctx* ctx, const uint8_t* xy, size_t n, bpgpu_points** out);
int bpgpu_points_download(bpgpu_ctx* ctx, const bpgpu_points* p, size_t off, size_t n, uint8_t* xy);
/* G1Vector of hash-derived generators: out[i] = ECP::mapit(hashes[i]) where hashes[i] = SHAKE256(msg_i)[..MODBYTES] is
 * computed by the caller (G1::from_msg_hash = hash_msg + mapit; utils/mod.rs:16-23 get_generators builds every
 * generator table this way).  Try-and-increment, square root, even-y choice and cofactor clearing run on the device. */
int bpgpu_points_from_hashes(bpgpu_ctx* ctx, const uint8_t* hashes, size_t n, bpgpu_points** out);
/* Window tables for a generator table: T[i][w][d] = d * 2^(8w) * P_i (32 windows x 255 multiples, 8160 affine points =
 * 765 KB (BLS12-381) / 510 KB (BN254) per generator, built once on the device).  Afterwards every MSM whose points come from
 * this handle -- bpgpu_msm, bpgpu_msm_device, bpgpu_msm_parts(_batch) when ALL its parts have tables, and the rounds of
 * bpgpu_ipp_begin_fixed_q -- is a plain sum of 32 table entries per term: no doublings, no buckets, no Horner tail on the
 * host.  Meant for the fixed generators G, H of the proof system (utils/mod.rs:16-23).  It trades HBM (1.6 GB for the 2048
 * generators of a 1024-multiplier circuit) for latency: a clear win up to a few thousand generators (prove n = 64:
 * 10.4 -> 2.4 ms, n = 1024: 29.6 -> 5.3 ms), a loss in throughput mode at 2^14 -- callers precompute accordingly. */
int bpgpu_points_precompute(bpgpu_ctx* ctx, bpgpu_points* p);
/* WIDE tables on top of those: T16[i][W][d] = d * 2^(16 W) * P_i, 16 windows x 65535 multiples = 100 MB (BLS12-381) / 67 MB
 * (BN254) per generator, each entry ONE addition of two entries of the 8-bit tables.  The batch calls (bpgpu_r1cs_verify_batch,
 * bpgpu_msm_batch_is_identity, bpgpu_pbatch_*) then add 16 instead of 32 entries per fixed term.  For the few dozen generators
 * of a small statement proved / verified in bulk (G, H of a 64-bit range proof: 2 x 6.4 GB of the 180 GB); refused above
 * 48 GB per handle.  Single MSMs keep using the 8-bit tables.  has_tables: 0 none, 1 8-bit, 2 both. */
int bpgpu_points_precompute_wide(bpgpu_ctx* ctx, bpgpu_points* p);
int bpgpu_points_has_tables(const bpgpu_points* p);
size_t bpgpu_points_len(const bpgpu_points* p);
void bpgpu_points_free(bpgpu_points* p);
int bpgpu_scalars_upload(bpgpu_ctx* ctx, const uint8_t* be, size_t n, bpgpu_scalars** out);
int bpgpu_scalars_download(bpgpu_ctx* ctx, const bpgpu_scalars* s, size_t off, size_t n, uint8_t* be);
/* a second handle onto s[off .. off+n) (FieldElementVector::split_at without the copies, ipp.rs:70-73): the storage is
 * reference counted and released when the last handle -- s or any view, in any order -- is freed.  Lets several vectors
 * travel in ONE upload (one copy, one conversion launch, one synchronisation) and be handed out as separate vectors.
 * Handles of one allocation must be freed from one thread at a time (as everything on a ctx). */
int bpgpu_scalars_view(bpgpu_scalars* s, size_t off, size_t n, bpgpu_scalars** out);
size_t bpgpu_scalars_len(const bpgpu_scalars* s);
void bpgpu_scalars_free(bpgpu_scalars* s);

/* ---- multi-scalar multiplication: out = sum_i s_i * P_i, normalised affine ------------------
 * Replaces GroupElementVector::multi_scalar_mul_var_time (ipp.rs:251-253,372,471),
 * G1Vector::inner_product_var_time_with_ref_vecs (ipp.rs:91,104,158,170; verifier.rs:451) and
 * inner_product_const_time / commit_to_field_element_vectors (prover.rs:347-362,413-426).
 * The result is the unique affine form, so it does not depend on the summation algorithm. */
/* cached bases P[off .. off+n), host scalars */
int bpgpu_msm(bpgpu_ctx* ctx, const bpgpu_points* p, size_t off, size_t n, const uint8_t* scalars_be,
              uint8_t* out_xy);
/* cached bases, host scalars as 32-byte LITTLE-endian canonical integers (value < r, the caller's responsibility): the
 * wire format of callers that keep scalars as 4 x u64 limbs -- one 32-byte copy per term, no conversion pass */
int bpgpu_msm_le32(bpgpu_ctx* ctx, const bpgpu_points* p, size_t off, size_t n, const uint8_t* scalars_le32,
                   uint8_t* out_xy);
/* The same four calls in two halves, for callers that keep MORE THAN ONE MSM in flight: *_begin enqueues the device
 * pipeline (and the host-to-device copies) on the ctx stream and returns; bpgpu_msm_finish(ctx, out) waits for it and does
 * the host part (Horner over the window sums, one inversion).  One MSM in flight per ctx (a second begin before the finish
 * is BPGPU_E_ARG); host buffers passed to begin must stay valid until finish.  With two contexts on one device, one host
 * thread alternating begin(A) / begin(B) / finish(A) / begin(A) / finish(B) ... runs the sort stages, the copies and the
 * host finish of one MSM under the bucket accumulation of the other (bench.py: 8.6 -> 7.x ms per 2^20-term MSM). */
int bpgpu_msm_begin(bpgpu_ctx* ctx, const bpgpu_points* p, size_t off, size_t n, const uint8_t* scalars_be);
int bpgpu_msm_le32_begin(bpgpu_ctx* ctx, const bpgpu_points* p, size_t off, size_t n, const uint8_t* scalars_le32);
int bpgpu_msm_device_begin(bpgpu_ctx* ctx, const bpgpu_points* p, size_t poff, size_t n, const bpgpu_scalars* s, size_t soff);
int bpgpu_msm_refs_begin(bpgpu_ctx* ctx, const uint8_t* points_xy, const uint8_t* scalars_be, size_t n);
int bpgpu_msm_finish(bpgpu_ctx* ctx, uint8_t* out_xy);
/* cached bases and device-resident scalars */
int bpgpu_msm_device(bpgpu_ctx* ctx, const bpgpu_points* p, size_t poff, size_t n, const bpgpu_scalars* s,
                     size_t soff, uint8_t* out_xy);
/* ad-hoc host point and scalar lists (`Vec<&G1>`, `Vec<&FieldElement>`) */
int bpgpu_msm_refs(bpgpu_ctx* ctx, const uint8_t* points_xy, const uint8_t* scalars_be, size_t n,
                   uint8_t* out_xy);
/* one MSM over the concatenation of several sources.  Each part takes its n points either from a
 * device table (points != NULL, starting at points_off) or from host bytes (host_points_xy), and its n
 * scalars either from a device vector (scalars != NULL) or from host bytes (host_scalars_be).  This is
 * how the reference's `arg2`/`arg1` reference lists are assembled (verifier.rs:394-449: proof points and
 * host scalars around the big G, H tables) and how commit_to_field_element_vectors(G, H, h, a, b, c)
 * (prover.rs:347-362) maps onto cached generator tables. */
typedef struct bpgpu_msm_part {
  const bpgpu_points* points;   size_t points_off;   const uint8_t* host_points_xy;
  const bpgpu_scalars* scalars; size_t scalars_off;  const uint8_t* host_scalars_be;
  size_t n;
} bpgpu_msm_part;
int bpgpu_msm_parts(bpgpu_ctx* ctx, const bpgpu_msm_part* parts, size_t nparts, uint8_t* out_xy);
/* nmsm independent composite MSMs in one call: MSM m takes the next counts[m] entries of `parts`; out_xy receives nmsm
 * points.  This is the prover's A_I, A_O, S triple (prover.rs:347-362, 404-427): with precomputed generator tables the
 * three are evaluated by one launch pair. */
int bpgpu_msm_parts_batch(bpgpu_ctx* ctx, const bpgpu_msm_part* parts, const size_t* counts, size_t nmsm, uint8_t* out_xy);
/* ---- batches of verification MSMs sharing the fixed generators (verifier.rs:451-456 for many proofs at once) -------
 * R_b = sum_i f[b][i] * F_i + sum_j v[b][j] * V[b][j] for b < batch, reduced to is_identity[b] = (R_b == identity).
 * The fixed terms F are the concatenation of `runs` (device tables WITH window tables: bpgpu_points_precompute; or a single
 * host base, whose table is built and cached in the ctx) -- F = sum of runs[k].n; fixed_scalars_be is batch x F scalars,
 * var_points_xy / var_scalars_be are batch x vn proof-specific points and scalars.  Every proof keeps its own verdict:
 * nothing is merged with random weights (the reference has no batch API). */
typedef struct bpgpu_fixed_run { const bpgpu_points* points; size_t off; size_t n; const uint8_t* host_base_xy; } bpgpu_fixed_run;
int bpgpu_msm_batch_is_identity(bpgpu_ctx* ctx, const bpgpu_fixed_run* runs, size_t nruns, size_t batch, const uint8_t* fixed_scalars_be,
                                const uint8_t* var_points_xy, const uint8_t* var_scalars_be, size_t vn, uint8_t* is_identity);
/* ---- batched R1CS verification, per-proof scalars built on the device (csrc/verifybatch.cu) -----------------------------
 * A bpgpu_circuit is the constraint system of a ONE-PHASE circuit in the form Verifier::flattened_constraints
 * (verifier.rs:149-193; Prover: prover.rs:142-184) consumes it, as a sparse matrix shared by every proof of a batch:
 * rows = variables [wL 0..n) | wR 0..n) | wO 0..n) | wV 0..m) | wc] (row_start: 3n + m + 2 offsets), entry e of a row =
 * (ent_q[e] & 0x3fffffff = constraint index k, ent_coeff_be[e]) meaning coefficient * z^(k+1), with the signs of
 * verifier.rs:166-190 applied by the caller (committed and constant terms negated).  Bits 31 / 30 of ent_q flag the
 * coefficients +1 / -1 (the product is skipped).  q = number of constraints.  host/r1cs.hpp Verifier::export_csr builds it.
 * bpgpu_circuit_flatten: [wL | wR | wO | wV | wc] for one challenge z (3n + m + 1 device scalars). */
int bpgpu_circuit_create(bpgpu_ctx* ctx, size_t n, size_t m, size_t q, const uint32_t* row_start, const uint32_t* ent_q,
                         const uint8_t* ent_coeff_be, bpgpu_circuit** out);
void bpgpu_circuit_free(bpgpu_circuit* c);
size_t bpgpu_circuit_multipliers(const bpgpu_circuit* c);
size_t bpgpu_circuit_commitments(const bpgpu_circuit* c);
int bpgpu_circuit_flatten(bpgpu_ctx* ctx, const bpgpu_circuit* c, const uint8_t* z_be, bpgpu_scalars** out);
/* A recorded circuit kept WITH a context under a caller-chosen key (the host layer keys range statements by (m, bits)):
 * put hands ownership to the ctx (freed by bpgpu_ctx_destroy; BPGPU_E_ARG if the key is taken or the circuit belongs to
 * another ctx), get returns NULL when absent.  With it a single proof of a known statement never rebuilds
 * LinearCombinations on the host: weights come from bpgpu_circuit_flatten (prover.rs:142-184), the witness of a range
 * statement from bpgpu_range_witness. */
int bpgpu_ctx_circuit_put(bpgpu_ctx* ctx, uint64_t key, bpgpu_circuit* c);
const bpgpu_circuit* bpgpu_ctx_circuit_get(const bpgpu_ctx* ctx, uint64_t key);
/* [a_L | a_R | a_O] (3 * m * bits device scalars) of m positive_no gadgets over `values` (gadgets/positive_no.rs:18-24:
 * per bit a_L = 1 - bit, a_R = bit, a_O = 0), in the allocation order of the gadget: value-major, least significant bit first */
int bpgpu_range_witness(bpgpu_ctx* ctx, const uint64_t* values, size_t m, size_t bits, bpgpu_scalars** out);
/* Verifier::verify (verifier.rs:267-457) for `count` independent proofs of `circuit`: verdicts[i] = BPGPU_OK,
 * BPGPU_E_VERIFY (VerificationError) or BPGPU_E_FORMAT (missing 0x04 tag, scalar >= r, coordinate >= p, point off the
 * curve).  proofs = count records of proof_stride bytes in the flat form A_I1 A_O1 S1 A_I2 A_O2 S2 T_1 T_3 T_4 T_5 T_6
 * (0x04||X||Y each) | t_x t_x_blinding e_blinding | L_1..L_lg | R_1..R_lg | a | b; comms_xy = count x m commitments X||Y.
 * Challenges come from EXACTLY ONE of
 *   transcript_state: the 203-byte exported Merlin state (200 STROBE state bytes | pos | pos_begin | 0) after
 *                     Transcript::new(label) and r1cs_domain_sep(); the device replays every proof's transcript from there
 *                     (verifier.rs:124-132,279-323; ipp.rs:278-288), one thread per proof;
 *   challenges_be:    count x (5 + lg) scalars y, z, u, x, w, u_1..u_lg from the caller's own transcripts.
 * Everything else -- the circuit's weights from z, the s vector (ipp.rs:295-312), y^-i, delta, g/h scalars
 * (verifier.rs:341-390), the 13 + m head scalars (:392-429) -- is built on the device, one block per proof, straight into
 * the scalar slab of the batched MSM check.  The verifier's random scalar of proof i (verifier.rs:392) is draw i of the
 * counter-mode stream SHAKE256(rnd_key || le64(i)) mod r (host layer: Rng); pass fresh entropy as rnd_key.
 * G and H get window tables on first use.  Nothing is merged across proofs. */
int bpgpu_r1cs_verify_batch(bpgpu_ctx* ctx, const bpgpu_circuit* circuit, bpgpu_points* G, bpgpu_points* H, const uint8_t* g_xy,
                            const uint8_t* h_xy, size_t count, const uint8_t* proofs, size_t proof_stride, const uint8_t* comms_xy,
                            const uint8_t* transcript_state, const uint8_t* challenges_be, const uint8_t* rnd_key, size_t rnd_key_len,
                            int32_t* verdicts);
/* test hook: the same call (count <= 4096) that also returns the scalars the device built: count x (2N + 2) for
 * [G | H | g | h] and count x (6 + m + 5 + 2 lg) for [A_I1 A_O1 S1 A_I2 A_O2 S2 | V | T_1 T_3 T_4 T_5 T_6 | L | R] */
int bpgpu_r1cs_verify_batch_terms(bpgpu_ctx* ctx, const bpgpu_circuit* circuit, bpgpu_points* G, bpgpu_points* H, const uint8_t* g_xy,
                                  const uint8_t* h_xy, size_t count, const uint8_t* proofs, size_t proof_stride, const uint8_t* comms_xy,
                                  const uint8_t* transcript_state, const uint8_t* challenges_be, const uint8_t* rnd_key,
                                  size_t rnd_key_len, int32_t* verdicts, uint8_t* fixed_scalars_be, uint8_t* var_scalars_be);
/* window width the MSM would use for n terms (reporting only) */
int bpgpu_msm_window_bits(size_t n);

/* ---- fixed-base commitments ----------------------------------------------------------------------
 * Replaces commitment::commit_to_field_element(g, h, v, r) = g.binary_scalar_mul(h, v, r) and `g * w`
 * (prover.rs:123 V_i, :496-500 T_1..T_6, :550 Q = g*w; gadgets/poseidon_hash.rs:49-62) for bases that are reused:
 * create builds T[j][w][d] = d * 2^(4w) * B_j once (k bases, k <= 64); commit then evaluates `count` independent
 * combinations out_i = sum_j s[i*k + j] * B_j in one launch with no doublings.  bpgpu_fixed_bases_get returns a table
 * cached inside the ctx (built on first use; valid until the ctx is destroyed or 16 newer tables evict it). */
int bpgpu_fixed_bases_create(bpgpu_ctx* ctx, const uint8_t* bases_xy, size_t k, bpgpu_fixed_bases** out);
int bpgpu_fixed_bases_get(bpgpu_ctx* ctx, const uint8_t* bases_xy, size_t k, bpgpu_fixed_bases** out);
int bpgpu_fixed_bases_commit(bpgpu_ctx* ctx, bpgpu_fixed_bases* fb, const uint8_t* scalars_be, size_t count, uint8_t* out_xy);
void bpgpu_fixed_bases_free(bpgpu_fixed_bases* fb);

/* ---- FieldElementVector algebra on device-resident vectors ------------------------------------
 * Replaces FieldElementVector::{new_vandermonde_vector, hadamard_product, plus, scaled_by,
 * inner_product} and FieldElement::batch_invert (ipp.rs:77-82,94-95,145-146,295; prover.rs:463,
 * 472-485,513; verifier.rs:342-352,416) and VecPoly3::special_inner_product (vector_poly.rs:79-97).
 * Length/offset violations return BPGPU_E_LEN (ValueError::UnequalSizeVectors). */
int bpgpu_scalars_alloc(bpgpu_ctx* ctx, size_t n, bpgpu_scalars** out);                    /* zero vector */
int bpgpu_fr_vandermonde(bpgpu_ctx* ctx, const uint8_t* x_be, size_t n, bpgpu_scalars** out); /* [1,x,..,x^(n-1)] */
int bpgpu_fr_hadamard(bpgpu_ctx* ctx, const bpgpu_scalars* a, size_t aoff, const bpgpu_scalars* b, size_t boff,
                      size_t n, bpgpu_scalars* out, size_t ooff);
int bpgpu_fr_add(bpgpu_ctx* ctx, const bpgpu_scalars* a, size_t aoff, const bpgpu_scalars* b, size_t boff,
                 size_t n, bpgpu_scalars* out, size_t ooff);
int bpgpu_fr_sub(bpgpu_ctx* ctx, const bpgpu_scalars* a, size_t aoff, const bpgpu_scalars* b, size_t boff,
                 size_t n, bpgpu_scalars* out, size_t ooff);
int bpgpu_fr_scale(bpgpu_ctx* ctx, const bpgpu_scalars* a, size_t aoff, size_t n, const uint8_t* s_be,
                   bpgpu_scalars* out, size_t ooff);
int bpgpu_fr_inner_product(bpgpu_ctx* ctx, const bpgpu_scalars* a, size_t aoff, const bpgpu_scalars* b, size_t boff,
                           size_t n, uint8_t* out_be);
/* t1..t6 (6 x MODBYTES) from l1,l2,l3,r0,r1,r3: nine inner products in one pass */
int bpgpu_fr_poly3_special_inner_product(bpgpu_ctx* ctx, const bpgpu_scalars* l1, const bpgpu_scalars* l2,
                                         const bpgpu_scalars* l3, const bpgpu_scalars* r0, const bpgpu_scalars* r1,
                                         const bpgpu_scalars* r3, size_t n, uint8_t* t_be);
/* elementwise inverses (0 -> 0) into out, product of all inverses into prod_inv_be (may be NULL) */
int bpgpu_fr_batch_invert(bpgpu_ctx* ctx, const bpgpu_scalars* a, size_t aoff, size_t n, bpgpu_scalars* out,
                          size_t ooff, uint8_t* prod_inv_be);
/* FieldElementVector::random (the blinding vectors s_L, s_R: prover.rs:340-341, 401-402) generated on the device:
 * out[i] = be_int(SHAKE256(key || le64(ctr0 + i))[..MODBYTES]) mod r, i < n -- the counter-mode stream of the host
 * layer's Rng (host/curve.hpp), so a prover that draws a vector here gets exactly the scalars n host draws would give.
 * key_len <= 64.  The vector never exists on the host. */
int bpgpu_fr_random(bpgpu_ctx* ctx, const uint8_t* key, size_t key_len, uint64_t ctr0, size_t n, bpgpu_scalars** out);

/* ---- lock-step proving of a BATCH of independent one-phase R1CS proofs of one shape (n multipliers; csrc/provebatch.cu) ----
 * Every stage of Prover::prove (prover.rs:322-593) and every round of IPP::create_ipp (ipp.rs:68-194) is ONE device call
 * for all `batch` proofs; the caller runs the transcripts in between (host/prove_batch.hpp is that caller).  G and H get
 * window tables (bpgpu_points_precompute); g, h are the Pedersen pair.  All scalar arguments are big endian, MODBYTES
 * each, proof-major:
 *   commit3   : witness = batch x [a_L | a_R | a_O] (3n), keys = batch x key_len bytes and ctr0[batch] = where each proof's
 *               blinding stream stands (s_L, s_R = the next 2n draws, made on the device as by bpgpu_fr_random),
 *               blind = batch x [i_b, o_b, s_b]; out = batch x [A_I, A_O, S]
 *   polys     : weights = batch x [wL | wR | wO] (3n), y = batch x [y, y^-1]; out t = batch x [t_1 .. t_6]
 *   eval      : xuw = batch x [x, u, w] -> l_vec, r_vec, G_factors, H_factors become the IPP state (Q = w * g)
 *   ipp_round : uv = batch x [u, u^-1] of the previous round (NULL for the first); out = batch x [L, R]
 *   ipp_finish: uv of the last round; out = batch x [a, b] */
typedef struct bpgpu_pbatch bpgpu_pbatch;
int bpgpu_pbatch_create(bpgpu_ctx* ctx, bpgpu_points* G, bpgpu_points* H, const uint8_t* g_xy, const uint8_t* h_xy, size_t batch, size_t n,
                        bpgpu_pbatch** out);
void bpgpu_pbatch_free(bpgpu_pbatch* pb);
int bpgpu_pbatch_commit3(bpgpu_pbatch* pb, const uint8_t* witness_be, const uint8_t* keys, size_t key_len, const uint64_t* ctr0,
                         const uint8_t* blind_be, uint8_t* out_xy);
int bpgpu_pbatch_polys(bpgpu_pbatch* pb, const uint8_t* weights_be, const uint8_t* y_be, uint8_t* t_be);
int bpgpu_pbatch_eval(bpgpu_pbatch* pb, const uint8_t* xuw_be);
int bpgpu_pbatch_ipp_round(bpgpu_pbatch* pb, const uint8_t* uv_be, uint8_t* out_xy);
int bpgpu_pbatch_ipp_finish(bpgpu_pbatch* pb, const uint8_t* uv_be, uint8_t* ab_be);

/* The whole slab in ONE call, transcripts on the device (SURVEY.md section 8 f3 for slabs): `batch` range proofs of m values x
 * `bits` bits each (m * bits = the pbatch's n; `circuit` = the same statement's constraint matrix, bph_range_circuit_csr).
 * values = batch x m u64; transcript_state = the exported Merlin state after Transcript::new(label) + r1cs_domain_sep()
 * (203 bytes); keys = batch x key_len bytes, proof i's blinding stream SHAKE256(key_i || le64(counter)) consumed in the
 * reference's draw order.  The device derives the witness bits (positive_no.rs:18-24), draws every
 * blinding, replays the transcripts (one thread per proof), flattens the constraints from each proof's z, normalises every
 * commitment (one inversion per 4 points) and assembles the flat proof records: proofs = batch x proof_stride bytes,
 * comms_xy = batch x m commitments.  Proof i is byte-identical to gen_proof_of_positive_nums with the same stream. */
int bpgpu_pbatch_prove_range(bpgpu_pbatch* pb, const bpgpu_circuit* circuit, const uint64_t* values, size_t m, size_t bits,
                             const uint8_t* transcript_state, const uint8_t* keys, size_t key_len, uint8_t* proofs,
                             size_t proof_stride, uint8_t* comms_xy);

/* ---- inner-product argument, device-resident across rounds (IPP::create_ipp, ipp.rs:35-202) ----
 * begin: clones G[goff..goff+n), H[hoff..), a, b and the factor vectors (ipp.rs:57-60); n must be a power
 *        of two (BPGPU_E_NOT_POW2 = the assert at ipp.rs:48) and all vectors length n (BPGPU_E_LEN).
 * round_LR: L and R of the current round (ipp.rs:80-104 / 148-170), as X||Y.  The host commits them to
 *        its transcript, draws u (ipp.rs:106-113 / 172-179) and calls
 * fold:  a, b and the generator coefficients are folded by (u, u^-1) (ipp.rs:115-130 / 181-188).
 * finish: after lg n rounds, the proof's a and b (ipp.rs:196-201).
 * The G_factors / H_factors are applied in the first round only, as in the reference (ipp.rs:74-129). */
int bpgpu_ipp_begin(bpgpu_ctx* ctx, const bpgpu_points* G, size_t goff, const bpgpu_points* H, size_t hoff,
                    const uint8_t* Q_xy, const bpgpu_scalars* G_factors, const bpgpu_scalars* H_factors,
                    const bpgpu_scalars* a, const bpgpu_scalars* b, size_t n, bpgpu_ipp** out);
/* the same with Q given as q_scalar * q_base for a fixed base (the R1CS prover's Q = g * w, prover.rs:549-550): with
 * precomputed G and H every round is then table-only */
int bpgpu_ipp_begin_fixed_q(bpgpu_ctx* ctx, const bpgpu_points* G, size_t goff, const bpgpu_points* H, size_t hoff,
                            const uint8_t* q_base_xy, const uint8_t* q_scalar_be, const bpgpu_scalars* G_factors,
                            const bpgpu_scalars* H_factors, const bpgpu_scalars* a, const bpgpu_scalars* b, size_t n, bpgpu_ipp** out);
size_t bpgpu_ipp_len(const bpgpu_ipp* ipp);           /* current vector length (n, n/2, ..., 1) */
int bpgpu_ipp_round_LR(bpgpu_ipp* ipp, uint8_t* L_xy, uint8_t* R_xy);
int bpgpu_ipp_fold(bpgpu_ipp* ipp, const uint8_t* u_be, const uint8_t* u_inv_be);
int bpgpu_ipp_finish(bpgpu_ipp* ipp, uint8_t* a_be, uint8_t* b_be);
void bpgpu_ipp_free(bpgpu_ipp* ipp);
/* the s vector of IPP::verification_scalars (ipp.rs:303-312) from the lg challenges u_k (creation order);
 * lg >= 32 -> BPGPU_E_VERIFY (ipp.rs:269-273).  The transcript replay that yields u_k stays on the host. */
int bpgpu_ipp_verification_scalars(bpgpu_ctx* ctx, const uint8_t* u_be, size_t lg, bpgpu_scalars** s_out);
/* expected_P of IPP::verify_ipp (ipp.rs:220-253): one MSM over [Q | G | H | L | R] with the scalars
 * [a*b | a*s_i*Gf_i | b*s_(n-1-i)*Hf_i | -u_k^2 | -u_k^-2] built on the device; n = 2^lg. */
int bpgpu_ipp_verify_msm(bpgpu_ctx* ctx, const bpgpu_points* G, size_t goff, const bpgpu_points* H, size_t hoff,
                         const uint8_t* Q_xy, const bpgpu_scalars* G_factors, const bpgpu_scalars* H_factors,
                         const uint8_t* a_be, const uint8_t* b_be, const uint8_t* u_be, const uint8_t* L_xy,
                         const uint8_t* R_xy, size_t lg, uint8_t* out_xy);

/* ---- fused Fr kernels of the R1CS prover / verifier ------------------------------------------------
 * prover_polys (prover.rs:469-486): from a_L, a_R, s_R and the flattened weights wL, wR, wO (length n) and
 *   the challenge y: l1 = a_L + y^-i*wR, r0 = wO - y^i, r1 = y^i*a_R + wL, r3 = y^i*s_R (l2 = a_O, l3 = s_L).
 * prover_eval (prover.rs:524-535,552-563): l_vec = l(x) | 0^pad, r_vec = r(x) | (-y^i)_{i>=n},
 *   G_factors = 1^n1 | u^(N-n1), H_factors = y^-i * G_factors, all of length padded_n, ready for bpgpu_ipp_begin.
 * verifier_scalars (verifier.rs:341-390): g_scalars | h_scalars (2*padded_n) and delta from wL, wR, wO (length n),
 *   the s vector (bpgpu_ipp_verification_scalars) and the challenges / proof scalars y, x, a, b, u. */
int bpgpu_r1cs_prover_polys(bpgpu_ctx* ctx, size_t n, const bpgpu_scalars* a_L, const bpgpu_scalars* a_R,
                            const bpgpu_scalars* s_R, const bpgpu_scalars* wL, const bpgpu_scalars* wR,
                            const bpgpu_scalars* wO, const uint8_t* y_be, bpgpu_scalars** l1, bpgpu_scalars** r0,
                            bpgpu_scalars** r1, bpgpu_scalars** r3);
int bpgpu_r1cs_prover_eval(bpgpu_ctx* ctx, size_t n, size_t n1, size_t padded_n, const bpgpu_scalars* l1,
                           const bpgpu_scalars* l2, const bpgpu_scalars* l3, const bpgpu_scalars* r0,
                           const bpgpu_scalars* r1, const bpgpu_scalars* r3, const uint8_t* x_be, const uint8_t* u_be,
                           const uint8_t* y_be, bpgpu_scalars** l_vec, bpgpu_scalars** r_vec,
                           bpgpu_scalars** G_factors, bpgpu_scalars** H_factors);
int bpgpu_r1cs_verifier_scalars(bpgpu_ctx* ctx, size_t n, size_t n1, size_t padded_n, const bpgpu_scalars* wL,
                                const bpgpu_scalars* wR, const bpgpu_scalars* wO, const bpgpu_scalars* s,
                                const uint8_t* y_be, const uint8_t* x_be, const uint8_t* a_be, const uint8_t* b_be,
                                const uint8_t* u_be, bpgpu_scalars** gh_scalars, uint8_t* delta_be);

/* ---- witness generation for the MiMC hash gadget (gadgets/helper_constraints/mimc.rs:10-29,53-77) ----
 * `count` independent instances of mimc(xl, xr, constants, rounds): per round xl, xr := xr + (xl + c_i)^3, xl.
 * image[i] = the hash of instance i; a_L / a_R / a_O (optional, all three or none) receive the multiplier assignments of
 * enforce_mimc_2_inputs, instance-major, 2 * rounds multipliers each: (l, l, l^2) and (l^2, l, l^3) with l = xl + c_i --
 * device-resident, ready for the commitment MSMs.  constants must hold at least `rounds` scalars (mimc.rs:18). */
int bpgpu_mimc_witness(bpgpu_ctx* ctx, const bpgpu_scalars* xl, const bpgpu_scalars* xr, size_t count, const bpgpu_scalars* constants,
                       size_t rounds, bpgpu_scalars** image, bpgpu_scalars** a_L, bpgpu_scalars** a_R, bpgpu_scalars** a_O);

/* Poseidon_permutation (gadgets/helper_constraints/poseidon.rs:202-293) for `count` independent states of `width` (3, 5
 * or 9) scalars: full_rounds_beginning full rounds, partial_rounds partial rounds (S-box on the last element only),
 * full_rounds_end full rounds; round_keys = (total rounds) x width scalars consumed in order, mds = width x width row-major
 * (new[i] = sum_j state[j] * mds[j][i]); sbox: 0 cube, 1 inverse, 2 quint (poseidon.rs:122-138).  The round keys and the
 * matrix are the caller's parameters (the reference's tables: gadgets/poseidon_constants.rs).  out = count x width. */
int bpgpu_poseidon_permutation(bpgpu_ctx* ctx, const bpgpu_scalars* input, size_t count, size_t width, size_t full_rounds_beginning,
                               size_t partial_rounds, size_t full_rounds_end, int sbox, const bpgpu_scalars* round_keys,
                               const bpgpu_scalars* mds, bpgpu_scalars** out);

/* ---- self-test / measurement hooks (used by tests/ and bench.py; not part of the drop-in) ---- */
/* field: 0 Fq, 1 Fr of the ctx curve; op: 0 mul 1 add 2 sub 3 inv 4 sqr; operands are canonical
 * big-endian MODBYTES values (converted to/from Montgomery form on the device). */
int bpgpu_selftest_field(bpgpu_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b, size_t n,
                         uint8_t* out);
/* group ops on the device: op 0: P+Q, 1: 2P, 2: k*P (k = scalars_be[i]) ; points as xy */
int bpgpu_selftest_group(bpgpu_ctx* ctx, int op, const uint8_t* p_xy, const uint8_t* q_xy,
                         const uint8_t* scalars_be, size_t n, uint8_t* out_xy);
/* integer-pipe microbenchmark: returns achieved 32x32+64 multiply-adds per second through
 * IMAD.WIDE.U32 chains (kind 0) or Fq Montgomery products per second (kind 1). */
int bpgpu_int_pipe_bench(bpgpu_ctx* ctx, int kind, int iters, double* ops_per_s, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* BPGPU_H */
