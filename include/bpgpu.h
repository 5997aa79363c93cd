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

const char* bpgpu_strerror(int code);
int bpgpu_modbytes(int curve);              /* amcl_wrapper::constants::MODBYTES */
int bpgpu_device_count(void);

int bpgpu_ctx_create(int curve, int device, bpgpu_ctx** out);
void bpgpu_ctx_destroy(bpgpu_ctx* ctx);
void* bpgpu_ctx_stream(bpgpu_ctx* ctx);     /* the cudaStream_t all work of this ctx is issued on */
int bpgpu_ctx_sync(bpgpu_ctx* ctx);
int bpgpu_ctx_curve(const bpgpu_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
uint64_t bpgpu_ctx_launches(const bpgpu_ctx* ctx);

/* per-stage CUDA-event timing of the MSM pipeline on the ctx stream (roofline evidence for bench.py).
 * stages: 0 digits, 1 scan, 2 scatter, 3 chunk_acc, 4 giant, 5 reduce_l1, 6 reduce_l2.
 * bpgpu_msm_stage_ms writes the average ms per stage since set_profile(1) and returns the run count. */
int bpgpu_ctx_set_profile(bpgpu_ctx* ctx, int on);
int bpgpu_msm_stage_ms(const bpgpu_ctx* ctx, double* avg_ms, int cap);

void* bpgpu_host_alloc(size_t bytes);       /* pinned host memory */
void bpgpu_host_free(void* p);

/* ---- G1Vector / FieldElementVector residency (G1Vector::from / FieldElementVector::from) ---- */
int bpgpu_points_upload(bpgpu_ctx* ctx, const uint8_t* xy, size_t n, bpgpu_points** out);
int bpgpu_points_download(bpgpu_ctx* ctx, const bpgpu_points* p, size_t off, size_t n, uint8_t* xy);
size_t bpgpu_points_len(const bpgpu_points* p);
void bpgpu_points_free(bpgpu_points* p);
int bpgpu_scalars_upload(bpgpu_ctx* ctx, const uint8_t* be, size_t n, bpgpu_scalars** out);
int bpgpu_scalars_download(bpgpu_ctx* ctx, const bpgpu_scalars* s, size_t off, size_t n, uint8_t* be);
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
/* cached bases and device-resident scalars */
int bpgpu_msm_device(bpgpu_ctx* ctx, const bpgpu_points* p, size_t poff, size_t n, const bpgpu_scalars* s,
                     size_t soff, uint8_t* out_xy);
/* ad-hoc host point and scalar lists (`Vec<&G1>`, `Vec<&FieldElement>`) */
int bpgpu_msm_refs(bpgpu_ctx* ctx, const uint8_t* points_xy, const uint8_t* scalars_be, size_t n,
                   uint8_t* out_xy);
/* window width the MSM would use for n terms (reporting only) */
int bpgpu_msm_window_bits(size_t n);

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
