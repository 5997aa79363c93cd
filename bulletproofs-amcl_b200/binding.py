"""ctypes binding of include/bpgpu.h.  Byte conventions are the ABI's: scalars are
FieldElement::to_bytes() (MODBYTES big endian), points are G1::to_bytes()[1:] (X||Y)."""
import ctypes
import os
import subprocess
import weakref

BLS12_381 = 0
BN254 = 1

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# (name, restype, argtypes) for every symbol include/bpgpu.h declares
_c = ctypes
_VP, _SZ, _U8P, _INT = _c.c_void_p, _c.c_size_t, _c.c_char_p, _c.c_int
SYMBOLS = [
    ("bpgpu_strerror", _c.c_char_p, [_INT]),
    ("bpgpu_modbytes", _INT, [_INT]),
    ("bpgpu_device_count", _INT, []),
    ("bpgpu_ctx_create", _INT, [_INT, _INT, _c.POINTER(_VP)]),
    ("bpgpu_ctx_destroy", None, [_VP]),
    ("bpgpu_ctx_aux", _INT, [_VP, _c.POINTER(_VP)]),
    ("bpgpu_ctx_stream", _VP, [_VP]),
    ("bpgpu_ctx_sync", _INT, [_VP]),
    ("bpgpu_ctx_curve", _INT, [_VP]),
    ("bpgpu_ctx_device", _INT, [_VP]),
    ("bpgpu_ctx_launches", _c.c_uint64, [_VP]),
    ("bpgpu_ctx_set_profile", _INT, [_VP, _INT]),
    ("bpgpu_ctx_set_fixed_schedule", _INT, [_VP, _INT]),
    ("bpgpu_ctx_set_blocking_sync", _INT, [_VP, _INT]),
    ("bpgpu_msm_stage_ms", _INT, [_VP, _c.POINTER(_c.c_double), _INT]),
    ("bpgpu_host_alloc", _VP, [_SZ]),
    ("bpgpu_host_free", None, [_VP]),
    ("bpgpu_points_upload", _INT, [_VP, _VP, _SZ, _c.POINTER(_VP)]),
    ("bpgpu_points_download", _INT, [_VP, _VP, _SZ, _SZ, _VP]),
    ("bpgpu_points_from_hashes", _INT, [_VP, _VP, _SZ, _c.POINTER(_VP)]),
    ("bpgpu_points_precompute", _INT, [_VP, _VP]),
    ("bpgpu_points_precompute_wide", _INT, [_VP, _VP]),
    ("bpgpu_points_has_tables", _INT, [_VP]),
    ("bpgpu_points_len", _SZ, [_VP]),
    ("bpgpu_points_free", None, [_VP]),
    ("bpgpu_scalars_upload", _INT, [_VP, _VP, _SZ, _c.POINTER(_VP)]),
    ("bpgpu_scalars_download", _INT, [_VP, _VP, _SZ, _SZ, _VP]),
    ("bpgpu_scalars_len", _SZ, [_VP]),
    ("bpgpu_scalars_free", None, [_VP]),
    ("bpgpu_msm", _INT, [_VP, _VP, _SZ, _SZ, _VP, _VP]),
    ("bpgpu_msm_le32", _INT, [_VP, _VP, _SZ, _SZ, _VP, _VP]),
    ("bpgpu_msm_begin", _INT, [_VP, _VP, _SZ, _SZ, _VP]),
    ("bpgpu_msm_le32_begin", _INT, [_VP, _VP, _SZ, _SZ, _VP]),
    ("bpgpu_msm_device_begin", _INT, [_VP, _VP, _SZ, _SZ, _VP, _SZ]),
    ("bpgpu_msm_refs_begin", _INT, [_VP, _VP, _VP, _SZ]),
    ("bpgpu_msm_finish", _INT, [_VP, _VP]),
    ("bpgpu_msm_device", _INT, [_VP, _VP, _SZ, _SZ, _VP, _SZ, _VP]),
    ("bpgpu_msm_refs", _INT, [_VP, _VP, _VP, _SZ, _VP]),
    ("bpgpu_msm_parts_batch", _INT, [_VP, _VP, _VP, _SZ, _VP]),
    ("bpgpu_msm_batch_is_identity", _INT, [_VP, _VP, _SZ, _SZ, _VP, _VP, _VP, _SZ, _VP]),
    ("bpgpu_msm_window_bits", _INT, [_SZ]),
    ("bpgpu_circuit_create", _INT, [_VP, _SZ, _SZ, _SZ, _VP, _VP, _VP, _c.POINTER(_VP)]),
    ("bpgpu_circuit_free", None, [_VP]),
    ("bpgpu_circuit_multipliers", _SZ, [_VP]),
    ("bpgpu_circuit_commitments", _SZ, [_VP]),
    ("bpgpu_circuit_flatten", _INT, [_VP, _VP, _VP, _c.POINTER(_VP)]),
    ("bpgpu_ctx_circuit_put", _INT, [_VP, _c.c_uint64, _VP]),
    ("bpgpu_ctx_circuit_get", _VP, [_VP, _c.c_uint64]),
    ("bpgpu_range_witness", _INT, [_VP, _c.POINTER(_c.c_uint64), _SZ, _SZ, _c.POINTER(_VP)]),
    ("bpgpu_r1cs_verify_batch", _INT, [_VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP, _SZ, _VP, _VP, _VP, _VP, _SZ, _VP]),
    ("bpgpu_r1cs_verify_batch_terms", _INT, [_VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP, _SZ, _VP, _VP, _VP, _VP, _SZ, _VP, _VP, _VP]),
    ("bpgpu_pbatch_create", _INT, [_VP, _VP, _VP, _VP, _VP, _SZ, _SZ, _c.POINTER(_VP)]),
    ("bpgpu_pbatch_free", None, [_VP]),
    ("bpgpu_pbatch_prove_range", _INT, [_VP, _VP, _VP, _SZ, _SZ, _VP, _VP, _SZ, _VP, _SZ, _VP]),
    ("bpgpu_pbatch_commit3", _INT, [_VP, _VP, _VP, _SZ, _VP, _VP, _VP]),
    ("bpgpu_pbatch_polys", _INT, [_VP, _VP, _VP, _VP]),
    ("bpgpu_pbatch_eval", _INT, [_VP, _VP]),
    ("bpgpu_pbatch_ipp_round", _INT, [_VP, _VP, _VP]),
    ("bpgpu_pbatch_ipp_finish", _INT, [_VP, _VP, _VP]),
    ("bpgpu_scalars_view", _INT, [_VP, _SZ, _SZ, _c.POINTER(_VP)]),
    ("bpgpu_fr_random", _INT, [_VP, _VP, _SZ, _c.c_uint64, _SZ, _c.POINTER(_VP)]),
    ("bpgpu_fixed_bases_create", _INT, [_VP, _VP, _SZ, _c.POINTER(_VP)]),
    ("bpgpu_fixed_bases_get", _INT, [_VP, _VP, _SZ, _c.POINTER(_VP)]),
    ("bpgpu_fixed_bases_commit", _INT, [_VP, _VP, _VP, _SZ, _VP]),
    ("bpgpu_fixed_bases_free", None, [_VP]),
    ("bpgpu_scalars_alloc", _INT, [_VP, _SZ, _c.POINTER(_VP)]),
    ("bpgpu_fr_vandermonde", _INT, [_VP, _VP, _SZ, _c.POINTER(_VP)]),
    ("bpgpu_fr_hadamard", _INT, [_VP, _VP, _SZ, _VP, _SZ, _SZ, _VP, _SZ]),
    ("bpgpu_fr_add", _INT, [_VP, _VP, _SZ, _VP, _SZ, _SZ, _VP, _SZ]),
    ("bpgpu_fr_sub", _INT, [_VP, _VP, _SZ, _VP, _SZ, _SZ, _VP, _SZ]),
    ("bpgpu_fr_scale", _INT, [_VP, _VP, _SZ, _SZ, _VP, _VP, _SZ]),
    ("bpgpu_fr_inner_product", _INT, [_VP, _VP, _SZ, _VP, _SZ, _SZ, _VP]),
    ("bpgpu_fr_poly3_special_inner_product", _INT, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    ("bpgpu_fr_batch_invert", _INT, [_VP, _VP, _SZ, _SZ, _VP, _SZ, _VP]),
    ("bpgpu_ipp_begin", _INT, [_VP, _VP, _SZ, _VP, _SZ, _VP, _VP, _VP, _VP, _VP, _SZ, _c.POINTER(_VP)]),
    ("bpgpu_ipp_begin_fixed_q", _INT, [_VP, _VP, _SZ, _VP, _SZ, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _c.POINTER(_VP)]),
    ("bpgpu_ipp_len", _SZ, [_VP]),
    ("bpgpu_ipp_round_LR", _INT, [_VP, _VP, _VP]),
    ("bpgpu_ipp_fold", _INT, [_VP, _VP, _VP]),
    ("bpgpu_ipp_finish", _INT, [_VP, _VP, _VP]),
    ("bpgpu_ipp_free", None, [_VP]),
    ("bpgpu_ipp_verification_scalars", _INT, [_VP, _VP, _SZ, _c.POINTER(_VP)]),
    ("bpgpu_ipp_verify_msm", _INT, [_VP, _VP, _SZ, _VP, _SZ, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    ("bpgpu_msm_parts", _INT, [_VP, _VP, _SZ, _VP]),
    ("bpgpu_r1cs_prover_polys", _INT, [_VP, _SZ, _VP, _VP, _VP, _VP, _VP, _VP, _VP] + [_c.POINTER(_VP)] * 4),
    ("bpgpu_r1cs_prover_eval", _INT, [_VP, _SZ, _SZ, _SZ] + [_VP] * 6 + [_VP, _VP, _VP] + [_c.POINTER(_VP)] * 4),
    ("bpgpu_r1cs_verifier_scalars", _INT, [_VP, _SZ, _SZ, _SZ] + [_VP] * 4 + [_VP] * 5 + [_c.POINTER(_VP), _VP]),
    ("bpgpu_mimc_witness", _INT, [_VP, _VP, _VP, _SZ, _VP, _SZ] + [_c.POINTER(_VP)] * 4),
    ("bpgpu_poseidon_permutation", _INT, [_VP, _VP, _SZ, _SZ, _SZ, _SZ, _SZ, _INT, _VP, _VP, _c.POINTER(_VP)]),
    ("bpgpu_selftest_field", _INT, [_VP, _INT, _INT, _VP, _VP, _SZ, _VP]),
    ("bpgpu_selftest_group", _INT, [_VP, _INT, _VP, _VP, _VP, _SZ, _VP]),
    ("bpgpu_int_pipe_bench", _INT, [_VP, _INT, _INT, _c.POINTER(_c.c_double), _c.POINTER(_c.c_double)]),
]

# include/bphost.h: the host layer (C++ mirror of the reference's prover / verifier API)
_CS, _U64 = _c.c_char_p, _c.c_uint64
SYMBOLS_HOST = [
    ("bph_merlin_kat", _INT, [_CS, _CS, _VP, _SZ, _CS, _VP, _SZ]),
    ("bph_transcript_kat", _INT, [_INT, _CS, _VP, _VP, _VP]),
    ("bph_get_generators", _INT, [_VP, _CS, _SZ, _c.POINTER(_VP)]),
    ("bph_g1_from_msg_hash", _INT, [_VP, _VP, _SZ, _VP]),
    ("bph_ipp_create", _INT, [_VP, _CS, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP, _SZ, _c.POINTER(_SZ)]),
    ("bph_ipp_verify", _INT, [_VP, _CS, _SZ, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ]),
    ("bph_bound_check_prove", _INT, [_VP, _CS, _VP, _VP, _VP, _VP, _U64, _VP, _U64, _U64, _SZ, _INT, _U64, _VP, _SZ,
                                     _c.POINTER(_SZ), _VP]),
    ("bph_bound_check_verify", _INT, [_VP, _CS, _VP, _VP, _VP, _VP, _U64, _U64, _SZ, _VP, _SZ, _VP, _VP]),
    ("bph_shuffle_prove", _INT, [_VP, _CS, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _SZ, _INT, _U64, _VP, _SZ, _c.POINTER(_SZ), _VP]),
    ("bph_shuffle_verify", _INT, [_VP, _CS, _VP, _VP, _VP, _VP, _SZ, _SZ, _VP, _SZ, _VP, _VP]),
    ("bph_range_prove", _INT, [_VP, _CS, _VP, _VP, _VP, _VP, _VP, _SZ, _SZ, _INT, _U64, _VP, _SZ, _c.POINTER(_SZ), _VP]),
    ("bph_range_verify", _INT, [_VP, _CS, _VP, _VP, _VP, _VP, _SZ, _SZ, _VP, _SZ, _VP, _VP]),
    ("bph_range_proof_len", _SZ, [_INT, _SZ, _SZ]),
    ("bph_range_prove_many", _INT, [_VP, _SZ, _CS, _VP, _VP, _VP, _VP, _VP, _SZ, _SZ, _SZ, _INT, _U64, _VP, _SZ, _VP]),
    ("bph_range_prove_batch", _INT, [_VP, _CS, _VP, _VP, _VP, _VP, _VP, _SZ, _SZ, _SZ, _INT, _U64, _SZ, _VP, _SZ, _VP]),
    ("bph_range_prove_batch_mode", _INT, [_VP, _CS, _VP, _VP, _VP, _VP, _VP, _SZ, _SZ, _SZ, _INT, _U64, _INT, _SZ, _VP, _SZ, _VP]),
    ("bph_range_verify_batch", _INT, [_VP, _CS, _VP, _VP, _VP, _VP, _SZ, _SZ, _SZ, _VP, _SZ, _VP, _SZ, _VP]),
    ("bph_range_verify_batch_mode", _INT, [_VP, _CS, _VP, _VP, _VP, _VP, _SZ, _SZ, _SZ, _VP, _SZ, _VP, _INT, _SZ, _VP]),
    ("bph_bound_check_verify_batch", _INT, [_VP, _CS, _VP, _VP, _VP, _VP, _SZ, _U64, _U64, _SZ, _VP, _SZ, _VP, _INT, _SZ, _VP]),
    ("bph_range_circuit_csr", _INT, [_INT, _SZ, _SZ] + [_c.POINTER(_SZ)] * 4 + [_VP, _VP, _VP]),
    ("bph_bound_check_circuit_csr", _INT, [_INT, _U64, _U64, _SZ] + [_c.POINTER(_SZ)] * 4 + [_VP, _VP, _VP]),
    ("bph_set_secret_fixed_schedule", None, [_INT]),
    ("bph_r1cs_transcript_state", None, [_CS, _VP]),
    ("bph_r1cs_replay_challenges", _INT, [_INT, _CS, _VP, _VP, _SZ, _SZ, _VP]),
    ("bph_msm_sharded", _INT, [_VP, _SZ, _VP, _VP, _VP, _VP]),
    ("bph_g1_sum", _INT, [_INT, _VP, _SZ, _VP]),
    ("bph_range_verify_many", _INT, [_VP, _SZ, _CS, _VP, _VP, _VP, _VP, _SZ, _SZ, _SZ, _VP, _SZ, _VP, _VP]),
]


class FixedRun(ctypes.Structure):
    """bpgpu_fixed_run"""
    _fields_ = [("points", _VP), ("off", _SZ), ("n", _SZ), ("host_base_xy", _VP)]


class MsmPart(ctypes.Structure):
    """bpgpu_msm_part"""
    _fields_ = [("points", _VP), ("points_off", _SZ), ("host_points_xy", _VP),
                ("scalars", _VP), ("scalars_off", _SZ), ("host_scalars_be", _VP), ("n", _SZ)]


def _ctx_array(ctxs):
    return (ctypes.c_void_p * len(ctxs))(*[c.handle for c in ctxs])


def range_prove_many(ctxs, label, g_xy, h_xy, G, H, values, m, bits, seed=None):
    """count = len(values)/m independent range proofs over len(ctxs) contexts (one host thread each).
    Returns (proofs bytes, stride, commitments bytes)."""
    count = len(values) // m
    c0 = ctxs[0]
    stride = lib().bph_range_proof_len(c0.curve, m, bits)
    proofs = ctypes.create_string_buffer(max(1, count * stride))
    comms = ctypes.create_string_buffer(max(1, count * m * 2 * c0.modbytes))
    arr = (ctypes.c_uint64 * max(1, len(values)))(*values)
    rc = lib().bph_range_prove_many(_ctx_array(ctxs), len(ctxs), label, _buf(g_xy), _buf(h_xy), G.handle, H.handle,
                                    ctypes.cast(arr, ctypes.c_void_p), count, m, bits, 0 if seed is None else 1, seed or 0, proofs, stride,
                                    comms)
    if rc:
        raise BpgpuError(rc, "range_prove_many")
    return proofs.raw[:count * stride], stride, comms.raw[:count * m * 2 * c0.modbytes]


def range_prove_batch(ctx, label, g_xy, h_xy, G, H, values, m, bits, seed=None, nthreads=0, mode=0):
    """the same proofs as range_prove_many, proved in lock-step on ONE context (one device call per prover stage and IPP
    round for the whole batch; mode 0: transcripts on the device too, mode 1: on host threads).
    Returns (proofs bytes, stride, commitments bytes)."""
    count = len(values) // m
    stride = lib().bph_range_proof_len(ctx.curve, m, bits)
    proofs = ctypes.create_string_buffer(max(1, count * stride))
    comms = ctypes.create_string_buffer(max(1, count * m * 2 * ctx.modbytes))
    arr = (ctypes.c_uint64 * max(1, len(values)))(*values)
    rc = lib().bph_range_prove_batch_mode(ctx.handle, label, _buf(g_xy), _buf(h_xy), G.handle, H.handle, ctypes.cast(arr, ctypes.c_void_p),
                                          count, m, bits, 0 if seed is None else 1, seed or 0, mode, nthreads, proofs, stride, comms)
    if rc:
        raise BpgpuError(rc, "range_prove_batch")
    return proofs.raw[:count * stride], stride, comms.raw[:count * m * 2 * ctx.modbytes]


def range_verify_many(ctxs, label, g_xy, h_xy, G, H, count, m, bits, proofs, stride, comms):
    """per-proof verdicts (0 = Ok, -4 = VerificationError, ...) of `count` independent proofs."""
    verdicts = (ctypes.c_int32 * max(1, count))()
    rc = lib().bph_range_verify_many(_ctx_array(ctxs), len(ctxs), label, _buf(g_xy), _buf(h_xy), G.handle, H.handle, count, m, bits,
                                     _buf(proofs), stride, _buf(comms), verdicts)
    if rc:
        raise BpgpuError(rc, "range_verify_many")
    return list(verdicts)[:count]


def msm_sharded(ctxs, shards, scalars_be):
    """sum over all shards of <shard points, their scalars>: shards[k] is a DevicePoints resident on ctxs[k]'s device, the
    scalars follow the shards in order.  One host thread per context; the partial sums are combined on the host."""
    n = len(ctxs)
    mb = ctxs[0].modbytes
    arr_c = (ctypes.c_void_p * n)(*[c.handle for c in ctxs])
    arr_p = (ctypes.c_void_p * n)(*[s.handle for s in shards])
    arr_n = (ctypes.c_size_t * n)(*[len(s) for s in shards])
    out = ctypes.create_string_buffer(2 * mb)
    rc = lib().bph_msm_sharded(arr_c, n, arr_p, arr_n, _buf(scalars_be), out)
    if rc:
        raise BpgpuError(rc, "msm_sharded")
    return out.raw


def g1_sum(curve, points_xy):
    """host-side sum of affine points X||Y (the combine step of a sharded MSM)"""
    mb = 48 if curve == BLS12_381 else 32
    out = ctypes.create_string_buffer(2 * mb)
    rc = lib().bph_g1_sum(curve, _buf(points_xy), len(points_xy) // (2 * mb), out)
    if rc:
        raise BpgpuError(rc, "g1_sum")
    return out.raw


def range_verify_batch(ctx, label, g_xy, h_xy, G, H, count, m, bits, proofs, stride, comms, nthreads=0, mode=0):
    """per-proof verdicts of `count` independent proofs from batched device calls (bph_range_verify_batch_mode):
    mode 0 = transcripts and scalars on the device, 1 = transcripts on host threads, 2 = round 1's host-built scalars."""
    verdicts = (ctypes.c_int32 * max(1, count))()
    rc = lib().bph_range_verify_batch_mode(ctx.handle, label, _buf(g_xy), _buf(h_xy), G.handle, H.handle, count, m, bits, _buf(proofs), stride,
                                           _buf(comms), mode, nthreads, verdicts)
    if rc:
        raise BpgpuError(rc, "range_verify_batch")
    return list(verdicts)[:count]


def bound_check_verify_batch(ctx, label, g_xy, h_xy, G, H, count, lower, upper, bits, proofs, stride, comms, nthreads=0, mode=0):
    """verify_proof_of_bounded_num for `count` proofs over the same bounds, one batched device call per slab"""
    verdicts = (ctypes.c_int32 * max(1, count))()
    rc = lib().bph_bound_check_verify_batch(ctx.handle, label, _buf(g_xy), _buf(h_xy), G.handle, H.handle, count, lower, upper, bits,
                                            _buf(proofs), stride, _buf(comms), mode, nthreads, verdicts)
    if rc:
        raise BpgpuError(rc, "bound_check_verify_batch")
    return list(verdicts)[:count]


def _csr_call(fn, *args):
    n, m, q, nnz = (ctypes.c_size_t() for _ in range(4))
    refs = [ctypes.byref(x) for x in (n, m, q, nnz)]
    rc = fn(*args, *refs, None, None, None)
    if rc:
        raise BpgpuError(rc, "circuit_csr")
    mb = 48 if args[0] == BLS12_381 else 32
    rows = (ctypes.c_uint32 * (3 * n.value + m.value + 2))()
    eq = (ctypes.c_uint32 * max(1, nnz.value))()
    ec = ctypes.create_string_buffer(max(1, nnz.value * mb))
    rc = fn(*args, *refs, ctypes.cast(rows, ctypes.c_void_p), ctypes.cast(eq, ctypes.c_void_p), ctypes.cast(ec, ctypes.c_void_p))
    if rc:
        raise BpgpuError(rc, "circuit_csr")
    return {"n": n.value, "m": m.value, "q": q.value, "row_start": list(rows), "ent_q": list(eq)[:nnz.value], "ent_c_be": ec.raw[:nnz.value * mb]}


def range_circuit_csr(curve, m, bits):
    """the flattened-constraints matrix of m x positive_no_gadget(bits) as CSR arrays (host only)"""
    return _csr_call(lib().bph_range_circuit_csr, curve, m, bits)


def bound_check_circuit_csr(curve, lower, upper, bits):
    return _csr_call(lib().bph_bound_check_circuit_csr, curve, lower, upper, bits)


def r1cs_transcript_state(label):
    """exported Merlin state after Transcript::new(label) + r1cs_domain_sep() (203 bytes)"""
    out = ctypes.create_string_buffer(203)
    lib().bph_r1cs_transcript_state(label, out)
    return out.raw


def r1cs_replay_challenges(curve, label, proof, comms, m, lg):
    """y, z, u, x, w, u_1..u_lg of one proof from a host transcript (big-endian scalars)"""
    mb = 48 if curve == BLS12_381 else 32
    out = ctypes.create_string_buffer((5 + lg) * mb)
    rc = lib().bph_r1cs_replay_challenges(curve, label, _buf(proof), _buf(comms), m, lg, out)
    if rc:
        raise BpgpuError(rc, "r1cs_replay_challenges")
    return out.raw


class Circuit:
    """bpgpu_circuit: a one-phase constraint system as a device-resident sparse matrix"""

    def __init__(self, ctx, csr):
        self.ctx, self.csr = ctx, csr
        rows = (ctypes.c_uint32 * len(csr["row_start"]))(*csr["row_start"])
        eq = (ctypes.c_uint32 * max(1, len(csr["ent_q"])))(*csr["ent_q"])
        h = ctypes.c_void_p()
        ctx._check(lib().bpgpu_circuit_create(ctx.handle, csr["n"], csr["m"], csr["q"], ctypes.cast(rows, ctypes.c_void_p),
                                              ctypes.cast(eq, ctypes.c_void_p), _buf(csr["ent_c_be"]), ctypes.byref(h)), "circuit_create")
        self.handle = h
        ctx._track(self)

    def flatten(self, z_be):
        """[wL | wR | wO | wV | wc] for the challenge z, as a DeviceScalars of 3n + m + 1 entries"""
        h = ctypes.c_void_p()
        self.ctx._check(lib().bpgpu_circuit_flatten(self.ctx.handle, self.handle, _buf(z_be), ctypes.byref(h)), "circuit_flatten")
        return DeviceScalars(self.ctx, h)

    def verify_batch(self, G, H, g_xy, h_xy, count, proofs, stride, comms, state=None, challenges=None, key=b"", terms=False):
        """bpgpu_r1cs_verify_batch(_terms): verdict list (and, with terms=True, the fixed and variable scalars the device built)"""
        verdicts = (ctypes.c_int32 * max(1, count))()
        st = _buf(state) if state is not None else None
        ch = _buf(challenges) if challenges is not None else None
        if not terms:
            self.ctx._check(lib().bpgpu_r1cs_verify_batch(self.ctx.handle, self.handle, G.handle, H.handle, _buf(g_xy), _buf(h_xy), count,
                                                          _buf(proofs), stride, _buf(comms), st, ch, _buf(key), len(key), verdicts),
                            "r1cs_verify_batch")
            return list(verdicts)[:count]
        n, m, mb = self.csr["n"], self.csr["m"], self.ctx.modbytes
        N = 1
        while N < n:
            N <<= 1
        lg = N.bit_length() - 1
        F, vn = 2 * N + 2, 6 + m + 5 + 2 * lg
        fixed = ctypes.create_string_buffer(max(1, count * F * mb))
        var = ctypes.create_string_buffer(max(1, count * vn * mb))
        self.ctx._check(lib().bpgpu_r1cs_verify_batch_terms(self.ctx.handle, self.handle, G.handle, H.handle, _buf(g_xy), _buf(h_xy), count,
                                                            _buf(proofs), stride, _buf(comms), st, ch, _buf(key), len(key), verdicts, fixed,
                                                            var), "r1cs_verify_batch_terms")
        return list(verdicts)[:count], fixed.raw[:count * F * mb], var.raw[:count * vn * mb]

    def free(self):
        if self.handle:
            lib().bpgpu_circuit_free(self.handle)
            self.handle = None


class BpgpuError(RuntimeError):
    def __init__(self, code, what=""):
        self.code = code
        msg = lib().bpgpu_strerror(code).decode() if _LIB is not None else str(code)
        super().__init__(f"{what}: {msg} ({code})")


def library_path():
    return os.path.join(_HERE, "libbpgpu.so")


def build_library(verbose=False):
    """Compile csrc/ into libbpgpu.so for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j4"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("building libbpgpu.so failed")
    return library_path()


def lib():
    """Load libbpgpu.so; raises loudly if it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        L = ctypes.CDLL(path)
        for name, res, args in SYMBOLS + SYMBOLS_HOST:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _buf(b):
    """bytes/bytearray/memoryview/int address -> something ctypes accepts as void*."""
    if isinstance(b, int):
        return ctypes.c_void_p(b)
    if isinstance(b, (bytes, bytearray)):
        return ctypes.cast(ctypes.c_char_p(bytes(b)) if isinstance(b, bytearray) else ctypes.c_char_p(b), ctypes.c_void_p)
    return ctypes.cast(ctypes.c_char_p(bytes(b)), ctypes.c_void_p)


class DevicePoints:
    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle
        ctx._track(self)

    def __len__(self):
        return lib().bpgpu_points_len(self.handle)

    def precompute(self):
        """build the window tables (bpgpu_points_precompute): later MSMs over this table need no doublings"""
        self.ctx._check(lib().bpgpu_points_precompute(self.ctx.handle, self.handle), "points_precompute")
        return self

    def precompute_wide(self):
        """16-bit window tables on top (bpgpu_points_precompute_wide): the batch calls add half as many entries per term"""
        self.ctx._check(lib().bpgpu_points_precompute_wide(self.ctx.handle, self.handle), "points_precompute_wide")
        return self

    @property
    def has_tables(self):
        return bool(lib().bpgpu_points_has_tables(self.handle))

    @property
    def table_level(self):
        """0: no tables, 1: 8-bit window tables, 2: also the wide (16-bit window) tables"""
        return int(lib().bpgpu_points_has_tables(self.handle))

    def download(self, off=0, n=None):
        n = len(self) - off if n is None else n
        out = ctypes.create_string_buffer(max(1, n * 2 * self.ctx.modbytes))
        self.ctx._check(lib().bpgpu_points_download(self.ctx.handle, self.handle, off, n, out), "points_download")
        return out.raw[:n * 2 * self.ctx.modbytes]

    def free(self):
        if self.handle:
            lib().bpgpu_points_free(self.handle)
            self.handle = None


class DeviceScalars:
    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle
        ctx._track(self)

    def __len__(self):
        return lib().bpgpu_scalars_len(self.handle)

    def download(self, off=0, n=None):
        n = len(self) - off if n is None else n
        out = ctypes.create_string_buffer(max(1, n * self.ctx.modbytes))
        self.ctx._check(lib().bpgpu_scalars_download(self.ctx.handle, self.handle, off, n, out), "scalars_download")
        return out.raw[:n * self.ctx.modbytes]

    def view(self, off, n):
        """a second handle onto self[off:off+n]; the storage lives until the last handle is freed"""
        h = ctypes.c_void_p()
        self.ctx._check(lib().bpgpu_scalars_view(self.handle, off, n, ctypes.byref(h)), "scalars_view")
        return DeviceScalars(self.ctx, h)

    def free(self):
        if self.handle:
            lib().bpgpu_scalars_free(self.handle)
            self.handle = None


class Context:
    """One CUDA stream + scratch on one device, for one curve (bpgpu_ctx)."""

    def __init__(self, curve=BLS12_381, device=0):
        h = ctypes.c_void_p()
        rc = lib().bpgpu_ctx_create(curve, device, ctypes.byref(h))
        if rc:
            raise BpgpuError(rc, "bpgpu_ctx_create")
        self.handle, self.curve, self.device = h, curve, device
        self.modbytes = lib().bpgpu_modbytes(curve)
        self._live = weakref.WeakSet()          # device handles created on this context: released before the context is

    def _track(self, obj):
        self._live.add(obj)

    def _check(self, rc, what):
        if rc:
            raise BpgpuError(rc, what)

    def close(self):
        """destroy the context; handles still alive on it are freed first (freeing one after its context would be a
        use-after-free on the C side)"""
        if self.handle:
            for obj in list(self._live):
                obj.free()
            lib().bpgpu_ctx_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    @property
    def stream(self):
        return lib().bpgpu_ctx_stream(self.handle)

    @property
    def launches(self):
        return lib().bpgpu_ctx_launches(self.handle)

    STAGES = ["digits", "scan", "scatter", "chunk_acc", "giant", "merge", "reduce_l1", "reduce_l2"]

    def range_witness(self, values, bits):
        """[a_L | a_R | a_O] of len(values) positive_no gadgets, built on the device (DeviceScalars of 3 * m * bits entries)"""
        m = len(values)
        arr = (ctypes.c_uint64 * m)(*values)
        h = ctypes.c_void_p()
        self._check(lib().bpgpu_range_witness(self.handle, arr, m, bits, ctypes.byref(h)), "range_witness")
        return DeviceScalars(self, h)

    def set_blocking_sync(self, on):
        """host waits of this context sleep on an event instead of spinning (more contexts than cores)"""
        self._check(lib().bpgpu_ctx_set_blocking_sync(self.handle, 1 if on else 0), "set_blocking_sync")

    def set_fixed_schedule(self, on):
        """table-path MSMs of this context with a fixed trip count per term (secret scalars)"""
        self._check(lib().bpgpu_ctx_set_fixed_schedule(self.handle, 1 if on else 0), "set_fixed_schedule")

    def set_profile(self, min_n):
        """record per-stage times of every MSM with at least min_n terms (0 / False = off)"""
        self._check(lib().bpgpu_ctx_set_profile(self.handle, int(min_n)), "set_profile")

    def msm_stage_ms(self):
        arr = (ctypes.c_double * 8)()
        runs = lib().bpgpu_msm_stage_ms(self.handle, arr, 8)
        return runs, dict(zip(self.STAGES, list(arr)[:8]))

    def host_alloc(self, nbytes):
        p = lib().bpgpu_host_alloc(nbytes)
        if not p:
            raise BpgpuError(-6, "bpgpu_host_alloc")
        return p

    def host_free(self, p):
        lib().bpgpu_host_free(p)

    def sync(self):
        self._check(lib().bpgpu_ctx_sync(self.handle), "sync")

    # ---- residency
    def upload_points(self, xy, n=None):
        n = len(xy) // (2 * self.modbytes) if n is None else n
        h = ctypes.c_void_p()
        self._check(lib().bpgpu_points_upload(self.handle, _buf(xy), n, ctypes.byref(h)), "points_upload")
        return DevicePoints(self, h)

    def points_from_hashes(self, hashes, n=None):
        """ECP::mapit on n SHAKE256 digests (MODBYTES each) -> device generator table."""
        n = len(hashes) // self.modbytes if n is None else n
        h = ctypes.c_void_p()
        self._check(lib().bpgpu_points_from_hashes(self.handle, _buf(hashes), n, ctypes.byref(h)), "points_from_hashes")
        return DevicePoints(self, h)

    def upload_scalars(self, be, n=None):
        n = len(be) // self.modbytes if n is None else n
        h = ctypes.c_void_p()
        self._check(lib().bpgpu_scalars_upload(self.handle, _buf(be), n, ctypes.byref(h)), "scalars_upload")
        return DeviceScalars(self, h)

    # ---- FieldElementVector algebra (device resident)
    def alloc_scalars(self, n):
        h = ctypes.c_void_p()
        self._check(lib().bpgpu_scalars_alloc(self.handle, n, ctypes.byref(h)), "scalars_alloc")
        return DeviceScalars(self, h)

    def fr_random(self, key, ctr0, n):
        """n scalars of the counter-mode stream SHAKE256(key || le64(ctr0 + i)) mod r, generated on the device"""
        h = ctypes.c_void_p()
        self._check(lib().bpgpu_fr_random(self.handle, _buf(key), len(key), ctr0, n, ctypes.byref(h)), "fr_random")
        return DeviceScalars(self, h)

    def fr_vandermonde(self, x_be, n):
        h = ctypes.c_void_p()
        self._check(lib().bpgpu_fr_vandermonde(self.handle, _buf(x_be), n, ctypes.byref(h)), "fr_vandermonde")
        return DeviceScalars(self, h)

    def _fr_binop(self, fn, name, a, b, n=None, aoff=0, boff=0, out=None, ooff=0):
        n = len(a) - aoff if n is None else n
        out = out or self.alloc_scalars(n + ooff)
        self._check(fn(self.handle, a.handle, aoff, b.handle, boff, n, out.handle, ooff), name)
        return out

    def fr_hadamard(self, a, b, **kw):
        return self._fr_binop(lib().bpgpu_fr_hadamard, "fr_hadamard", a, b, **kw)

    def fr_add(self, a, b, **kw):
        return self._fr_binop(lib().bpgpu_fr_add, "fr_add", a, b, **kw)

    def fr_sub(self, a, b, **kw):
        return self._fr_binop(lib().bpgpu_fr_sub, "fr_sub", a, b, **kw)

    def fr_scale(self, a, s_be, n=None, aoff=0, out=None, ooff=0):
        n = len(a) - aoff if n is None else n
        out = out or self.alloc_scalars(n + ooff)
        self._check(lib().bpgpu_fr_scale(self.handle, a.handle, aoff, n, _buf(s_be), out.handle, ooff), "fr_scale")
        return out

    def fr_inner_product(self, a, b, n=None, aoff=0, boff=0):
        n = len(a) - aoff if n is None else n
        out = ctypes.create_string_buffer(self.modbytes)
        self._check(lib().bpgpu_fr_inner_product(self.handle, a.handle, aoff, b.handle, boff, n, out), "fr_inner_product")
        return out.raw

    def fr_poly3_special_inner_product(self, l1, l2, l3, r0, r1, r3, n):
        out = ctypes.create_string_buffer(6 * self.modbytes)
        self._check(lib().bpgpu_fr_poly3_special_inner_product(self.handle, l1.handle, l2.handle, l3.handle, r0.handle,
                                                               r1.handle, r3.handle, n, out), "fr_poly3_special_inner_product")
        return out.raw

    def fr_batch_invert(self, a, n=None, aoff=0):
        n = len(a) - aoff if n is None else n
        out = self.alloc_scalars(n)
        prod = ctypes.create_string_buffer(self.modbytes)
        self._check(lib().bpgpu_fr_batch_invert(self.handle, a.handle, aoff, n, out.handle, 0, prod), "fr_batch_invert")
        return out, prod.raw

    # ---- inner-product argument rounds (device resident)
    def ipp_begin(self, G, H, Q_xy, Gf, Hf, a, b, n, goff=0, hoff=0):
        h = ctypes.c_void_p()
        self._check(lib().bpgpu_ipp_begin(self.handle, G.handle, goff, H.handle, hoff, _buf(Q_xy), Gf.handle, Hf.handle,
                                          a.handle, b.handle, n, ctypes.byref(h)), "ipp_begin")
        return h

    def ipp_round_LR(self, st):
        L, R = ctypes.create_string_buffer(2 * self.modbytes), ctypes.create_string_buffer(2 * self.modbytes)
        self._check(lib().bpgpu_ipp_round_LR(st, L, R), "ipp_round_LR")
        return L.raw, R.raw

    def ipp_fold(self, st, u_be, u_inv_be):
        self._check(lib().bpgpu_ipp_fold(st, _buf(u_be), _buf(u_inv_be)), "ipp_fold")

    def ipp_finish(self, st):
        a, b = ctypes.create_string_buffer(self.modbytes), ctypes.create_string_buffer(self.modbytes)
        self._check(lib().bpgpu_ipp_finish(st, a, b), "ipp_finish")
        lib().bpgpu_ipp_free(st)
        return a.raw, b.raw

    def ipp_verification_scalars(self, u_be, lg):
        h = ctypes.c_void_p()
        self._check(lib().bpgpu_ipp_verification_scalars(self.handle, _buf(u_be), lg, ctypes.byref(h)), "ipp_verification_scalars")
        return DeviceScalars(self, h)

    def ipp_verify_msm(self, G, H, Q_xy, Gf, Hf, a_be, b_be, u_be, L_xy, R_xy, lg, goff=0, hoff=0):
        out = ctypes.create_string_buffer(2 * self.modbytes)
        self._check(lib().bpgpu_ipp_verify_msm(self.handle, G.handle, goff, H.handle, hoff, _buf(Q_xy), Gf.handle, Hf.handle,
                                               _buf(a_be), _buf(b_be), _buf(u_be), _buf(L_xy), _buf(R_xy), lg, out), "ipp_verify_msm")
        return out.raw

    # ---- MSM (G1Vector::multi_scalar_mul_var_time & friends)
    def msm(self, points, scalars_be, off=0, n=None):
        n = len(scalars_be) // self.modbytes if n is None else n
        out = ctypes.create_string_buffer(2 * self.modbytes)
        self._check(lib().bpgpu_msm(self.handle, points.handle, off, n, _buf(scalars_be), out), "msm")
        return out.raw

    def msm_le32(self, points, scalars_le32, off=0, n=None):
        """resident bases, scalars as 32-byte little-endian canonical integers"""
        n = len(scalars_le32) // 32 if n is None else n
        out = ctypes.create_string_buffer(2 * self.modbytes)
        self._check(lib().bpgpu_msm_le32(self.handle, points.handle, off, n, _buf(scalars_le32), out), "msm_le32")
        return out.raw

    # two-halves form: begin enqueues, finish returns the point (one MSM in flight per context)
    def msm_device_begin(self, points, scalars, poff=0, soff=0, n=None):
        n = len(scalars) - soff if n is None else n
        self._check(lib().bpgpu_msm_device_begin(self.handle, points.handle, poff, n, scalars.handle, soff), "msm_device_begin")

    def msm_begin(self, points, scalars_be, off=0, n=None):
        n = len(scalars_be) // self.modbytes if n is None else n
        self._check(lib().bpgpu_msm_begin(self.handle, points.handle, off, n, _buf(scalars_be)), "msm_begin")

    def msm_le32_begin(self, points, scalars_le32, off=0, n=None):
        n = len(scalars_le32) // 32 if n is None else n
        self._check(lib().bpgpu_msm_le32_begin(self.handle, points.handle, off, n, _buf(scalars_le32)), "msm_le32_begin")

    def msm_refs_begin(self, points_xy, scalars_be, n=None):
        n = len(scalars_be) // self.modbytes if n is None else n
        self._check(lib().bpgpu_msm_refs_begin(self.handle, _buf(points_xy), _buf(scalars_be), n), "msm_refs_begin")

    def msm_finish(self):
        out = ctypes.create_string_buffer(2 * self.modbytes)
        self._check(lib().bpgpu_msm_finish(self.handle, out), "msm_finish")
        return out.raw

    def msm_device(self, points, scalars, poff=0, soff=0, n=None):
        n = len(scalars) - soff if n is None else n
        out = ctypes.create_string_buffer(2 * self.modbytes)
        self._check(lib().bpgpu_msm_device(self.handle, points.handle, poff, n, scalars.handle, soff, out), "msm_device")
        return out.raw

    def msm_refs(self, points_xy, scalars_be, n=None):
        n = len(scalars_be) // self.modbytes if n is None else n
        out = ctypes.create_string_buffer(2 * self.modbytes)
        self._check(lib().bpgpu_msm_refs(self.handle, _buf(points_xy), _buf(scalars_be), n, out), "msm_refs")
        return out.raw

    def commit_batch(self, bases_xy, scalars_be, count):
        """count independent combinations sum_j s[i*k+j] * B_j over k fixed bases (cached window tables)."""
        k = len(bases_xy) // (2 * self.modbytes)
        fb = ctypes.c_void_p()
        self._check(lib().bpgpu_fixed_bases_get(self.handle, _buf(bases_xy), k, ctypes.byref(fb)), "fixed_bases_get")
        out = ctypes.create_string_buffer(max(1, count * 2 * self.modbytes))
        self._check(lib().bpgpu_fixed_bases_commit(self.handle, fb, _buf(scalars_be), count, out), "fixed_bases_commit")
        return out.raw[:count * 2 * self.modbytes]

    def msm_batch_is_identity(self, runs, batch, fixed_scalars_be, var_points_xy, var_scalars_be, vn):
        """runs: list of (DevicePoints with tables, off, n) or bytes (one host base X||Y).  Returns `batch` verdict bytes."""
        arr = (FixedRun * max(1, len(runs)))()
        keep = []
        for i, r in enumerate(runs):
            if isinstance(r, (bytes, bytearray)):
                buf = ctypes.create_string_buffer(bytes(r), len(r))
                keep.append(buf)
                arr[i].host_base_xy, arr[i].n = ctypes.cast(buf, ctypes.c_void_p), 1
            else:
                arr[i].points, arr[i].off, arr[i].n = r[0].handle, r[1], r[2]
        out = ctypes.create_string_buffer(max(1, batch))
        self._check(lib().bpgpu_msm_batch_is_identity(self.handle, arr, len(runs), batch, _buf(fixed_scalars_be), _buf(var_points_xy),
                                                      _buf(var_scalars_be), vn, out), "msm_batch_is_identity")
        return out.raw[:batch]

    def msm_parts(self, parts):
        """parts: list of (points, scalars, n[, points_off, scalars_off]); points is DevicePoints or bytes (X||Y),
        scalars is DeviceScalars or bytes (big endian)."""
        arr = (MsmPart * len(parts))()
        keep = []
        for i, p in enumerate(parts):
            pts, sc, n = p[0], p[1], p[2]
            poff = p[3] if len(p) > 3 else 0
            soff = p[4] if len(p) > 4 else 0
            if isinstance(pts, DevicePoints):
                arr[i].points, arr[i].points_off = pts.handle, poff
            else:
                buf = ctypes.create_string_buffer(bytes(pts), len(pts))
                keep.append(buf)
                arr[i].host_points_xy = ctypes.cast(buf, ctypes.c_void_p)
            if isinstance(sc, DeviceScalars):
                arr[i].scalars, arr[i].scalars_off = sc.handle, soff
            else:
                buf = ctypes.create_string_buffer(bytes(sc), len(sc))
                keep.append(buf)
                arr[i].host_scalars_be = ctypes.cast(buf, ctypes.c_void_p)
            arr[i].n = n
        out = ctypes.create_string_buffer(2 * self.modbytes)
        self._check(lib().bpgpu_msm_parts(self.handle, arr, len(parts), out), "msm_parts")
        return out.raw

    # ---- fused R1CS Fr kernels
    def r1cs_prover_polys(self, n, aL, aR, sR, wL, wR, wO, y_be):
        hs = [ctypes.c_void_p() for _ in range(4)]
        self._check(lib().bpgpu_r1cs_prover_polys(self.handle, n, aL.handle, aR.handle, sR.handle, wL.handle, wR.handle, wO.handle,
                                                  _buf(y_be), *[ctypes.byref(h) for h in hs]), "r1cs_prover_polys")
        return [DeviceScalars(self, h) for h in hs]

    def r1cs_prover_eval(self, n, n1, N, l1, l2, l3, r0, r1, r3, x_be, u_be, y_be):
        hs = [ctypes.c_void_p() for _ in range(4)]
        self._check(lib().bpgpu_r1cs_prover_eval(self.handle, n, n1, N, l1.handle, l2.handle, l3.handle, r0.handle, r1.handle, r3.handle,
                                                 _buf(x_be), _buf(u_be), _buf(y_be), *[ctypes.byref(h) for h in hs]), "r1cs_prover_eval")
        return [DeviceScalars(self, h) for h in hs]

    def r1cs_verifier_scalars(self, n, n1, N, wL, wR, wO, s, y_be, x_be, a_be, b_be, u_be):
        h = ctypes.c_void_p()
        delta = ctypes.create_string_buffer(self.modbytes)
        self._check(lib().bpgpu_r1cs_verifier_scalars(self.handle, n, n1, N, wL.handle, wR.handle, wO.handle, s.handle, _buf(y_be),
                                                      _buf(x_be), _buf(a_be), _buf(b_be), _buf(u_be), ctypes.byref(h), delta),
                    "r1cs_verifier_scalars")
        return DeviceScalars(self, h), delta.raw

    def mimc_witness(self, xl, xr, constants, rounds, count=None, with_multipliers=True):
        """batched mimc(xl, xr, constants, rounds) -> (image, a_L, a_R, a_O) device vectors (a_* None without multipliers)"""
        count = len(xl) if count is None else count
        hs = [ctypes.c_void_p() for _ in range(4)]
        refs = [ctypes.byref(h) for h in hs]
        if not with_multipliers:
            refs[1:] = [None, None, None]
        self._check(lib().bpgpu_mimc_witness(self.handle, xl.handle, xr.handle, count, constants.handle, rounds, *refs), "mimc_witness")
        out = [DeviceScalars(self, hs[0])]
        out += [DeviceScalars(self, h) for h in hs[1:]] if with_multipliers else [None, None, None]
        return out

    def poseidon_permutation(self, inputs, count, width, frb, pr, fre, sbox, round_keys, mds):
        h = ctypes.c_void_p()
        self._check(lib().bpgpu_poseidon_permutation(self.handle, inputs.handle, count, width, frb, pr, fre, sbox, round_keys.handle, mds.handle,
                                                     ctypes.byref(h)), "poseidon_permutation")
        return DeviceScalars(self, h)

    # ---- host layer (include/bphost.h): the reference's API end to end
    def get_generators(self, prefix, n, precompute=False):
        """utils::get_generators(prefix, n) as a device table (utils/mod.rs:16-23); precompute=True also builds the window
        tables (bpgpu_points_precompute)."""
        h = ctypes.c_void_p()
        self._check(lib().bph_get_generators(self.handle, prefix.encode(), n, ctypes.byref(h)), "get_generators")
        t = DevicePoints(self, h)
        if os.environ.get("BPGPU_PRECOMPUTE") == "1" and n <= 4096:      # test switch: run every caller on the table path
            precompute = True
        return t.precompute() if precompute else t

    def g1_from_msg_hash(self, msg):
        out = ctypes.create_string_buffer(2 * self.modbytes)
        self._check(lib().bph_g1_from_msg_hash(self.handle, _buf(msg), len(msg), out), "g1_from_msg_hash")
        return out.raw

    def _proof_call(self, fn, what, *args_before, extra_after=()):
        cap = 1 << 16
        while True:
            buf, ln = ctypes.create_string_buffer(cap), ctypes.c_size_t()
            rc = fn(*args_before, buf, cap, ctypes.byref(ln), *extra_after)
            if rc == -10 and ln.value > cap:
                cap = ln.value
                continue
            self._check(rc, what)
            return buf.raw[:ln.value]

    def ipp_create(self, label, G, H, Q_xy, Gf_be, Hf_be, a_be, b_be, n):
        """IPP::create_ipp with a fresh Transcript::new(label) -> L.. | R.. | a | b"""
        return self._proof_call(lib().bph_ipp_create, "ipp_create", self.handle, label, G.handle, H.handle, _buf(Q_xy), _buf(Gf_be),
                                _buf(Hf_be), _buf(a_be), _buf(b_be), n)

    def ipp_verify(self, label, n, Gf_be, Hf_be, P_xy, Q_xy, G, H, proof):
        """IPP::verify_ipp -> True / False (VerificationError); other errors raise."""
        rc = lib().bph_ipp_verify(self.handle, label, n, _buf(Gf_be), _buf(Hf_be), _buf(P_xy), _buf(Q_xy), G.handle, H.handle,
                                  _buf(proof), len(proof))
        if rc == -4:
            return False
        self._check(rc, "ipp_verify")
        return True

    def bound_check_prove(self, label, g_xy, h_xy, G, H, val, lower, upper, bits, seed=None, randomness_be=None):
        """gen_proof_of_bounded_num -> (proof bytes, 3 commitments X||Y).  seed=None: OS entropy."""
        comms = ctypes.create_string_buffer(3 * 2 * self.modbytes)
        rnd = _buf(randomness_be) if randomness_be is not None else None
        proof = self._proof_call(lib().bph_bound_check_prove, "bound_check_prove", self.handle, label, _buf(g_xy), _buf(h_xy), G.handle,
                                 H.handle, val, rnd, lower, upper, bits, 0 if seed is None else 1, seed or 0, extra_after=(comms,))
        return proof, comms.raw

    def bound_check_verify(self, label, g_xy, h_xy, G, H, lower, upper, bits, proof, comms, r_be=None):
        rc = lib().bph_bound_check_verify(self.handle, label, _buf(g_xy), _buf(h_xy), G.handle, H.handle, lower, upper, bits, _buf(proof),
                                          len(proof), _buf(comms), _buf(r_be) if r_be is not None else None)
        if rc == -4:
            return False
        self._check(rc, "bound_check_verify")
        return True

    def range_prove(self, label, g_xy, h_xy, G, H, values, bits, seed=None):
        """m x positive_no_gadget in one constraint system -> (proof bytes, m commitments X||Y)."""
        m = len(values)
        arr = (ctypes.c_uint64 * max(1, m))(*values)
        comms = ctypes.create_string_buffer(max(1, m * 2 * self.modbytes))
        proof = self._proof_call(lib().bph_range_prove, "range_prove", self.handle, label, _buf(g_xy), _buf(h_xy), G.handle, H.handle,
                                 ctypes.cast(arr, ctypes.c_void_p), m, bits, 0 if seed is None else 1, seed or 0, extra_after=(comms,))
        return proof, comms.raw[:m * 2 * self.modbytes]

    def range_verify(self, label, g_xy, h_xy, G, H, m, bits, proof, comms, r_be=None):
        rc = lib().bph_range_verify(self.handle, label, _buf(g_xy), _buf(h_xy), G.handle, H.handle, m, bits, _buf(proof), len(proof),
                                    _buf(comms), _buf(r_be) if r_be is not None else None)
        if rc == -4:
            return False
        self._check(rc, "range_verify")
        return True

    def shuffle_prove(self, label, g_xy, h_xy, G, H, xs, ys, bits=0, seed=None):
        """two-phase circuit: {ys} is a permutation of {xs}, xs[0] range-checked to `bits` bits -> (proof, 2k commitments)"""
        k = len(xs)
        ax, ay = (ctypes.c_uint64 * k)(*xs), (ctypes.c_uint64 * k)(*ys)
        comms = ctypes.create_string_buffer(2 * k * 2 * self.modbytes)
        proof = self._proof_call(lib().bph_shuffle_prove, "shuffle_prove", self.handle, label, _buf(g_xy), _buf(h_xy), G.handle, H.handle,
                                 ctypes.cast(ax, ctypes.c_void_p), ctypes.cast(ay, ctypes.c_void_p), k, bits, 0 if seed is None else 1,
                                 seed or 0, extra_after=(comms,))
        return proof, comms.raw

    def shuffle_verify(self, label, g_xy, h_xy, G, H, k, bits, proof, comms, r_be=None):
        rc = lib().bph_shuffle_verify(self.handle, label, _buf(g_xy), _buf(h_xy), G.handle, H.handle, k, bits, _buf(proof), len(proof),
                                      _buf(comms), _buf(r_be) if r_be is not None else None)
        if rc == -4:
            return False
        self._check(rc, "shuffle_verify")
        return True

    # ---- self-test hooks
    def selftest_field(self, field, op, a, b):
        n = len(a) // self.modbytes
        out = ctypes.create_string_buffer(max(1, len(a)))
        self._check(lib().bpgpu_selftest_field(self.handle, field, op, _buf(a), _buf(b), n, out), "selftest_field")
        return out.raw[:len(a)]

    def selftest_group(self, op, p_xy, q_xy, scalars_be):
        n = len(p_xy) // (2 * self.modbytes)
        out = ctypes.create_string_buffer(max(1, len(p_xy)))
        self._check(lib().bpgpu_selftest_group(self.handle, op, _buf(p_xy), _buf(q_xy), _buf(scalars_be), n, out),
                    "selftest_group")
        return out.raw[:len(p_xy)]

    def int_pipe_bench(self, kind, iters):
        ops, ms = ctypes.c_double(), ctypes.c_double()
        self._check(lib().bpgpu_int_pipe_bench(self.handle, kind, iters, ctypes.byref(ops), ctypes.byref(ms)), "int_pipe_bench")
        return ops.value, ms.value
