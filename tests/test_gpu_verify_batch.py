"""GPU parity tests of the batched verifier whose per-proof scalars are built on the device (csrc/verifybatch.cu):
bpgpu_circuit_* (SURVEY.md section 8 f2), bpgpu_r1cs_verify_batch (rows a3, a12, f3 for slabs, config 5)."""
import json
import os

import pytest

from oracle import r1cs as or1cs
from oracle.curves import CURVES
from oracle.merlin import Transcript
from tests.util import curve_of, dec_scalars

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle_arg1(C, label, proof_b, comms_b, build, rnd):
    proof = or1cs.R1CSProof.from_bytes(C, proof_b)
    m = len(comms_b) // (2 * C.MODBYTES)
    comms = [C.g1_from_xy_bytes(comms_b[i * 2 * C.MODBYTES:(i + 1) * 2 * C.MODBYTES]) for i in range(m)]
    v = or1cs.Verifier(C, Transcript(label, C))
    build(v, comms)
    n = v.num_vars
    N = 1 << max(0, (n - 1).bit_length())
    G = [C.INF] * N
    _, arg1 = v.verification_msm(proof, C.g1_from_msg_hash(b"g"), C.g1_from_msg_hash(b"h"), G, G, rnd)
    head = 6 + m + 5
    return arg1[head + 2:head + 2 + 2 * N] + arg1[head:head + 2], arg1[:head] + arg1[head + 2 + 2 * N:], v


@pytest.mark.parametrize("name,key", [("range_small.json", "bls_m2_b8"), ("range_small.json", "bn_m3_b5"),
                                      ("range_config5_unit.json", "bls_m1_b64"), ("range_config3_reduced.json", "bn_m8_b64")])
def test_device_built_verification_scalars_match_oracle(bp, ctx_bls, ctx_bn, name, key):
    """every scalar of the verification MSM (g_scalars, h_scalars, the scalars of g and h, of A_I.., V_j, T_i, L_k, R_k:
    verifier.rs:341-429, ipp.rs:295-312) as the DEVICE builds it, element by element against the oracle -- with the
    transcripts replayed on the device and with challenges from a host transcript; the verdicts are Ok."""
    d = json.load(open(os.path.join(GOLD, name)))[key]
    ctx = ctx_bls if d["curve"] == "BLS12_381" else ctx_bn
    C = curve_of(ctx)
    m, bits, label = d["m"], d["bits"], d["label"].encode()
    proof_b, comms_b = bytes.fromhex(d["proof"]), bytes.fromhex(d["commitments"])
    n = m * bits
    N = 1 << max(0, (n - 1).bit_length())
    lg = N.bit_length() - 1

    def build(v, comms):
        for com in comms:
            or1cs.positive_no_gadget(v, or1cs.AllocatedQuantity(v.commit(com), None), bits)
    count = 3                                             # the same proof three times: proof i takes draw i of the r stream
    key_b = (777).to_bytes(8, "little") + b"blind"
    circ = bp.Circuit(ctx, bp.range_circuit_csr(ctx.curve, m, bits))
    dG, dH = ctx.get_generators("G", N), ctx.get_generators("H", N)
    gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
    state = bp.r1cs_transcript_state(label)
    chal = bp.r1cs_replay_challenges(ctx.curve, label, proof_b, comms_b, m, lg) * count
    for kw in ({"state": state}, {"challenges": chal}):
        verdicts, fixed, var = circ.verify_batch(dG, dH, gx, hx, count, proof_b * count, len(proof_b), comms_b * count, key=key_b,
                                                 terms=True, **kw)
        F, vn = 2 * N + 2, 6 + m + 5 + 2 * lg
        gf, gv = dec_scalars(C, fixed), dec_scalars(C, var)
        for i in range(count):
            exp_fixed, exp_var, _ = _oracle_arg1(C, label, proof_b, comms_b, build, C.synth_scalar(777, i, b"blind"))
            bad_v = [k for k in range(vn) if gv[i * vn + k] != exp_var[k]]
            bad_f = [k for k in range(F) if gf[i * F + k] != exp_fixed[k]]
            assert not bad_v and not bad_f, (i, list(kw.keys()), bad_v, bad_f)
        assert verdicts == [0] * count
    circ.free()


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_circuit_flatten_matches_oracle(bp, ctx_bls, ctx_bn, which):
    """flattened_constraints (verifier.rs:149-193 / prover.rs:142-184) as a device sparse mat-vec over the recorded circuit"""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    for kind in ("range", "bound"):
        v = or1cs.Verifier(C, Transcript(b"x", C))
        if kind == "range":
            m, bits = 3, 7
            for _ in range(m):
                or1cs.positive_no_gadget(v, or1cs.AllocatedQuantity(v.commit(C.INF), None), bits)
            csr = bp.range_circuit_csr(ctx.curve, m, bits)
        else:
            or1cs.verify_bounded_num(v, 100, 1000, 10, [C.INF] * 3)
            csr = bp.bound_check_circuit_csr(ctx.curve, 100, 1000, 10)
            m = 3
        n = v.num_vars
        assert (csr["n"], csr["m"], csr["q"]) == (n, m, len(v.constraints))
        circ = bp.Circuit(ctx, csr)
        for z in (1, 2, C.r - 1, C.synth_scalar(5, 0)):
            wL, wR, wO, wV, wc = v.flattened_constraints(z)
            out = circ.flatten(C.fr_to_bytes(z))
            assert dec_scalars(C, out.download()) == wL + wR + wO + wV + [wc]
            out.free()
        circ.free()


def test_large_batch_keeps_per_proof_verdicts_across_slabs(bp, ctx_bls):
    """4100 proofs (more than one 4096-proof device slab), tampered proofs on both sides of the slab boundary and inside:
    every proof keeps its own verdict in all three modes (device transcripts / host transcripts / round 1's host scalars)."""
    ctx = ctx_bls
    m, bits, count = 1, 8, 4100
    dG, dH = ctx.get_generators("G", 8), ctx.get_generators("H", 8)
    gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
    values = [(37 * i + 11) % 256 for i in range(count)]
    proofs, stride, comms = bp.range_prove_batch(ctx, b"Big", gx, hx, dG, dH, values, m, bits, seed=4000)
    assert bp.range_verify_batch(ctx, b"Big", gx, hx, dG, dH, count, m, bits, proofs, stride, comms) == [0] * count
    bad = bytearray(proofs)
    PB, mb = 97, 48
    exp = [0] * count
    bad[17 * stride + 11 * PB + mb - 1] ^= 1               # t_x of proof 17
    exp[17] = -4
    bad[4095 * stride + stride - 1] ^= 2                   # b of the last proof of slab 0
    exp[4095] = -4
    bad[4096 * stride + 6 * PB + 1 + 2 * mb - 1] ^= 1      # y of T_1 of the first proof of slab 1: off the curve
    exp[4096] = -5
    bad[4099 * stride + 3 * PB] = 3                        # tag of A_I2 of the last proof
    exp[4099] = -5
    bad[2048 * stride + 11 * PB + 2 * mb:2048 * stride + 11 * PB + 3 * mb] = (curve_of(ctx).r + 5).to_bytes(mb, "big")   # e_blinding >= r
    exp[2048] = -5
    badc = bytearray(comms)
    badc[1000 * 2 * mb + 5] ^= 4                           # x of the commitment of proof 1000: off the curve
    exp[1000] = -5
    for mode in (0, 1):
        v = bp.range_verify_batch(ctx, b"Big", gx, hx, dG, dH, count, m, bits, bytes(bad), stride, bytes(badc), nthreads=4, mode=mode)
        assert v == exp, mode
    v2 = bp.range_verify_batch(ctx, b"Big", gx, hx, dG, dH, count, m, bits, bytes(bad), stride, bytes(badc), nthreads=4, mode=2)
    assert [x != 0 for x in v2] == [x != 0 for x in exp]
    # a strided record layout (proof_stride > proof length)
    wide = b"".join(proofs[i * stride:(i + 1) * stride] + b"\xee" * 13 for i in range(40))
    assert bp.range_verify_batch(ctx, b"Big", gx, hx, dG, dH, 40, m, bits, wide, stride + 13, comms) == [0] * 40


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_bound_check_batch(bp, ctx_bls, ctx_bn, which):
    """verify_proof_of_bounded_num (bound_check.rs:163-178; config 5's 32-bit variant at reduced width) in bulk: three
    commitments per proof, constant terms in the constraints (wc != 0)"""
    ctx = ctx_bls if which == "bls" else ctx_bn
    bits, lower, upper = 8, 10, 100
    dG, dH = ctx.get_generators("G", 2 * bits), ctx.get_generators("H", 2 * bits)
    gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
    vals = [10, 11, 57, 99, 100, 75]
    proofs, comms = b"", b""
    for i, v in enumerate(vals):
        p, c = ctx.bound_check_prove(b"Bounds", gx, hx, dG, dH, v, lower, upper, bits, seed=50 + i)
        assert ctx.bound_check_verify(b"Bounds", gx, hx, dG, dH, lower, upper, bits, p, c)
        proofs, comms = proofs + p, comms + c
    stride = len(proofs) // len(vals)
    for mode in (0, 1):
        assert bp.bound_check_verify_batch(ctx, b"Bounds", gx, hx, dG, dH, len(vals), lower, upper, bits, proofs, stride, comms,
                                           mode=mode) == [0] * len(vals)
        # other bounds: a different circuit (the constants differ), nothing verifies
        assert bp.bound_check_verify_batch(ctx, b"Bounds", gx, hx, dG, dH, len(vals), lower, upper + 1, bits, proofs, stride, comms,
                                           mode=mode) == [-4] * len(vals)
    mbytes = ctx.modbytes
    swapped = comms[:3 * 2 * mbytes] + comms[6 * 2 * mbytes:9 * 2 * mbytes] + comms[3 * 2 * mbytes:6 * 2 * mbytes] + comms[9 * 2 * mbytes:]
    assert bp.bound_check_verify_batch(ctx, b"Bounds", gx, hx, dG, dH, len(vals), lower, upper, bits, proofs, stride, swapped) == \
        [0, -4, -4, 0, 0, 0]


@pytest.mark.gpu
@pytest.mark.parametrize("which,m,bits", [("bls", 1, 64), ("bn", 2, 16)])
def test_wide_generator_tables_change_no_byte(which, m, bits, bp, ctx_bls, ctx_bn, monkeypatch):
    """16-bit-window tables (bpgpu_points_precompute_wide) halve the additions of the batch calls' fixed terms; the batch
    prover's proofs and the batch verifier's verdicts (incl. tampered proofs) are the same with and without them, and equal
    the single-proof prover's bytes."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    n = m * bits
    count = 300
    gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
    vals = [(0x9E3779B97F4A7C15 * (i + 1)) % (1 << min(bits, 63)) for i in range(count * m)]
    out = {}
    for wide in ("0", "64"):
        monkeypatch.setenv("BPH_WIDE_TABLES", wide)
        G, H = ctx.get_generators("G", n, precompute=True), ctx.get_generators("H", n, precompute=True)
        proofs, stride, comms = bp.range_prove_batch(ctx, b"wide", gx, hx, G, H, vals, m, bits, seed=5)
        assert (G.table_level == 2) == (wide != "0")
        bad = bytearray(proofs)
        for i in (3, 150, 299):
            bad[i * stride + stride - 1] ^= 1
        v = bp.range_verify_batch(ctx, b"wide", gx, hx, G, H, count, m, bits, bytes(bad), stride, comms)
        assert [i for i, x in enumerate(v) if x != 0] == [3, 150, 299]
        out[wide] = (proofs, comms)
        if wide != "0":
            single, scomms = ctx.range_prove(b"wide", gx, hx, G, H, vals[7 * m:8 * m], bits, seed=5 + 7)
            assert proofs[7 * stride:7 * stride + len(single)] == single
        G.free()
        H.free()
    assert out["0"] == out["64"]
