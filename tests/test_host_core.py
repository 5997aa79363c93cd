"""CPU tests of the per-proof DEVICE logic of the batched verifier (csrc/merlin.cuh, csrc/verify_core.cuh): the functions
the kernels loop over, compiled for the host (tests/host_check2.cpp), against the oracle's verifier
(oracle/r1cs.py Verifier.verification_msm = verifier.rs:267-449) on the committed golden proofs.  The circuit matrices
come from the product's own host code (bph_range_circuit_csr / bph_bound_check_circuit_csr: host only, no device)."""
import ctypes
import json
import os
import subprocess

import pytest

from oracle import r1cs as or1cs
from oracle.curves import CURVES
from oracle.merlin import Transcript

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def hc2(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hc2") / "host_check2.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "host_check2.cpp")])
    return ctypes.CDLL(so)


def test_merlin_kat_through_device_transcript_code(hc2):
    """Merlin's published `equivalence_simple` vector through StrobeHD (the code k_vb_transcript runs per proof)"""
    out = ctypes.create_string_buffer(32)
    hc2.hc2_merlin_kat(b"test protocol", b"some label", b"some data", 9, b"challenge", out, 32)
    assert out.raw.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_device_transcript_code_long_messages(hc2):
    """messages that cross the 166-byte STROBE rate, against the oracle's transcript"""
    for ln in (0, 1, 97, 165, 166, 167, 400):
        msg = bytes((7 * i + ln) & 0xFF for i in range(ln))
        t = Transcript(b"x" * 30)
        t.append_message(b"a fairly long label", msg)
        exp = t.challenge_bytes(b"c", 48)
        out = ctypes.create_string_buffer(48)
        hc2.hc2_merlin_kat(b"x" * 30, b"a fairly long label", msg, ln, b"c", out, 48)
        assert out.raw == exp, ln


def _oracle_terms(C, d, bp_lib_csr, build_verifier):
    """oracle arg1 split the way the device lays it out: fixed = g_scalars | h_scalars | g | h, var = the proof's own points"""
    proof = or1cs.R1CSProof.from_bytes(C, bytes.fromhex(d["proof"]))
    comms_b = bytes.fromhex(d["commitments"])
    m = len(comms_b) // (2 * C.MODBYTES)
    comms = [C.g1_from_xy_bytes(comms_b[i * 2 * C.MODBYTES:(i + 1) * 2 * C.MODBYTES]) for i in range(m)]
    v = or1cs.Verifier(C, Transcript(d["label"].encode(), C))
    build_verifier(v, comms)
    n = v.num_vars
    N = 1 << max(0, (n - 1).bit_length())
    lg = N.bit_length() - 1
    g, h = C.g1_from_msg_hash(b"g"), C.g1_from_msg_hash(b"h")
    G = [C.INF] * N                                       # the generators do not enter the scalars
    key = (12345).to_bytes(8, "little") + b"blind"
    rnd = C.synth_scalar(12345, 3, b"blind")              # draw 3 of the stream keyed `key`
    arg2, arg1 = v.verification_msm(proof, g, h, G, G, rnd)
    head = 6 + m + 5
    var = arg1[:head] + arg1[head + 2 + 2 * N:]
    fixed = arg1[head + 2:head + 2 + 2 * N] + arg1[head:head + 2]
    return proof, comms_b, m, n, N, lg, key, fixed, var


def _run_core(hc2, bp, C, cid, d, csr, key, ctr, mode_state=True):
    proof_b, comms_b = bytes.fromhex(d["proof"]), bytes.fromhex(d["commitments"])
    n, m = csr["n"], csr["m"]
    N = 1 << max(0, (n - 1).bit_length())
    lg = N.bit_length() - 1
    mb = C.MODBYTES
    rows = (ctypes.c_uint32 * len(csr["row_start"]))(*csr["row_start"])
    eq = (ctypes.c_uint32 * max(1, len(csr["ent_q"])))(*csr["ent_q"])
    fixed = ctypes.create_string_buffer((2 * N + 2) * mb)
    vn = 6 + m + 5 + 2 * lg
    var = ctypes.create_string_buffer(vn * mb)
    ok = ctypes.create_string_buffer(vn)
    state = bp.r1cs_transcript_state(d["label"].encode())
    chal = None
    if not mode_state:
        # challenges from a host transcript, as a caller with its own Merlin would supply them
        chal = bp.r1cs_replay_challenges(cid, d["label"].encode(), proof_b, comms_b, m, lg)
        state = None
    st = hc2.hc2_verify_terms(cid, state, chal, proof_b, comms_b, m, lg, n, csr["q"], rows, eq, csr["ent_c_be"], key, len(key),
                              ctypes.c_uint64(ctr), fixed, var, ok)
    return st, fixed.raw, var.raw, ok.raw


CASES = [("range_small.json", "bls_m2_b8"), ("range_small.json", "bn_m3_b5"), ("range_config5_unit.json", "bls_m1_b64"),
         ("range_config3_reduced.json", "bn_m8_b64")]


@pytest.mark.parametrize("name,key", CASES)
def test_device_verifier_scalars_match_oracle(hc2, bp, name, key):
    """transcript replay + CSR flattening + s / y^-i / g / h / head scalars, element by element"""
    d = json.load(open(os.path.join(GOLD, name)))[key]
    C = CURVES[d["curve"]]
    cid = 0 if d["curve"] == "BLS12_381" else 1
    m, bits = d["m"], d["bits"]

    def build(v, comms):
        for com in comms:
            or1cs.positive_no_gadget(v, or1cs.AllocatedQuantity(v.commit(com), None), bits)
    proof, comms_b, m_, n, N, lg, rkey, exp_fixed, exp_var = _oracle_terms(C, d, None, build)
    csr = bp.range_circuit_csr(cid, m, bits)
    assert (csr["n"], csr["m"]) == (n, m)
    st, fixed, var, ok = _run_core(hc2, bp, C, cid, d, csr, rkey, 3)
    assert st == 0
    mb = C.MODBYTES
    got_fixed = [int.from_bytes(fixed[i * mb:(i + 1) * mb], "big") for i in range(2 * N + 2)]
    got_var = [int.from_bytes(var[i * mb:(i + 1) * mb], "big") for i in range(len(exp_var))]
    assert got_var == exp_var
    assert got_fixed == exp_fixed
    assert ok == b"\x01" * len(ok)
    # host transcripts + uploaded challenges (bpgpu_r1cs_verify_batch's second mode) arrive at the same scalars
    st2, fixed2, var2, _ = _run_core(hc2, bp, C, cid, d, csr, rkey, 3, mode_state=False)
    assert st2 == 0 and fixed2 == fixed and var2 == var


def test_device_verifier_scalars_bound_check(hc2, bp):
    fx = json.load(open(os.path.join(GOLD, "bound_check_8bit.json")))
    for cname, d in fx.items():
        C = CURVES[cname]
        cid = 0 if cname == "BLS12_381" else 1

        def build(v, comms):
            or1cs.verify_bounded_num(v, d["lower"], d["upper"], d["bits"], comms)
        proof, comms_b, m, n, N, lg, rkey, exp_fixed, exp_var = _oracle_terms(C, d, None, build)
        csr = bp.bound_check_circuit_csr(cid, d["lower"], d["upper"], d["bits"])
        assert (csr["n"], csr["m"]) == (n, 3)
        st, fixed, var, ok = _run_core(hc2, bp, C, cid, d, csr, rkey, 3)
        mb = C.MODBYTES
        assert st == 0
        assert [int.from_bytes(var[i * mb:(i + 1) * mb], "big") for i in range(len(exp_var))] == exp_var
        assert [int.from_bytes(fixed[i * mb:(i + 1) * mb], "big") for i in range(2 * N + 2)] == exp_fixed


def test_device_format_checks(hc2, bp):
    """what the device rejects before any arithmetic: a missing 0x04 tag, a scalar >= r, a coordinate >= p, a point off the curve"""
    d = json.load(open(os.path.join(GOLD, "range_small.json")))["bls_m2_b8"]
    C = CURVES["BLS12_381"]
    csr = bp.range_circuit_csr(0, d["m"], d["bits"])
    key = b"k" * 8
    proof = bytearray(bytes.fromhex(d["proof"]))
    PB, mb = 97, 48
    good = dict(d)
    assert _run_core(hc2, bp, C, 0, good, csr, key, 0)[0] == 0
    bad = bytearray(proof); bad[2 * PB] = 2                                    # tag of S1
    assert _run_core(hc2, bp, C, 0, dict(d, proof=bytes(bad).hex()), csr, key, 0)[0] == -5
    bad = bytearray(proof); bad[11 * PB:11 * PB + mb] = (C.r).to_bytes(mb, "big")   # t_x = r: not canonical
    assert _run_core(hc2, bp, C, 0, dict(d, proof=bytes(bad).hex()), csr, key, 0)[0] == -5
    bad = bytearray(proof); bad[11 * PB] = 1                                   # t_x >= 2^376
    assert _run_core(hc2, bp, C, 0, dict(d, proof=bytes(bad).hex()), csr, key, 0)[0] == -5
    # points: flagged per point by the decoder (k_vb_points turns the flag into E_FORMAT)
    bad = bytearray(proof); bad[1 + 95] ^= 1                                   # y of A_I1 off by one: off the curve
    st, _, _, ok = _run_core(hc2, bp, C, 0, dict(d, proof=bytes(bad).hex()), csr, key, 0)
    assert st == 0 and ok[0] == 0 and ok[1:] == b"\x01" * (len(ok) - 1)
    x = int.from_bytes(proof[1:49], "big")
    bad = bytearray(proof); bad[1:49] = (x + C.p).to_bytes(48, "big")          # x + p: same residue, not canonical
    st, _, _, ok = _run_core(hc2, bp, C, 0, dict(d, proof=bytes(bad).hex()), csr, key, 0)
    assert ok[0] == 0
    # the identity (0, 1) IS a point (A_I2, A_O2, S2 of every one-phase proof)
    assert ok[3:6] == b"\x01\x01\x01"
    for cid, Cc in ((0, CURVES["BLS12_381"]), (1, CURVES["BN254"])):
        m_ = Cc.MODBYTES
        assert hc2.hc2_point_valid(cid, Cc.g1_xy_bytes(Cc.INF)) == 1
        assert hc2.hc2_point_valid(cid, Cc.g1_xy_bytes(Cc.from_affine(Cc.g))) == 1
        assert hc2.hc2_point_valid(cid, bytes(2 * m_)) == 0                    # (0, 0)
        gx, gy = Cc.g
        assert hc2.hc2_point_valid(cid, gx.to_bytes(m_, "big") + ((gy + 1) % Cc.p).to_bytes(m_, "big")) == 0
        assert hc2.hc2_point_valid(cid, gx.to_bytes(m_, "big") + (Cc.p - gy).to_bytes(m_, "big")) == 1      # -G


def test_wide_scalar_reduction(hc2):
    """FieldElement::from(&[u8; MODBYTES]) for values up to 2^384 (a transcript challenge is 48 uniform bytes on BLS12-381)"""
    import random
    rnd = random.Random(9)
    for cid, C in ((0, CURVES["BLS12_381"]), (1, CURVES["BN254"])):
        mb = C.MODBYTES
        vals = [0, 1, C.r - 1, C.r, C.r + 1, (1 << (8 * mb)) - 1, 1 << 256 if mb > 32 else 5, (1 << 255) + 12345]
        vals += [rnd.getrandbits(8 * mb) for _ in range(200)]
        for v in vals:
            v %= 1 << (8 * mb)
            out = ctypes.create_string_buffer(mb)
            hc2.hc2_fr_from_be_wide(cid, v.to_bytes(mb, "big"), out)
            assert int.from_bytes(out.raw, "big") == v % C.r


def test_glv_constants_and_split(hc2):
    """GLV on BLS12-381 G1 (csrc/curves.cuh): phi(P) = (beta x, y) = lambda P checked with the oracle's group law, and the
    device's split k = k1 + k2 lambda with both halves below 2^128, on edge cases and random scalars"""
    import random
    C = CURVES["BLS12_381"]
    lam = 0xac45a4010001a40200000000ffffffff
    assert (lam * lam + lam + 1) % C.r == 0
    G = C.from_affine(C.g)
    for P in (G, C.mul(G, 7777)):
        x, y = C.to_affine(P)
        out = ctypes.create_string_buffer(48)
        hc2.hc2_glv_phi_x(x.to_bytes(48, "big"), out)
        assert (int.from_bytes(out.raw, "big"), y) == C.to_affine(C.mul(P, lam))
    rnd = random.Random(5)
    ks = [0, 1, lam - 1, lam, lam + 1, C.r - 1, C.r - 2, lam * ((C.r - 1) // lam), lam * lam % C.r, (1 << 254) + 3]
    ks += [rnd.randrange(C.r) for _ in range(3000)]
    for k in ks:
        a, b = ctypes.create_string_buffer(16), ctypes.create_string_buffer(16)
        hc2.hc2_glv_split(k.to_bytes(32, "big"), a, b)
        k1, k2 = int.from_bytes(a.raw, "big"), int.from_bytes(b.raw, "big")
        assert k1 + k2 * lam == k and k1 < lam, hex(k)
