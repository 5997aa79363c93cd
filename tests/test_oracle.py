"""CPU tests of the oracle itself (no GPU): pins against public known answers and invariants.

The reference holds NO golden vectors: what can be pinned is pinned here."""
import hashlib
import os

import pytest

from oracle import ipp, r1cs
from oracle.curves import BLS12_381, BN254
from oracle.merlin import Transcript, shake256

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_keccak_matches_hashlib():
    for m in [b"", b"abc", b"x" * 135, b"y" * 136, b"z" * 500]:
        assert shake256(m, 200) == hashlib.shake_256(m).digest(200)


def test_merlin_equivalence_simple_kat():
    # merlin 1.x tests::equivalence_simple (public known answer)
    t = Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


@pytest.mark.parametrize("C", [BLS12_381, BN254])
def test_curve_constants(C):
    G = C.from_affine(C.g)
    assert C.on_curve(C.g)
    assert C.is_inf(C.mul(G, C.r))
    assert not C.is_inf(C.mul(G, C.r - 1))
    assert C.to_affine(C.add(C.mul(G, C.r - 1), G)) is None
    # identity encoding = AMCL (0, 1)
    assert C.g1_to_bytes(C.INF) == b"\x04" + bytes(C.MODBYTES) + bytes(C.MODBYTES - 1) + b"\x01"
    assert C.is_inf(C.g1_from_bytes(C.g1_to_bytes(C.INF)))


def test_bls_generator_small_multiple_kat():
    # 2*G1 of BLS12-381 (public constant, e.g. in the IETF pairing-friendly-curves draft test data)
    C = BLS12_381
    x2 = 0x0572CBEA904D67468808C8EB50A9450C9721DB309128012543902D0AC358A62AE28F75BB8F1C7C42C39A8C5529BF0F4E
    y2 = 0x166A9D8CABC673A322FDA673779D8E3822BA3ECB8670E461F73BB9021D5FD76A4C56D9D4CD16BD1BBA86881979749D28
    assert C.to_affine(C.dbl(C.from_affine(C.g))) == (x2, y2)


@pytest.mark.parametrize("C", [BLS12_381, BN254])
def test_hash_to_curve_properties(C):
    for msg in (b"g", b"h", b"Q", b"G1", b"H17"):
        P = C.g1_from_msg_hash(msg)
        assert C.on_curve(C.to_affine(P)) and C.is_inf(C.mul(P, C.r)) and not C.is_inf(P)
    gens = C.get_generators("G", 3)
    assert [C.to_affine(p) for p in gens] == [C.to_affine(C.g1_from_msg_hash(b"G%d" % i)) for i in (1, 2, 3)]


@pytest.mark.parametrize("C", [BLS12_381, BN254])
def test_msm_linearity(C):
    G = C.from_affine(C.g)
    pts = [C.mul(G, k) for k in (3, 5, 7, 11, 13)]
    s = C.synth_scalars(4, 5)
    assert C.eq(C.msm(pts, s), C.mul(G, sum(a * b for a, b in zip(s, (3, 5, 7, 11, 13))) % C.r))
    with pytest.raises(ValueError):
        C.msm(pts, s[:4])


@pytest.mark.parametrize("C", [BLS12_381, BN254])
def test_ipp_roundtrip_and_padding(C):
    # ipp.rs:325-390 (test_ipp) and ipp.rs:393-489 (test_ipp_non_power_of_2)
    for a, b in (([1, 2, 3, 4], [5, 6, 7, 8]), ([1, 2, 3, 4, 9, 0, 0, 0], [5, 6, 7, 8, 10, 0, 0, 0])):
        n = len(a)
        G, H, Q = C.get_generators("g", n), C.get_generators("h", n), C.g1_from_msg_hash(b"Q")
        Gf, Hf = [1] * n, C.vandermonde(C.synth_scalar(9, 0), n)
        pr = ipp.create_ipp(C, Transcript(b"innerproduct", C), Q, Gf, Hf, G, H, a, b)
        assert len(pr.L) == len(pr.R) == n.bit_length() - 1
        k = 5 if n == 8 else 4          # P built from the un-padded vectors
        bp_ = [x * y % C.r for x, y in zip(b[:k], Hf)]
        P = C.msm(G[:k] + H[:k] + [Q], a[:k] + bp_ + [C.inner_product(a, b)])
        ipp.verify_ipp(C, n, Transcript(b"innerproduct", C), Gf, Hf, P, Q, G, H, pr.a, pr.b, pr.L, pr.R)
        with pytest.raises(ipp.VerificationError):
            ipp.verify_ipp(C, n, Transcript(b"innerproduct", C), Gf, Hf, P, Q, G, H, (pr.a + 1) % C.r, pr.b, pr.L, pr.R)
    with pytest.raises(ipp.VerificationError):       # ipp.rs:274-276
        ipp.verification_scalars(C, pr.L, pr.R, 4, Transcript(b"innerproduct", C))


@pytest.mark.parametrize("C", [BLS12_381, BN254])
def test_r1cs_bound_check_roundtrip(C):
    # gadgets/bound_check.rs:188-225 at 8 bits
    g, h = C.g1_from_msg_hash(b"g"), C.g1_from_msg_hash(b"h")
    bits = 8
    G, H = C.get_generators("G", 2 * bits), C.get_generators("H", 2 * bits)
    rng = r1cs.make_rng(C, 1)
    p = r1cs.Prover(C, g, h, Transcript(b"BoundChecks", C))
    comms = r1cs.prove_bounded_num(p, 75, rng(), 10, 100, bits, rng)
    proof = p.prove(G, H, rng)
    v = r1cs.Verifier(C, Transcript(b"BoundChecks", C))
    r1cs.verify_bounded_num(v, 10, 100, bits, comms)
    v.verify(proof, g, h, G, H, rng())
    proof.t_x = (proof.t_x + 1) % C.r
    v = r1cs.Verifier(C, Transcript(b"BoundChecks", C))
    r1cs.verify_bounded_num(v, 10, 100, bits, comms)
    with pytest.raises(r1cs.R1CSError):
        v.verify(proof, g, h, G, H, rng())
    with pytest.raises(r1cs.InvalidGeneratorsLength):
        p2 = r1cs.Prover(C, g, h, Transcript(b"BoundChecks", C))
        r1cs.prove_bounded_num(p2, 75, rng(), 10, 100, bits, rng)
        p2.prove(G[:5], H[:5], rng)


def test_c_oracle_matches_python():
    from oracle import cref
    for C in (BLS12_381, BN254):
        G = C.from_affine(C.g)
        n = 60
        xy = cref.multiples(C.id, C.g1_xy_bytes(C.mul(G, 5)), n)
        m = 2 * C.MODBYTES
        pts = [C.g1_from_xy_bytes(xy[i * m:(i + 1) * m]) for i in range(n)]
        assert C.to_affine(pts[7]) == C.to_affine(C.mul(G, 40))
        s = C.synth_scalars(1, n)
        s[0], s[1], s[2] = 0, 1, C.r - 1
        sb = b"".join(C.fr_to_bytes(x) for x in s)
        exp = C.g1_xy_bytes(C.msm(pts, s))
        assert cref.msm(C.id, xy, sb, n, 1) == exp
        assert cref.msm(C.id, xy, sb, n, 4) == exp
        assert cref.msm(C.id, xy, sb, 0, 1) == C.g1_xy_bytes(C.INF)


def test_two_phase_circuit_round_trip():
    """Randomised constraints (prover.rs:300-319,384-436; verifier.rs:245-264) in the oracle: a 3-shuffle with a 4-bit range
    check in the first phase proves and verifies (n = 4 + 4 = 8), the second phase commits something, and a
    non-permutation is rejected.  Group operations through oracle/c to keep the CPU suite fast."""
    from oracle import fast, r1cs as or1cs
    from oracle.curves import BN254
    from oracle.merlin import Transcript
    C = fast.FastCurve(BN254)
    with fast.c_keccak():
        g, h = C.g1_from_msg_hash(b"g"), C.g1_from_msg_hash(b"h")
        k, bits = 3, 4
        G, H = C.get_generators("G", 8), C.get_generators("H", 8)
        xs = [7, 20, 33]
        for ys, ok in (([20, 33, 7], True), ([21, 33, 7], False)):
            rng = or1cs.make_rng(C, 5)
            p = or1cs.Prover(C, g, h, Transcript(b"S", C))
            comms, vs = [], []
            for v in xs + ys:
                com, var = p.commit(v, rng())
                comms.append(com)
                vs.append(var)
            or1cs.positive_no_gadget(p, or1cs.AllocatedQuantity(vs[0], xs[0]), bits)
            or1cs.shuffle_gadget(p, vs[:k], vs[k:])
            proof = p.prove(G, H, rng)
            assert len(p.a_L) == 8 and not C.is_inf(proof.A_I2) and not C.is_inf(proof.S2)
            v = or1cs.Verifier(C, Transcript(b"S", C))
            vv = [v.commit(c) for c in comms]
            or1cs.positive_no_gadget(v, or1cs.AllocatedQuantity(vv[0], None), bits)
            or1cs.shuffle_gadget(v, vv[:k], vv[k:])
            if ok:
                v.verify(proof, g, h, G, H, 12345)
            else:
                with pytest.raises(or1cs.R1CSError):
                    v.verify(proof, g, h, G, H, 12345)
