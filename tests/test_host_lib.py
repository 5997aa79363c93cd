"""CPU tests of the host layer's transcript (bulletproofs-amcl_b200/host/merlin.hpp, curve.hpp) through
include/bphost.h: Merlin's published known answer, and TranscriptProtocol (transcript.rs:29-61) against the
oracle's restatement, including the 48-byte challenge reduced mod r on BLS12-381."""
import ctypes

import pytest

from oracle.curves import BLS12_381, BN254
from oracle.merlin import Transcript


def test_merlin_equivalence_simple_kat(bp):
    """merlin's own test `equivalence_simple`: new("test protocol"), append("some label","some data"),
    32 challenge bytes under "challenge"."""
    out = ctypes.create_string_buffer(32)
    rc = bp.lib().bph_merlin_kat(b"test protocol", b"some label", b"some data", 9, b"challenge", out, 32)
    assert rc == 0
    assert out.raw.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_merlin_long_messages_cross_rate_boundary(bp):
    """messages longer than the STROBE rate (166) exercise run_f inside absorb / squeeze."""
    msg = bytes(range(256)) * 3
    out = ctypes.create_string_buffer(400)
    assert bp.lib().bph_merlin_kat(b"L", b"m", msg, len(msg), b"c", out, 400) == 0
    t = Transcript(b"L")
    t.append_message(b"m", msg)
    assert out.raw == t.challenge_bytes(b"c", 400)


@pytest.mark.parametrize("C", [BLS12_381, BN254], ids=lambda c: c.name)
def test_transcript_protocol_matches_oracle(bp, C):
    P = C.mul(C.from_affine(C.g), 123456789)
    for s in (0, 1, C.r - 1, C.synth_scalar(9, 0)):
        out = ctypes.create_string_buffer(C.MODBYTES)
        rc = bp.lib().bph_transcript_kat(C.id, b"kat", C.g1_xy_bytes(P), C.fr_to_bytes(s), out)
        assert rc == 0
        t = Transcript(b"kat", C)
        t.innerproduct_domain_sep(8)
        t.commit_point(b"P", P)
        t.commit_scalar(b"s", s)
        assert out.raw == C.fr_to_bytes(t.challenge_scalar(b"c"))
    # the identity point is absorbed as 04 || 0..0 || 0..01 (prover.rs:429-434)
    out = ctypes.create_string_buffer(C.MODBYTES)
    assert bp.lib().bph_transcript_kat(C.id, b"kat", C.g1_xy_bytes(C.INF), C.fr_to_bytes(5), out) == 0
    t = Transcript(b"kat", C)
    t.innerproduct_domain_sep(8)
    t.commit_point(b"P", C.INF)
    t.commit_scalar(b"s", 5)
    assert out.raw == C.fr_to_bytes(t.challenge_scalar(b"c"))


@pytest.mark.parametrize("C", [BLS12_381, BN254], ids=lambda c: c.name)
def test_host_g1_sum_is_the_combine_step(bp, C):
    """bph_g1_sum: the host-side addition of per-shard partial sums, incl. identity entries, a repeated
    point (P + P), an inverse pair and the empty list."""
    G = C.from_affine(C.g)
    pts = [C.mul(G, k) for k in (5, 11, 11, 1234567)] + [C.INF, C.neg(C.mul(G, 5))]
    exp = C.INF
    for P in pts:
        exp = C.add(exp, P)
    xy = b"".join(C.g1_xy_bytes(P) for P in pts)
    cid = bp.BLS12_381 if C.id == 0 else bp.BN254
    assert bp.g1_sum(cid, xy) == C.g1_xy_bytes(exp)
    assert bp.g1_sum(cid, b"") == C.g1_xy_bytes(C.INF)
    assert bp.g1_sum(cid, C.g1_xy_bytes(G) + C.g1_xy_bytes(C.neg(G))) == C.g1_xy_bytes(C.INF)
