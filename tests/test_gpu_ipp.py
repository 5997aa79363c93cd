"""GPU parity: device-resident IPP rounds (bpgpu_ipp_*) vs the oracle's restatement of ipp.rs.

The test plays the host role of IPP::create_ipp (ipp.rs:35-202): Merlin transcript and challenges on
the host (here: the oracle's transcript), everything else through the C ABI.  Proof elements must be
byte-identical to the oracle's, including with non-trivial G_factors/H_factors as R1CS uses them
(prover.rs:552-563) and for the reference's own test shapes (ipp.rs:325-489)."""
import pytest

from oracle import ipp as oipp
from oracle.merlin import Transcript
from tests.util import curve_of, dec_scalars, enc_points, enc_scalars

pytestmark = pytest.mark.gpu


def gpu_create_ipp(ctx, C, transcript, Q, Gf, Hf, G, H, a, b):
    n = len(G)
    dG, dH = ctx.upload_points(enc_points(C, G)), ctx.upload_points(enc_points(C, H))
    dGf, dHf = ctx.upload_scalars(enc_scalars(C, Gf)), ctx.upload_scalars(enc_scalars(C, Hf))
    da, db = ctx.upload_scalars(enc_scalars(C, a)), ctx.upload_scalars(enc_scalars(C, b))
    transcript.innerproduct_domain_sep(n)
    st = ctx.ipp_begin(dG, dH, C.g1_xy_bytes(Q), dGf, dHf, da, db, n)
    Ls, Rs = [], []
    m = n
    while m != 1:
        L, R = ctx.ipp_round_LR(st)
        transcript.append_message(b"L", b"\x04" + L)
        transcript.append_message(b"R", b"\x04" + R)
        Ls.append(L)
        Rs.append(R)
        u = transcript.challenge_scalar(b"u")
        ctx.ipp_fold(st, C.fr_to_bytes(u), C.fr_to_bytes(C.fr_inv(u)))
        m //= 2
    a_out, b_out = ctx.ipp_finish(st)
    for d in (dG, dH, dGf, dHf, da, db):
        d.free()
    return Ls, Rs, a_out, b_out


@pytest.mark.parametrize("which", ["bls", "bn"])
@pytest.mark.parametrize("n", [1, 2, 4, 8, 64])
def test_create_ipp_matches_oracle(which, n, ctx_bls, ctx_bn):
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    G, H, Q = C.get_generators("g", n), C.get_generators("h", n), C.g1_from_msg_hash(b"Q")
    a, b = C.synth_scalars(3, n, b"a"), C.synth_scalars(3, n, b"b")
    if n == 4:
        a, b = [1, 2, 3, 4], [5, 6, 7, 8]                 # ipp.rs:326-335
    if n == 8:
        a, b = [1, 2, 3, 4, 9, 0, 0, 0], [5, 6, 7, 8, 10, 0, 0, 0]   # ipp.rs:395-402 (zero padded)
    u = C.synth_scalar(5, 0)
    Gf = [1] * (n // 2) + [u] * (n - n // 2)               # prover.rs:552-556 shape
    Hf = [y * g % C.r for y, g in zip(C.vandermonde(C.synth_scalar(5, 1), n), Gf)]
    exp = oipp.create_ipp(C, Transcript(b"innerproduct", C), Q, Gf, Hf, G, H, a, b)
    Ls, Rs, a_out, b_out = gpu_create_ipp(ctx, C, Transcript(b"innerproduct", C), Q, Gf, Hf, G, H, a, b)
    assert Ls == [C.g1_xy_bytes(p) for p in exp.L]
    assert Rs == [C.g1_xy_bytes(p) for p in exp.R]
    assert a_out == C.fr_to_bytes(exp.a) and b_out == C.fr_to_bytes(exp.b)

    # verifier side: s vector and expected_P through the ABI (ipp.rs:204-315)
    lg = len(exp.L)
    t = Transcript(b"innerproduct", C)
    u_sq, u_inv_sq, s = oipp.verification_scalars(C, exp.L, exp.R, n, t)
    t2 = Transcript(b"innerproduct", C)
    t2.innerproduct_domain_sep(n)
    ch = []
    for L, R in zip(exp.L, exp.R):
        t2.commit_point(b"L", L)
        t2.commit_point(b"R", R)
        ch.append(t2.challenge_scalar(b"u"))
    ds = ctx.ipp_verification_scalars(enc_scalars(C, ch), lg)
    assert dec_scalars(C, ds.download()) == s
    dG, dH = ctx.upload_points(enc_points(C, G)), ctx.upload_points(enc_points(C, H))
    dGf, dHf = ctx.upload_scalars(enc_scalars(C, Gf)), ctx.upload_scalars(enc_scalars(C, Hf))
    got_P = ctx.ipp_verify_msm(dG, dH, C.g1_xy_bytes(Q), dGf, dHf, a_out, b_out, enc_scalars(C, ch), b"".join(Ls), b"".join(Rs), lg)
    # P as the prover would commit it: <a*Gf, G> + <b*Hf, H> + <a,b> Q   (ipp.rs:354-372)
    P = C.msm(G + H + [Q], [x * f % C.r for x, f in zip(a, Gf)] + [x * f % C.r for x, f in zip(b, Hf)] + [C.inner_product(a, b)])
    assert got_P == C.g1_xy_bytes(P)
    # a tampered proof scalar must change expected_P (no negative test exists upstream)
    bad = ctx.ipp_verify_msm(dG, dH, C.g1_xy_bytes(Q), dGf, dHf, C.fr_to_bytes((exp.a + 1) % C.r), b_out, enc_scalars(C, ch),
                             b"".join(Ls), b"".join(Rs), lg)
    assert bad != got_P


def test_ipp_argument_errors(ctx_bls):
    ctx = ctx_bls
    C = curve_of(ctx)
    G = C.get_generators("g", 4)
    dG = ctx.upload_points(enc_points(C, G))
    v3, v4 = ctx.upload_scalars(enc_scalars(C, [1, 2, 3])), ctx.upload_scalars(enc_scalars(C, [1, 2, 3, 4]))
    Q = C.g1_xy_bytes(G[0])
    with pytest.raises(Exception):        # assert!(n.is_power_of_two())  ipp.rs:48
        ctx.ipp_begin(dG, dG, Q, v3, v3, v3, v3, 3)
    with pytest.raises(Exception):        # assert_eq!(a_vec.len(), n)     ipp.rs:52
        ctx.ipp_begin(dG, dG, Q, v4, v4, v3, v4, 4)


@pytest.mark.parametrize("which", ["bls", "bn"])
@pytest.mark.parametrize("n", [1, 2, 16, 64])
def test_ipp_with_tables_and_arbitrary_q(which, n, ctx_bls, ctx_bn):
    """bph_ipp_create / bph_ipp_verify (the reference's own test shape, ipp.rs:325-390) with generators that carry window
    tables and a Q that is NOT a multiple of a cached base: Q gets a table of its own and every round runs on the table
    path; the proof bytes equal the oracle's, also for Q = identity"""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    G, H = C.get_generators("g", n), C.get_generators("h", n)
    dG, dH = ctx.get_generators("g", n, precompute=True), ctx.get_generators("h", n, precompute=True)
    a, b = C.synth_scalars(8, n, b"a"), C.synth_scalars(8, n, b"b")
    ones = [1] * n
    for Q in (C.g1_from_msg_hash(b"Q"), C.INF):
        exp = oipp.create_ipp(C, Transcript(b"tab", C), Q, ones, ones, G, H, a, b)
        want = b"".join(C.g1_to_bytes(p) for p in exp.L) + b"".join(C.g1_to_bytes(p) for p in exp.R) + C.fr_to_bytes(exp.a) + C.fr_to_bytes(exp.b)
        got = ctx.ipp_create(b"tab", dG, dH, C.g1_xy_bytes(Q), enc_scalars(C, ones), enc_scalars(C, ones), enc_scalars(C, a), enc_scalars(C, b), n)
        assert got == want
        P = C.msm(G + H + [Q], a + b + [C.inner_product(a, b)])
        assert ctx.ipp_verify(b"tab", n, enc_scalars(C, ones), enc_scalars(C, ones), C.g1_xy_bytes(P), C.g1_xy_bytes(Q), dG, dH, got)
        assert not ctx.ipp_verify(b"tab", n, enc_scalars(C, ones), enc_scalars(C, ones), C.g1_xy_bytes(C.dbl(P)), C.g1_xy_bytes(Q), dG, dH, got)


@pytest.mark.parametrize("which", ["bls", "bn"])
@pytest.mark.parametrize("n,k", [(64, 1), (64, 3), (64, 5), (16, 3), (256, 2)])
def test_ipp_hybrid_materialised_generators(which, n, k, ctx_bls, ctx_bn, monkeypatch):
    """Large-N table mode materialises the folded generators after k rounds (csrc/ipp.cu "hybrid") and continues on the
    bucket pipeline over 2 n_k + 1 points.  Forced on at small n here: same L, R, a, b bytes as the oracle, with factor
    vectors that are not 1 (they are folded into the materialised points) and for Q = identity."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    monkeypatch.setenv("BPGPU_IPP_HYBRID", str(k))
    monkeypatch.setenv("BPGPU_IPP_HYBRID_MIN", "2")
    G, H = C.get_generators("g", n), C.get_generators("h", n)
    dG, dH = ctx.get_generators("g", n, precompute=True), ctx.get_generators("h", n, precompute=True)
    a, b = C.synth_scalars(18, n, b"a"), C.synth_scalars(18, n, b"b")
    gf, hf = C.synth_scalars(18, n, b"gf"), C.synth_scalars(18, n, b"hf")
    for Q in (C.g1_from_msg_hash(b"Q"), C.INF):
        exp = oipp.create_ipp(C, Transcript(b"hyb", C), Q, gf, hf, G, H, a, b)
        want = b"".join(C.g1_to_bytes(p) for p in exp.L) + b"".join(C.g1_to_bytes(p) for p in exp.R) + C.fr_to_bytes(exp.a) + C.fr_to_bytes(exp.b)
        launches0 = ctx.launches
        got = ctx.ipp_create(b"hyb", dG, dH, C.g1_xy_bytes(Q), enc_scalars(C, gf), enc_scalars(C, hf), enc_scalars(C, a), enc_scalars(C, b), n)
        assert got == want
        assert ctx.launches > launches0
    # and the switch is what changed the path: with it off the same bytes come from table sums alone
    monkeypatch.setenv("BPGPU_IPP_HYBRID", "0")
    assert ctx.ipp_create(b"hyb", dG, dH, C.g1_xy_bytes(Q), enc_scalars(C, gf), enc_scalars(C, hf), enc_scalars(C, a), enc_scalars(C, b), n) == want
