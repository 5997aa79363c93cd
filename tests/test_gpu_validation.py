"""Untrusted inputs at the boundary: every externally supplied point must be a point as AMCL's
ECP::frombytes / ECP::new_bigs decide it (coordinates < p, on the curve, or the identity (0, 1)), every proof scalar
canonical.  Anything else is BPGPU_E_FORMAT (-5) -- never silently reduced, never fed to the a = 0 group law."""
import pytest

from tests.util import curve_of, enc_points, enc_scalars, rand_points

pytestmark = pytest.mark.gpu


def _bad_points(C):
    mb = C.MODBYTES
    gx, gy = C.g
    off_curve = gx.to_bytes(mb, "big") + ((gy + 1) % C.p).to_bytes(mb, "big")
    zero_zero = bytes(2 * mb)
    out = [off_curve, zero_zero]
    if gx + C.p < (1 << (8 * mb)):
        out.append((gx + C.p).to_bytes(mb, "big") + gy.to_bytes(mb, "big"))      # same residue, not canonical
    out.append(gx.to_bytes(mb, "big") + (C.p).to_bytes(mb, "big"))               # y = p
    return out


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_raw_abi_rejects_invalid_points(bp, ctx_bls, ctx_bn, which):
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    pts = rand_points(C, 5, 3)
    good = enc_points(C, pts)
    sc = enc_scalars(C, C.synth_scalars(3, 5))
    exp = C.g1_xy_bytes(C.msm(pts, C.synth_scalars(3, 5)))
    assert ctx.msm_refs(good, sc) == exp
    P = 2 * C.MODBYTES
    for bad in _bad_points(C):
        mixed = good[:2 * P] + bad + good[3 * P:]
        with pytest.raises(bp.BpgpuError) as e:
            ctx.upload_points(mixed)
        assert e.value.code == -5
        with pytest.raises(bp.BpgpuError) as e:
            ctx.msm_refs(mixed, sc)
        assert e.value.code == -5
        with pytest.raises(bp.BpgpuError) as e:
            ctx.commit_batch(bad + good[:P], sc[:2 * C.MODBYTES], 1)
        assert e.value.code == -5
        # the flag does not leak into the next call
        assert ctx.msm_refs(good, sc) == exp
    # the identity and the negation of a point are points
    ident = C.g1_xy_bytes(C.INF)
    assert ctx.msm_refs(good[:2 * P] + ident + good[3 * P:], sc) == C.g1_xy_bytes(C.msm(pts[:2] + pts[3:], C.synth_scalars(3, 5)[:2] + C.synth_scalars(3, 5)[3:]))


def test_scalars_are_reduced_not_truncated(ctx_bls):
    """48-byte scalars on BLS12-381: FieldElement::from(&[u8; 48]) reduces the whole integer mod r (round 1 read only the low 32 bytes)"""
    C = curve_of(ctx_bls)
    pts = rand_points(C, 3, 8)
    vals = [(1 << 380) + 12345, C.r + 7, (1 << 256) + 1]
    sc = b"".join(v.to_bytes(48, "big") for v in vals)
    assert ctx_bls.msm_refs(enc_points(C, pts), sc) == C.g1_xy_bytes(C.msm(pts, [v % C.r for v in vals]))
    d = ctx_bls.upload_scalars(sc)
    assert d.download() == enc_scalars(C, [v % C.r for v in vals])
    d.free()


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_single_verifier_rejects_malformed_proofs(ctx_bls, ctx_bn, which):
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    mb = C.MODBYTES
    PB = 1 + 2 * mb
    bits, m = 8, 2
    dG, dH = ctx.get_generators("G", 16), ctx.get_generators("H", 16)
    gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
    proof, comms = ctx.range_prove(b"V", gx, hx, dG, dH, [7, 200], bits, seed=11)
    assert ctx.range_verify(b"V", gx, hx, dG, dH, m, bits, proof, comms) is True

    def code(p=proof, c=comms, g=gx):
        try:
            return 0 if ctx.range_verify(b"V", g, hx, dG, dH, m, bits, p, c) else -4
        except Exception as e:          # BpgpuError
            return e.code
    for bad in _bad_points(C):
        for k in (0, 7, 11 + 2):                           # A_I1, T_3, L_3 (after the three scalars)
            off = k * PB + (3 * mb if k >= 11 else 0)
            assert code(p=proof[:off + 1] + bad + proof[off + PB:]) == -5
        assert code(c=bad + comms[2 * mb:]) == -5
        assert code(g=bad) == -5
    # scalars: t_x = r and a = r + 1 are not canonical
    o = 11 * PB
    assert code(p=proof[:o] + C.r.to_bytes(mb, "big") + proof[o + mb:]) == -5
    assert code(p=proof[:-2 * mb] + (C.r + 1).to_bytes(mb, "big") + proof[-mb:]) == -5
    # a well-formed but wrong scalar is a verification error, not a format error
    assert code(p=proof[:o] + (5).to_bytes(mb, "big") + proof[o + mb:]) == -4


def test_ipp_verifier_rejects_malformed_proofs(ctx_bls):
    ctx = ctx_bls
    C = curve_of(ctx)
    n = 8
    dG, dH = ctx.get_generators("g", n), ctx.get_generators("h", n)
    Q = ctx.g1_from_msg_hash(b"Q")
    a, b = C.synth_scalars(1, n, b"a"), C.synth_scalars(1, n, b"b")
    ones = enc_scalars(C, [1] * n)
    proof = ctx.ipp_create(b"ipp", dG, dH, Q, ones, ones, enc_scalars(C, a), enc_scalars(C, b), n)
    G, H = C.get_generators("g", n), C.get_generators("h", n)
    P = C.msm(G + H + [C.g1_from_xy_bytes(Q)], a + b + [C.inner_product(a, b)])
    Pb = C.g1_xy_bytes(P)
    assert ctx.ipp_verify(b"ipp", n, ones, ones, Pb, Q, dG, dH, proof)
    mb = 48
    for bad in _bad_points(C):
        for args in ((bad, Q, proof), (Pb, bad, proof), (Pb, Q, proof[:1] + bad + proof[97:])):
            with pytest.raises(Exception) as e:
                ctx.ipp_verify(b"ipp", n, ones, ones, args[0], args[1], dG, dH, args[2])
            assert e.value.code == -5
    with pytest.raises(Exception) as e:                     # b >= r (round 1 accepted it reduced)
        ctx.ipp_verify(b"ipp", n, ones, ones, Pb, Q, dG, dH, proof[:-mb] + (C.r + 3).to_bytes(mb, "big"))
    assert e.value.code == -5


def test_context_close_releases_outstanding_handles(bp):
    """freeing a handle after its context was a use-after-free in round 1: close() now frees what is still alive, and a
    later free() of the Python object is a no-op"""
    c = bp.Context(bp.BLS12_381, 0)
    C = curve_of(c)
    pts = c.upload_points(enc_points(C, rand_points(C, 4, 1)))
    sc = c.upload_scalars(enc_scalars(C, [1, 2, 3, 4]))
    view = sc.view(1, 2)
    c.close()
    assert pts.handle is None and sc.handle is None and view.handle is None
    pts.free()
    with bp.Context(bp.BN254, 0) as c2:
        assert c2.modbytes == 32
    assert c2.handle is None
