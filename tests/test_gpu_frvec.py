"""GPU parity: FieldElementVector algebra kernels vs the oracle's integers (bit exact).
Reference call sites: ipp.rs:77-82,295; prover.rs:463,472-485,513; verifier.rs:342-352,416;
vector_poly.rs:79-97."""
import random

import pytest

from tests.util import curve_of, dec_scalars, enc_scalars

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", ["bls", "bn"])
@pytest.mark.parametrize("n", [0, 1, 5, 64, 1000, 4097])
def test_fr_vector_ops(which, n, ctx_bls, ctx_bn):
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    r = C.r
    rnd = random.Random(n)
    a = [rnd.randrange(r) for _ in range(n)]
    b = [rnd.randrange(r) for _ in range(n)]
    if n > 3:
        a[0], a[1], a[2] = 0, 1, r - 1
        b[0], b[1], b[2] = r - 1, r - 1, r - 1
    da, db = ctx.upload_scalars(enc_scalars(C, a)), ctx.upload_scalars(enc_scalars(C, b))
    assert dec_scalars(C, ctx.fr_hadamard(da, db).download()) == [x * y % r for x, y in zip(a, b)]
    assert dec_scalars(C, ctx.fr_add(da, db).download()) == [(x + y) % r for x, y in zip(a, b)]
    assert dec_scalars(C, ctx.fr_sub(da, db).download()) == [(x - y) % r for x, y in zip(a, b)]
    s = rnd.randrange(r)
    assert dec_scalars(C, ctx.fr_scale(da, C.fr_to_bytes(s)).download()) == [x * s % r for x in a]
    assert int.from_bytes(ctx.fr_inner_product(da, db), "big") == C.inner_product(a, b)
    x = rnd.randrange(r)
    assert dec_scalars(C, ctx.fr_vandermonde(C.fr_to_bytes(x), n).download()) == C.vandermonde(x, n)
    if n >= 4:
        # offsets: <a[1..3), b[2..4)>
        assert int.from_bytes(ctx.fr_inner_product(da, db, n=2, aoff=1, boff=2), "big") == (a[1] * b[2] + a[2] * b[3]) % r
        with pytest.raises(Exception):
            ctx.fr_inner_product(da, db, n=n, aoff=1, boff=0)     # UnequalSizeVectors analogue


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_special_inner_product_and_batch_invert(which, ctx_bls, ctx_bn):
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    r = C.r
    n = 777
    rnd = random.Random(1)
    vs = [[rnd.randrange(r) for _ in range(n)] for _ in range(6)]
    dv = [ctx.upload_scalars(enc_scalars(C, v)) for v in vs]
    l1, l2, l3, r0, r1, r3 = vs
    ip = C.inner_product
    exp = [ip(l1, r0), (ip(l1, r1) + ip(l2, r0)) % r, (ip(l2, r1) + ip(l3, r0)) % r, (ip(l1, r3) + ip(l3, r1)) % r,
           ip(l2, r3), ip(l3, r3)]                      # vector_poly.rs:82-87
    assert dec_scalars(C, ctx.fr_poly3_special_inner_product(*dv, n)) == exp
    ch = [rnd.randrange(1, r) for _ in range(15)] + [0]
    inv, prod = ctx.fr_batch_invert(ctx.upload_scalars(enc_scalars(C, ch)))
    einv, eprod = C.fr_batch_invert(ch)
    assert dec_scalars(C, inv.download()) == einv and int.from_bytes(prod, "big") == eprod


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_device_random_vector_is_the_host_stream(which, ctx_bls, ctx_bn):
    """bpgpu_fr_random (FieldElementVector::random on the device, prover.rs:340-341): element i is
    SHAKE256(key || le64(ctr0 + i)) reduced mod r -- for key = seed_le64 || "blind" that is the oracle's blinding stream
    (oracle/r1cs.py make_rng), including the 384 -> 255 bit reduction on BLS12-381."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    for seed, ctr0, n in ((1, 0, 300), (77, 5, 33), (2**63 + 5, 2**40, 4), (3, 0, 0)):
        key = (seed % 2**64).to_bytes(8, "little") + b"blind"
        v = ctx.fr_random(key, ctr0, n)
        assert len(v) == n
        exp = [C.synth_scalar(seed % 2**64, ctr0 + i, b"blind") for i in range(n)]
        assert v.download() == enc_scalars(C, exp)
        v.free()
    # any key up to 64 bytes; distinct keys give distinct streams
    a, b = ctx.fr_random(b"k" * 64, 0, 8), ctx.fr_random(b"k" * 63 + b"j", 0, 8)
    assert a.download() != b.download()
    with pytest.raises(Exception):
        ctx.fr_random(b"k" * 65, 0, 1)


def test_scalar_views_share_storage(ctx_bls):
    """bpgpu_scalars_view: handles onto slices of one allocation, freed in any order."""
    ctx = ctx_bls
    C = curve_of(ctx)
    vals = C.synth_scalars(12, 10)
    s = ctx.upload_scalars(enc_scalars(C, vals))
    a, b = s.view(0, 4), s.view(4, 6)
    assert len(a) == 4 and len(b) == 6
    s.free()                                            # the views keep the storage alive
    assert dec_scalars(C, a.download()) == vals[:4]
    c = b.view(1, 2)
    b.free()
    assert dec_scalars(C, c.download()) == vals[5:7]
    out = ctx.fr_hadamard(c, c)
    assert dec_scalars(C, out.download()) == [v * v % C.r for v in vals[5:7]]
    a.free()
    c.free()
    out.free()
    with pytest.raises(Exception):
        ctx.upload_scalars(enc_scalars(C, vals)).view(8, 3)


@pytest.mark.parametrize("which,m,bits", [("bls", 2, 8), ("bn", 3, 5), ("bls", 1, 64)])
def test_prover_and_verifier_vector_kernels_match_oracle(which, m, bits, ctx_bls, ctx_bn):
    """Rows a11 / a12 element by element (round 1 covered them only through whole-proof bytes and verdicts):
    bpgpu_r1cs_prover_polys / _eval against the oracle prover's l1, r0, r1, r3, l_vec, r_vec, G_factors, H_factors
    (prover.rs:458-486, 524-535, 552-563) and bpgpu_r1cs_verifier_scalars against the oracle verifier's g_scalars,
    h_scalars and delta (verifier.rs:341-390), on a real range-proof witness incl. a padded circuit (n = 15 -> 16)."""
    from oracle import r1cs as or1cs
    from oracle.merlin import Transcript
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    g, h = C.g1_from_msg_hash(b"g"), C.g1_from_msg_hash(b"h")
    n = m * bits
    N = 1 << max(0, (n - 1).bit_length())
    G, H = C.get_generators("G", N), C.get_generators("H", N)
    rng = or1cs.make_rng(C, 91)
    vals = [C.synth_scalar(91, j, b"v") & ((1 << bits) - 1) for j in range(m)]
    p = or1cs.Prover(C, g, h, Transcript(b"Vec", C))
    comms = []
    for v in vals:
        com, var = p.commit(v, rng())
        comms.append(com)
        or1cs.positive_no_gadget(p, or1cs.AllocatedQuantity(var, v), bits)
    p.trace = {}
    proof = p.prove(G, H, rng)
    t = p.trace
    up = lambda v: ctx.upload_scalars(enc_scalars(C, v))
    fb = C.fr_to_bytes
    l1, r0, r1, r3 = ctx.r1cs_prover_polys(n, up(p.a_L), up(p.a_R), up(t["s_R"]), up(t["wL"]), up(t["wR"]), up(t["wO"]), fb(t["y"]))
    for got, key in ((l1, "l1"), (r0, "r0"), (r1, "r1"), (r3, "r3")):
        assert dec_scalars(C, got.download()) == t[key], key
    assert dec_scalars(C, ctx.fr_poly3_special_inner_product(l1, up(p.a_O), up(t["s_L"]), r0, r1, r3, n)) == t["t"]
    lv, rv, gf, hf = ctx.r1cs_prover_eval(n, t["n1"], N, l1, up(p.a_O), up(t["s_L"]), r0, r1, r3, fb(t["x"]), fb(t["u"]), fb(t["y"]))
    for got, key in ((lv, "l_vec"), (rv, "r_vec"), (gf, "G_factors"), (hf, "H_factors")):
        assert dec_scalars(C, got.download()) == t[key], key
    # verifier side
    v_ = or1cs.Verifier(C, Transcript(b"Vec", C))
    for com in comms:
        or1cs.positive_no_gadget(v_, or1cs.AllocatedQuantity(v_.commit(com), None), bits)
    v_.trace = {}
    v_.verify(proof, g, h, G, H, 12345)
    tv = v_.trace
    assert tv["y"] == t["y"] and tv["wL"] == t["wL"]
    s = ctx.upload_scalars(enc_scalars(C, tv["s"]))
    gh, delta = ctx.r1cs_verifier_scalars(n, tv["n1"], N, up(tv["wL"]), up(tv["wR"]), up(tv["wO"]), s, fb(tv["y"]), fb(tv["x"]),
                                          fb(tv["a"]), fb(tv["b"]), fb(tv["u"]))
    assert dec_scalars(C, gh.download()) == tv["g_scalars"] + tv["h_scalars"]
    assert int.from_bytes(delta, "big") == tv["delta"]


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_mimc_witness_generation(which, ctx_bls, ctx_bn):
    """batched MiMC hashing and the multiplier assignments of its gadget on the device, against the
    oracle's mimc() and against the a_L / a_R / a_O an oracle PROVER allocates when it runs enforce_mimc_2_inputs
    (helper_constraints/mimc.rs:10-29, 53-77) -- 10 rounds here, the reference's tests use 322."""
    from oracle import r1cs as or1cs
    from oracle.merlin import Transcript
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    rounds, count = 10, 37
    consts = C.synth_scalars(5, rounds, b"mimc")
    xl, xr = C.synth_scalars(6, count, b"xl"), C.synth_scalars(6, count, b"xr")
    xl[0], xr[0], xl[1] = 0, 0, C.r - 1
    up = lambda v: ctx.upload_scalars(enc_scalars(C, v))
    img, aL, aR, aO = ctx.mimc_witness(up(xl), up(xr), up(consts), rounds)
    assert dec_scalars(C, img.download()) == [or1cs.mimc(C, a, b, consts, rounds) for a, b in zip(xl, xr)]
    gl, gr, go = dec_scalars(C, aL.download()), dec_scalars(C, aR.download()), dec_scalars(C, aO.download())
    g, h = C.g1_from_msg_hash(b"g"), C.g1_from_msg_hash(b"h")
    for i in (0, 1, 5, count - 1):
        p = or1cs.Prover(C, g, h, Transcript(b"MiMC", C))
        _, vl = p.commit(xl[i], 11)
        _, vr = p.commit(xr[i], 12)
        or1cs.enforce_mimc_2_inputs(p, or1cs.LC.of(vl, C), or1cs.LC.of(vr, C), rounds, consts)
        w = 2 * rounds
        assert (gl[i * w:(i + 1) * w], gr[i * w:(i + 1) * w], go[i * w:(i + 1) * w]) == (p.a_L, p.a_R, p.a_O), i
    img2, _, _, _ = ctx.mimc_witness(up(xl), up(xr), up(consts), rounds, with_multipliers=False)
    assert img2.download() == img.download()
    with pytest.raises(Exception):
        ctx.mimc_witness(up(xl), up(xr), up(consts[:3]), rounds)


@pytest.mark.parametrize("which", ["bls", "bn"])
@pytest.mark.parametrize("width,sbox", [(3, 0), (5, 1), (9, 2), (3, 1)])
def test_poseidon_permutation(which, width, sbox, ctx_bls, ctx_bn):
    """Poseidon_permutation (poseidon.rs:202-293) batched on the device against the oracle's restatement: all three S-boxes,
    the three widths, full and partial rounds, a zero state element under the inverse S-box (0 -> 0)"""
    from oracle import r1cs as or1cs
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    frb, pr, fre, count = 2, 5, 2, 9
    keys = C.synth_scalars(21, (frb + pr + fre) * width, b"rk")
    mds = [[C.synth_scalar(22, j * width + i, b"mds") for i in range(width)] for j in range(width)]
    states = [[C.synth_scalar(23, b * width + i, b"in") for i in range(width)] for b in range(count)]
    states[0] = [0] * width
    states[1][width - 1] = (C.r - keys[width - 1]) % C.r            # the S-box input of the last element is 0 in round 1
    up = lambda v: ctx.upload_scalars(enc_scalars(C, v))
    out = ctx.poseidon_permutation(up([x for s in states for x in s]), count, width, frb, pr, fre, sbox, up(keys),
                                   up([mds[j][i] for j in range(width) for i in range(width)]))
    name = ["cube", "inverse", "quint"][sbox]
    exp = [x for s in states for x in or1cs.poseidon_permutation(C, s, width, frb, pr, fre, keys, mds, name)]
    assert dec_scalars(C, out.download()) == exp
    with pytest.raises(Exception):
        ctx.poseidon_permutation(up(states[0]), 1, 4, frb, pr, fre, sbox, up(keys), up([1] * 16))     # width 4: not supported
