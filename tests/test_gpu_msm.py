"""GPU parity: bpgpu_msm* vs the oracle's sum_i s_i*P_i (bit exact affine bytes).

Reference call sites replaced: ipp.rs:91,104,158,170,251-253; verifier.rs:451; prover.rs:347-362.
Edge cases follow SURVEY.md 8d: empty, single, all-zero / all-one / all r-1 scalars, 0/1 witness
scalars (positive_no.rs:18-24), u8 scalars (ipp.rs:326-335), repeated points, P with -P, identity
among the inputs (prover.rs:429; poseidon_hash.rs:49-53)."""
import random

import pytest

from tests.util import curve_of, dec_points, enc_points, enc_scalars, rand_points

pytestmark = pytest.mark.gpu


def _check(ctx, C, P, s):
    exp = C.g1_xy_bytes(C.msm(P, s))
    ep, es = enc_points(C, P), enc_scalars(C, s)
    assert ctx.msm_refs(ep, es) == exp
    dp = ctx.upload_points(ep)
    assert ctx.msm(dp, es) == exp
    ds = ctx.upload_scalars(es)
    assert ctx.msm_device(dp, ds) == exp
    dp.free()
    ds.free()


@pytest.mark.parametrize("which", ["bls", "bn"])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 64, 141, 1025])
def test_msm_random(which, n, ctx_bls, ctx_bn):
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    P = rand_points(C, n, 100 + n)
    s = C.synth_scalars(1, n)
    _check(ctx, C, P, s)


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_msm_edge_scalars(which, ctx_bls, ctx_bn):
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    n = 96
    P = rand_points(C, n, 7)
    rnd = random.Random(3)
    for s in ([0] * n, [1] * n, [C.r - 1] * n, [rnd.randrange(2) for _ in range(n)],
              [rnd.randrange(256) for _ in range(n)], [(1 << 255) % C.r] * n,
              [C.r - 1 - rnd.randrange(4) for _ in range(n)]):
        _check(ctx, C, P, s)


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_msm_edge_points(which, ctx_bls, ctx_bn):
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    base = rand_points(C, 8, 9)
    # repeated point, P and -P, identity among the inputs
    P = [base[0]] * 20 + [base[1], C.neg(base[1])] * 5 + [C.INF, base[2], C.INF, base[3]] + base
    s = C.synth_scalars(2, len(P))
    _check(ctx, C, P, s)
    # same scalar on P and -P cancels exactly; whole sum is the identity
    P2 = [base[4], C.neg(base[4]), base[5], C.neg(base[5])]
    _check(ctx, C, P2, [77, 77, C.r - 5, C.r - 5])
    # offsets into a resident table
    ep = enc_points(C, base)
    dp = ctx.upload_points(ep)
    s3 = C.synth_scalars(3, 5)
    assert ctx.msm(dp, enc_scalars(C, s3), off=2, n=5) == C.g1_xy_bytes(C.msm(base[2:7], s3))
    with pytest.raises(Exception):
        ctx.msm(dp, enc_scalars(C, s3 * 2), off=2, n=10)      # UnequalSizeVectors analogue
    dp.free()


@pytest.mark.parametrize("which,n", [("bls", 1 << 14), ("bn", 1 << 13)])
def test_msm_medium_vs_oracle(which, n, ctx_bls, ctx_bn):
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    P = rand_points(C, n, 21)
    s = C.synth_scalars(1, n)
    assert ctx.msm_refs(enc_points(C, P), enc_scalars(C, s)) == C.g1_xy_bytes(C.msm(P, s))


def _device_multiples(ctx, C, ks):
    """P_i = k_i * G computed ON THE DEVICE (selftest_group op 2, itself oracle-checked above)."""
    n = len(ks)
    g = C.g1_xy_bytes(C.from_affine(C.g))
    return ctx.selftest_group(2, g * n, g * n, enc_scalars(C, ks))


@pytest.mark.parametrize("which,lg", [("bls", 16), ("bls", 20), ("bn", 18)])
def test_msm_large_structured(which, lg, ctx_bls, ctx_bn):
    """Size-independent property at BASELINE.json sizes: sum s_i*(k_i*G) == (sum s_i*k_i mod r)*G,
    with 0/1-heavy and uniform scalar mixes."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    n = 1 << lg
    rnd = random.Random(lg)
    ks = [rnd.randrange(1, C.r) for _ in range(n)]
    xy = _device_multiples(ctx, C, ks)
    dp = ctx.upload_points(xy)
    G = C.from_affine(C.g)
    for mix in ("uniform", "bits"):
        if mix == "uniform":
            s = [rnd.randrange(C.r) for _ in range(n)]
        else:
            s = [rnd.randrange(2) for _ in range(n)]
        tot = sum(a * b for a, b in zip(s, ks)) % C.r
        assert ctx.msm(dp, enc_scalars(C, s)) == C.g1_xy_bytes(C.mul(G, tot)), (which, lg, mix)
    dp.free()


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_fixed_base_commitments(which, ctx_bls, ctx_bn):
    """commit_to_field_element(g, h, v, r) = v*g + r*h (prover.rs:123,496-500) and g*w (prover.rs:550) through the cached
    window tables, incl. zero scalars (identity result), r-1 and a 3-base table."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    g, h, k3 = C.g1_from_msg_hash(b"g"), C.g1_from_msg_hash(b"h"), C.g1_from_msg_hash(b"third")
    pairs = [(0, 0), (1, 0), (0, 1), (C.r - 1, C.r - 1), (5, C.r - 5), (15, 16), (1 << 252, (1 << 253) + 12345)]
    pairs += [(C.synth_scalar(8, 2 * i), C.synth_scalar(8, 2 * i + 1)) for i in range(9)]
    sb = b"".join(C.fr_to_bytes(v) + C.fr_to_bytes(r) for v, r in pairs)
    for _ in range(2):                                              # second call hits the ctx cache
        got = ctx.commit_batch(enc_points(C, [g, h]), sb, len(pairs))
        assert got == enc_points(C, [C.binary_scalar_mul(g, h, v, r) for v, r in pairs])
    trip = [(C.synth_scalar(9, 3 * i), C.synth_scalar(9, 3 * i + 1), C.synth_scalar(9, 3 * i + 2)) for i in range(4)]
    got = ctx.commit_batch(enc_points(C, [g, h, k3]), b"".join(b"".join(C.fr_to_bytes(x) for x in t) for t in trip), len(trip))
    assert got == enc_points(C, [C.msm([g, h, k3], list(t)) for t in trip])
    assert ctx.commit_batch(enc_points(C, [g, h]), b"", 0) == b""


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_msm_over_precomputed_tables(which, ctx_bls, ctx_bn):
    """bpgpu_points_precompute: the table path (no doublings, no buckets) gives the same bytes as the general path and as
    the oracle, for host scalars, device scalars, offsets, edge scalars, and mixed table / general parts."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    n = 100
    pts = rand_points(C, n, 77)
    s = C.synth_scalars(31, n)
    s[0], s[1], s[2], s[3] = 0, 1, C.r - 1, 15
    plain = ctx.upload_points(enc_points(C, pts))
    tab = ctx.upload_points(enc_points(C, pts)).precompute()
    assert tab.has_tables and not plain.has_tables
    exp = C.g1_xy_bytes(C.msm(pts, s))
    assert ctx.msm(plain, enc_scalars(C, s)) == exp
    assert ctx.msm(tab, enc_scalars(C, s)) == exp
    ds = ctx.upload_scalars(enc_scalars(C, s))
    assert ctx.msm_device(tab, ds) == exp
    # offsets into the table
    assert ctx.msm(tab, enc_scalars(C, s[10:30]), off=40, n=20) == C.g1_xy_bytes(C.msm(pts[40:60], s[10:30]))
    assert ctx.msm_device(tab, ds, poff=5, soff=7, n=33) == C.g1_xy_bytes(C.msm(pts[5:38], s[7:40]))
    # all-zero scalars -> identity; single term
    assert ctx.msm(tab, enc_scalars(C, [0] * n)) == C.g1_xy_bytes(C.INF)
    assert ctx.msm(tab, enc_scalars(C, [C.r - 2]), off=99, n=1) == C.g1_xy_bytes(C.mul(pts[99], C.r - 2))
    # mixed: table part + general host part + general device part, with cancelling terms
    extra = rand_points(C, 5, 78)
    es = C.synth_scalars(32, 5)
    got = ctx.msm_parts([(tab, ds, n), (enc_points(C, extra), enc_scalars(C, es), 5), (plain, enc_scalars(C, [C.r - x for x in s]), n)])
    assert got == C.g1_xy_bytes(C.msm(extra, es))
    plain.free()
    tab.free()


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_batch_identity_check(which, ctx_bls, ctx_bn):
    """bpgpu_msm_batch_is_identity: per-instance verdict of sum f_i*F_i + sum v_j*V_j == identity, with fixed terms from a
    precomputed table (at an offset) and from a host base, variable points incl. the identity and repeated points."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    nf, vn, batch = 9, 5, 7
    fixed = rand_points(C, nf + 3, 5)
    hb = C.g1_from_msg_hash(b"h")
    tab = ctx.upload_points(enc_points(C, fixed)).precompute()
    fs, vp, vs, expect = [], [], [], []
    for b in range(batch):
        f = C.synth_scalars(40 + b, nf + 1)
        if b == 2:
            f = [0] * (nf + 1)
        V = rand_points(C, vn - 1, 100 + b)
        if b == 3:
            V[1] = C.INF
            V[2] = V[0]
        v = C.synth_scalars(60 + b, vn - 1)
        R = C.msm(fixed[2:2 + nf] + [hb] + V, f + v)
        last_s = C.r - 1
        if b % 2 == 1:
            last_s = C.r - 2                                # wrong: sum is R - 2R = -R, not the identity
        fs += f
        vp += V + [R]
        vs += v + [last_s]
        expect.append(1 if (b % 2 == 0 or C.is_inf(R)) else 0)
    got = ctx.msm_batch_is_identity([(tab, 2, nf), C.g1_xy_bytes(hb)], batch, enc_scalars(C, fs), enc_points(C, vp), enc_scalars(C, vs), vn)
    assert list(got) == expect
    # no variable part at all: only the all-zero instance is the identity
    got = ctx.msm_batch_is_identity([(tab, 2, nf), C.g1_xy_bytes(hb)], batch, enc_scalars(C, fs), b"", b"", 0)
    assert list(got) == [1 if b == 2 else 0 for b in range(batch)]
    plain = ctx.upload_points(enc_points(C, fixed))
    with pytest.raises(Exception):                          # tables are required
        ctx.msm_batch_is_identity([(plain, 0, nf)], 1, enc_scalars(C, fs[:nf]), b"", b"", 0)


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_msm_sharded_over_contexts(which, bp, ctx_bls, ctx_bn):
    """bph_msm_sharded: the points split unevenly over three contexts (here three streams of one GPU; one
    per GPU in production), an empty shard included -- the bytes equal the unsharded MSM's and the oracle's."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    cid = bp.BLS12_381 if which == "bls" else bp.BN254
    n = 700
    P = rand_points(C, n, 31)
    s = C.synth_scalars(8, n)
    exp = C.g1_xy_bytes(C.msm(P, s))
    ctxs = [ctx, bp.Context(cid, 0), bp.Context(cid, 0), bp.Context(cid, 0)]
    cuts = [0, 300, 301, 301, n]
    shards = [ctxs[k].upload_points(enc_points(C, P[cuts[k]:cuts[k + 1]])) for k in range(4)]
    assert bp.msm_sharded(ctxs, shards, enc_scalars(C, s)) == exp
    assert ctx.msm_refs(enc_points(C, P), enc_scalars(C, s)) == exp
    for sh in shards:
        sh.free()
    for c in ctxs[1:]:
        c.close()


def _np_scalars(n, seed, mb, bits=253):
    """n uniform `bits`-bit scalars as (big-endian MODBYTES byte block, list of ints) -- numpy, for the 2^21 / 2^22 cases"""
    import numpy as np
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    keep = bits - 248
    raw[:, 0] &= (1 << keep) - 1 if keep > 0 else 0
    out = np.zeros((n, mb), dtype=np.uint8)
    out[:, mb - 32:] = raw
    b = out.tobytes()
    return b, [int.from_bytes(b[i * mb + mb - 32:(i + 1) * mb], "big") for i in range(n)]


@pytest.mark.parametrize("which,lg", [("bls", 21), ("bls", 22), ("bn", 22)])
def test_msm_config4_top_sizes(which, lg, ctx_bls, ctx_bn):
    """BASELINE.json config 4 ("standalone G1 MSM sweep 2^10-2^22") at its two largest sizes, which nothing exercised in
    round 1: sum s_i*(k_i*G) == (sum s_i*k_i mod r)*G for uniform scalars and for a 0/1-heavy mix (giant buckets), through
    the resident-bases entry point and through the host-buffer one (copy queue + event path)."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    n, mb = 1 << lg, C.MODBYTES
    kb, ks = _np_scalars(n, 100 + lg, mb)
    g = C.g1_xy_bytes(C.from_affine(C.g))
    xy = ctx.selftest_group(2, g * n, g * n, kb)
    dp = ctx.upload_points(xy)
    G = C.from_affine(C.g)
    sb, s = _np_scalars(n, 200 + lg, mb)
    tot = sum(a * b for a, b in zip(s, ks)) % C.r
    exp = C.g1_xy_bytes(C.mul(G, tot))
    assert ctx.msm(dp, sb) == exp
    assert ctx.msm_le32(dp, b"".join(x.to_bytes(32, "little") for x in s)) == exp
    if lg == 22:
        assert ctx.msm_refs(xy, sb) == exp
    import numpy as np
    bits = np.random.default_rng(5).integers(0, 2, size=n, dtype=np.uint8)
    blk = np.zeros((n, mb), dtype=np.uint8)
    blk[:, mb - 1] = bits
    tot = sum(k for k, b in zip(ks, bits.tolist()) if b) % C.r
    assert ctx.msm(dp, blk.tobytes()) == C.g1_xy_bytes(C.mul(G, tot))
    dp.free()


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_fixed_schedule_table_sums_give_the_same_bytes(which, bp, ctx_bls, ctx_bn):
    """bpgpu_ctx_set_fixed_schedule (secret scalars; the reference's inner_product_const_time, prover.rs:347-362): no digit or
    scalar is skipped, results are the same group elements -- zero scalars, 0/1 witnesses, short and full scalars, and a
    whole range proof with BPH's switch on."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    n = 70
    pts = rand_points(C, n, 5)
    tab = ctx.upload_points(enc_points(C, pts)).precompute()
    rnd = random.Random(4)
    cases = [[0] * n, [rnd.randrange(2) for _ in range(n)], [rnd.randrange(256) for _ in range(n)], C.synth_scalars(6, n),
             [0, 1, C.r - 1, 1 << 200] + [0] * (n - 4)]
    ctx.set_fixed_schedule(True)
    try:
        for s in cases:
            assert ctx.msm(tab, enc_scalars(C, s)) == C.g1_xy_bytes(C.msm(pts, s))
            ds = ctx.upload_scalars(enc_scalars(C, s))
            assert ctx.msm_device(tab, ds) == C.g1_xy_bytes(C.msm(pts, s))
            ds.free()
    finally:
        ctx.set_fixed_schedule(False)
    tab.free()
    dG, dH = ctx.get_generators("G", 16, precompute=True), ctx.get_generators("H", 16, precompute=True)
    gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
    ref, comms = ctx.range_prove(b"FS", gx, hx, dG, dH, [200, 7], 8, seed=3)
    bp.lib().bph_set_secret_fixed_schedule(1)
    try:
        got, comms2 = ctx.range_prove(b"FS", gx, hx, dG, dH, [200, 7], 8, seed=3)
    finally:
        bp.lib().bph_set_secret_fixed_schedule(0)
    assert got == ref and comms2 == comms


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_msm_begin_finish_pipelined_over_two_contexts(which, bp, ctx_bls, ctx_bn):
    """bpgpu_msm_*_begin / bpgpu_msm_finish: two MSMs in flight on two contexts of one device, all four input forms, with
    and without window tables; one MSM in flight per context (a second begin is an argument error)"""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    other = bp.Context(ctx.curve, 0)
    n = 3000
    P = rand_points(C, n, 12)
    xy = enc_points(C, P)
    dp = ctx.upload_points(xy)
    tab = ctx.upload_points(xy[:200 * 2 * C.MODBYTES]).precompute()
    sets = [C.synth_scalars(50 + k, n) for k in range(4)]
    exp = [C.g1_xy_bytes(C.msm(P, s)) for s in sets]
    ds = [ctx.upload_scalars(enc_scalars(C, s)) for s in sets[:2]]
    ctx.msm_device_begin(dp, ds[0])
    other.msm_device_begin(dp, ds[1])                      # the points live on ctx; any context of the device may read them
    with pytest.raises(Exception):
        ctx.msm_device_begin(dp, ds[1])
    assert ctx.msm_finish() == exp[0]
    assert other.msm_finish() == exp[1]
    with pytest.raises(Exception):
        ctx.msm_finish()                                   # nothing pending
    sb2, sb3 = enc_scalars(C, sets[2]), enc_scalars(C, sets[3])
    ctx.msm_begin(dp, sb2)
    other.msm_refs_begin(xy, sb3)
    assert other.msm_finish() == exp[3]
    assert ctx.msm_finish() == exp[2]
    le = b"".join(x.to_bytes(32, "little") for x in sets[0])
    ctx.msm_le32_begin(dp, le)
    other.msm_begin(tab, sb2[:200 * C.MODBYTES])           # table path
    assert ctx.msm_finish() == exp[0]
    assert other.msm_finish() == C.g1_xy_bytes(C.msm(P[:200], sets[2][:200]))
    # an invalid point surfaces at finish
    bad = bytearray(xy)
    bad[5] ^= 1
    other.msm_refs_begin(bytes(bad), sb3)
    with pytest.raises(bp.BpgpuError) as e:
        other.msm_finish()
    assert e.value.code == -5
    other.close()
    dp.free()
    tab.free()


@pytest.mark.parametrize("which", ["bls", "bn"])
@pytest.mark.parametrize("n", [300, 2049, 4096])
def test_small_msm_kernel_skewed_digits(which, n, ctx_bls, ctx_bn):
    """The one-launch small-MSM kernel (csrc/msm.cu k_msm_small, n <= 4096): equal scalars (ONE bucket per window, summed by
    a warp), 0/1 witnesses, small scalars (most windows empty), r - 1 (all digits negative), a few distinct values (wide and
    narrow buckets mixed), repeated points and identities among the inputs -- against the oracle's sum."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    base = rand_points(C, 37, 900 + n)
    P = [base[i % 37] if i % 53 else C.INF for i in range(n)]
    rnd = random.Random(n)
    few = [rnd.randrange(C.r) for _ in range(5)]
    for s in ([few[0]] * n, [rnd.randrange(2) for _ in range(n)], [rnd.randrange(1 << 20) for _ in range(n)], [C.r - 1] * n,
              [few[rnd.randrange(5)] for _ in range(n)], [rnd.randrange(C.r) if i % 3 else 0 for i in range(n)]):
        _check(ctx, C, P, s)
