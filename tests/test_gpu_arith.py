"""GPU parity: device Montgomery field arithmetic and group law vs the oracle (bit exact)."""
import random

import pytest

from tests.util import curve_of, dec_points, dec_scalars, enc_points, enc_scalars, rand_points

pytestmark = pytest.mark.gpu


def _field_cases(p, rnd, n):
    edge = [0, 1, 2, p - 1, p - 2, (1 << (p.bit_length() - 1)), (1 << 32) - 1, (1 << 64), p >> 1]
    vals = edge + [rnd.randrange(p) for _ in range(n)]
    a = vals + [rnd.choice(vals) for _ in range(n)]
    b = [rnd.choice(vals) for _ in range(len(a))]
    return a, b


@pytest.mark.parametrize("which", ["bls", "bn"])
@pytest.mark.parametrize("field", [0, 1])
def test_field_ops(which, field, ctx_bls, ctx_bn):
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    p = C.p if field == 0 else C.r
    m = C.MODBYTES
    rnd = random.Random(11 + field)
    a, b = _field_cases(p, rnd, 400)
    ea = b"".join(x.to_bytes(m, "big") for x in a)
    eb = b"".join(x.to_bytes(m, "big") for x in b)
    exp = {0: lambda x, y: x * y % p, 1: lambda x, y: (x + y) % p, 2: lambda x, y: (x - y) % p,
           3: lambda x, y: pow(x, p - 2, p), 4: lambda x, y: x * x % p}
    for op, f in exp.items():
        got = ctx.selftest_field(field, op, ea, eb)
        got = [int.from_bytes(got[i:i + m], "big") for i in range(0, len(got), m)]
        assert got == [f(x, y) for x, y in zip(a, b)], (which, field, op)


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_group_ops(which, ctx_bls, ctx_bn):
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    rnd = random.Random(5)
    P = rand_points(C, 24, 1) + [C.INF, C.INF]
    Q = rand_points(C, 24, 2) + [C.INF, P[0]]
    # add cases that hit the complete-formula branches: P+P, P+(-P), P+O, O+Q
    P += [P[0], P[1], P[2], C.INF]
    Q += [P[0], C.neg(P[1]), C.INF, Q[3]]
    ks = [rnd.randrange(C.r) for _ in P]
    ks[0], ks[1], ks[2] = 0, 1, C.r - 1
    ep, eq, ek = enc_points(C, P), enc_points(C, Q), enc_scalars(C, ks)
    got = dec_points(C, ctx.selftest_group(0, ep, eq, ek))
    assert [C.to_affine(g) for g in got] == [C.to_affine(C.add(a, b)) for a, b in zip(P, Q)]
    got = dec_points(C, ctx.selftest_group(1, ep, eq, ek))
    assert [C.to_affine(g) for g in got] == [C.to_affine(C.dbl(a)) for a in P]
    got = dec_points(C, ctx.selftest_group(2, ep, eq, ek))
    assert [C.to_affine(g) for g in got] == [C.to_affine(C.mul(a, k)) for a, k in zip(P, ks)]
    got = dec_points(C, ctx.selftest_group(3, ep, eq, ek))
    assert [C.to_affine(g) for g in got] == [C.to_affine(C.add(C.dbl(b), a)) for a, b in zip(P, Q)]


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_residency_roundtrip(which, ctx_bls, ctx_bn):
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    P = rand_points(C, 9, 3) + [C.INF]
    dp = ctx.upload_points(enc_points(C, P))
    assert len(dp) == 10
    assert dp.download() == enc_points(C, P)
    assert dp.download(3, 2) == enc_points(C, P[3:5])
    dp.free()
    xs = [0, 1, C.r - 1, 12345678901234567890] + [C.synth_scalar(1, i) for i in range(7)]
    ds = ctx.upload_scalars(enc_scalars(C, xs))
    assert dec_scalars(C, ds.download()) == xs
    ds.free()
