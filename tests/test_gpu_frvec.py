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
