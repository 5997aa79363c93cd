"""CPU tests: the device headers (fp.cuh / ec.cuh) compiled for the host run the SAME limb schedule and
group formulas as the kernels; checked against the oracle's integers."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle.curves import BLS12_381 as B, BN254 as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hc") / "host_check.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "host_check.cpp")])
    return ctypes.CDLL(so)


def _L(x, n):
    return (ctypes.c_uint32 * n)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def _I(arr):
    return sum(int(v) << (32 * i) for i, v in enumerate(arr))


@pytest.mark.parametrize("fid,p,n", [(0, B.p, 12), (1, B.r, 8), (2, N.p, 8), (3, N.r, 8)])
def test_montgomery_schedule(hc, fid, p, n):
    R = 1 << (32 * n)
    rnd = random.Random(fid)

    def op(o, a, b=0):
        out = (ctypes.c_uint32 * n)()
        hc.hc_fp_op(fid, o, _L(a, n), _L(b, n), out)
        return _I(out)
    edge = [0, 1, p - 1, R % p, p - 2, (1 << (32 * n - 1)) % p]
    # carry-heavy operands for the dedicated squaring (fp.cuh sqr_t): runs of 0xffffffff limbs, single high limbs, p/2
    carry = [(R - 1) % p, (1 << (32 * (n - 1))) - 1, ((1 << (32 * (n - 1))) - 1) ^ 0xffffffff, p >> 1, (p >> 1) + 1,
             ((1 << 32) - 1) << (32 * (n - 2)), sum(0xffffffff << (64 * k) for k in range(n // 2)) % p,
             sum(0xffffffff << (64 * k + 32) for k in range(n // 2)) % p]
    vals = edge + carry + [rnd.randrange(p) for _ in range(400)]
    Rinv = pow(R, -1, p)
    for a in vals:
        for b in rnd.sample(vals, 4) + edge:
            assert op(0, a, b) == a * b * Rinv % p
            assert op(1, a, b) == (a + b) % p
            assert op(2, a, b) == (a - b) % p
        assert op(3, a) == a * R % p
        assert op(4, a) == a * Rinv % p
        assert op(6, a) == a * a * Rinv % p
    for a in vals[:12]:
        assert op(5, a * R % p) == pow(a, p - 2, p) * R % p


@pytest.mark.parametrize("fid,p,n", [(0, B.p, 12), (1, B.r, 8), (2, N.p, 8), (3, N.r, 8)])
def test_host_field_ops(hc, fid, p, n):
    """host_fp.h on 64-bit limbs (the host finish of every MSM, challenge arithmetic): Montgomery product, sum, difference and
    the conversions against the integers, with the extreme operands 0, 1, p - 1 and R mod p."""
    R = 1 << (32 * n)
    Rinv = pow(R, -1, p)
    rnd = random.Random(70 + fid)

    def op(o, a, b=0):
        out = (ctypes.c_uint32 * n)()
        hc.hc_hostfp_op(fid, o, _L(a, n), _L(b, n), out)
        return _I(out)
    edge = [0, 1, p - 1, p - 2, R % p, (p - 1) // 2, ((1 << (32 * n - 1)) - 1) % p]
    vals = edge + [rnd.randrange(p) for _ in range(150)]
    for a in vals:
        for b in rnd.sample(vals, 5) + edge:
            assert op(0, a, b) == a * b * Rinv % p
            assert op(1, a, b) == (a + b) % p
            assert op(2, a, b) == (a - b) % p
        assert op(3, a) == a * R % p
        assert op(4, a) == a * Rinv % p
        assert op(5, a) == a * a * Rinv % p


@pytest.mark.parametrize("fid,p,n", [(0, B.p, 12), (1, B.r, 8), (2, N.p, 8), (3, N.r, 8)])
def test_host_field_inverse(hc, fid, p, n):
    """host_fp.h (the host layer's 64-bit-limb field code): inv() by binary extended Euclid equals the Fermat ladder and the
    integers, incl. 0 -> 0, 1, p - 1 and values with long runs of zero bits."""
    R = 1 << (32 * n)
    rnd = random.Random(50 + fid)
    vals = [0, 1, 2, p - 1, p - 2, (p + 1) // 2, 1 << 200, (1 << 250) % p, R % p] + [rnd.randrange(p) for _ in range(200)]
    for a in vals:
        out = (ctypes.c_uint32 * n)()
        same = hc.hc_hostfp_inv(fid, _L(a * R % p, n), out)
        assert same == 1
        assert _I(out) == (pow(a, p - 2, p) * R % p if a else 0)


@pytest.mark.parametrize("fid,p,n", [(0, B.p, 12), (2, N.p, 8)])
def test_fused_two_product_reduction(hc, fid, p, n):
    """Fp::mul2 = (a*b + c*d)/R mod p with one reduction (the Y3 term of the group law), incl. the extreme operands
    p-1 and the non-canonical multiplicand p that a negated zero could produce."""
    R = 1 << (32 * n)
    Rinv = pow(R, -1, p)
    rnd = random.Random(100 + fid)
    edge = [0, 1, p - 1, p - 2, R % p]
    vals = edge + [rnd.randrange(p) for _ in range(60)]

    def mul2(a, b, c, d):
        out = (ctypes.c_uint32 * n)()
        hc.hc_fp_mul2(fid, _L(a, n), _L(b, n), _L(c, n), _L(d, n), out)
        return _I(out)
    for a in vals:
        for _ in range(6):
            b, c, d = rnd.choice(vals), rnd.choice(vals), rnd.choice(vals)
            assert mul2(a, b, c, d) == (a * b + c * d) * Rinv % p
    for x in edge:
        assert mul2(p - 1, p - 1, p - 1, p - 1) == 2 * (p - 1) * (p - 1) * Rinv % p
        assert mul2(x, p - 1, p, p - 1) == (x * (p - 1)) * Rinv % p          # c = p acts as 0


@pytest.mark.parametrize("cid,C,n", [(0, B, 12), (1, N, 8)])
def test_xyzz_group_law(hc, cid, C, n):
    p, R = C.p, 1 << (32 * n)
    rnd = random.Random(cid)
    M = lambda x: x * R % p

    def pack(vals):
        arr = (ctypes.c_uint32 * (n * len(vals)))()
        for k, v in enumerate(vals):
            for i in range(n):
                arr[k * n + i] = (v >> (32 * i)) & 0xFFFFFFFF
        return arr

    def xyzz_of(P):
        if P[2] == 0:
            return [0, 0, 0, 0]
        a = C.to_affine(P)
        z = rnd.randrange(1, p)
        zz, zzz = z * z % p, z * z * z % p
        return [M(a[0] * zz % p), M(a[1] * zzz % p), M(zz), M(zzz)]

    def aff_of(P):
        a = C.to_affine(P)
        return [0, 0] if a is None else [M(a[0]), M(a[1])]

    def to_aff(arr):
        vals = [sum(int(arr[k * n + i]) << (32 * i) for i in range(n)) * pow(R, -1, p) % p for k in range(4)]
        x, y, zz, zzz = vals
        if zz == 0:
            return None
        assert pow(zz, 3, p) == zzz * zzz % p
        return (x * pow(zz, -1, p) % p, y * pow(zzz, -1, p) % p)

    def ec(o, P, Q=None, k=0, qa=False):
        out = (ctypes.c_uint32 * (4 * n))()
        q = pack(aff_of(Q)) if qa else pack(xyzz_of(Q) if Q is not None else [0, 0, 0, 0])
        hc.hc_ec_op(cid, o, pack(xyzz_of(P)), q, k, out)
        return to_aff(out)
    G = C.from_affine(C.g)
    pts = [C.mul(G, rnd.randrange(1, C.r)) for _ in range(4)] + [C.INF]
    for P in pts:
        for Q in pts + [P, C.neg(P)]:
            assert ec(0, P, Q, qa=True) == C.to_affine(C.add(P, Q))
            assert ec(1, P, Q) == C.to_affine(C.add(P, Q))
        assert ec(2, P) == C.to_affine(C.dbl(P))
        for k in (0, 1, 2, 77, 32767):
            assert ec(3, P, k=k) == C.to_affine(C.mul(P, k))
        assert ec(4, P) == C.to_affine(P)
