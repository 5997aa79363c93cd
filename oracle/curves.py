"""Oracle: Fr / Fq / G1 arithmetic and AMCL byte encodings (test infrastructure).

Restates, with Python integers, the behaviour of the `amcl_wrapper` types the
reference uses on its hot path (call sites
`/root/reference/src/ipp.rs:77-129`, `src/transcript.rs:47-60`,
`src/utils/mod.rs:16-23`).  The dependency source is not in the container, so
each function states the published AMCL algorithm it restates.

Curves (Cargo.toml:22-27 features `bls381` (default) and `bn254`):
  * BLS12_381 : y^2 = x^3 + 4 over the 381-bit prime, 255-bit order r.
  * BN254     : AMCL's "BN254" = the Nogami curve y^2 = x^3 + 2 (NOT alt_bn128).

Points are Jacobian triples (X, Y, Z) of ints with Z == 0 for the identity;
affine points are (x, y) or None for the identity.
"""
import hashlib


class Curve:
    def __init__(self, name, cid, p, r, b, gx, gy, cof, modbytes):
        self.name, self.id = name, cid
        self.p, self.r, self.b = p, r, b
        self.g = (gx, gy)
        self.cof = cof
        self.MODBYTES = modbytes
        assert p % 4 == 3
        assert (gy * gy - gx * gx * gx - b) % p == 0

    # ------------------------------------------------------------------ Fr
    def fr_from_bytes(self, buf):
        """FieldElement::from(&[u8; MODBYTES]) = BIG::frombytes (big endian) then rmod(r)
        (used by transcript.rs:55-60)."""
        assert len(buf) == self.MODBYTES
        return int.from_bytes(buf, "big") % self.r

    def fr_to_bytes(self, x):
        """FieldElement::to_bytes = BIG::tobytes: MODBYTES big endian (transcript.rs:47-49)."""
        return int(x % self.r).to_bytes(self.MODBYTES, "big")

    def fr_inv(self, x):
        """FieldElement::inverse; AMCL's invmodp maps 0 -> 0."""
        x %= self.r
        return pow(x, self.r - 2, self.r) if x else 0

    def fr_from_msg_hash(self, msg):
        """FieldElement::from_msg_hash = from(SHAKE256(msg)[..MODBYTES])."""
        return self.fr_from_bytes(hashlib.shake_256(msg).digest(self.MODBYTES))

    def fr_batch_invert(self, xs):
        """FieldElement::batch_invert -> (inverses, product of all inverses) (ipp.rs:295)."""
        invs = [self.fr_inv(x) for x in xs]
        prod = 1
        for i in invs:
            prod = prod * i % self.r
        return invs, prod

    def vandermonde(self, x, n):
        """FieldElementVector::new_vandermonde_vector(x, n) = [1, x, ..., x^(n-1)]."""
        out, cur = [], 1
        for _ in range(n):
            out.append(cur)
            cur = cur * x % self.r
        return out

    def inner_product(self, a, b):
        assert len(a) == len(b)
        return sum(x * y for x, y in zip(a, b)) % self.r

    # ------------------------------------------------------------------ G1 (Jacobian, a = 0)
    INF = (0, 1, 0)

    def is_inf(self, P):
        return P[2] == 0

    def from_affine(self, a):
        return self.INF if a is None else (a[0], a[1], 1)

    def to_affine(self, P):
        if P[2] == 0:
            return None
        p = self.p
        zi = pow(P[2], p - 2, p)
        zi2 = zi * zi % p
        return (P[0] * zi2 % p, P[1] * zi2 * zi % p)

    def on_curve(self, a):
        if a is None:
            return True
        x, y = a
        return (y * y - x * x * x - self.b) % self.p == 0

    def dbl(self, P):
        X, Y, Z = P
        if Z == 0 or Y == 0:
            return self.INF
        p = self.p
        A = X * X % p
        B = Y * Y % p
        C = B * B % p
        D = 2 * ((X + B) * (X + B) - A - C) % p
        E = 3 * A % p
        F = E * E % p
        X3 = (F - 2 * D) % p
        Y3 = (E * (D - X3) - 8 * C) % p
        Z3 = 2 * Y * Z % p
        return (X3, Y3, Z3)

    def add(self, P, Q):
        if P[2] == 0:
            return Q
        if Q[2] == 0:
            return P
        p = self.p
        X1, Y1, Z1 = P
        X2, Y2, Z2 = Q
        Z1Z1 = Z1 * Z1 % p
        Z2Z2 = Z2 * Z2 % p
        U1 = X1 * Z2Z2 % p
        U2 = X2 * Z1Z1 % p
        S1 = Y1 * Z2 * Z2Z2 % p
        S2 = Y2 * Z1 * Z1Z1 % p
        if U1 == U2:
            return self.dbl(P) if S1 == S2 else self.INF
        H = (U2 - U1) % p
        I = 4 * H * H % p
        J = H * I % p
        rr = 2 * (S2 - S1) % p
        V = U1 * I % p
        X3 = (rr * rr - J - 2 * V) % p
        Y3 = (rr * (V - X3) - 2 * S1 * J) % p
        Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % p
        return (X3, Y3, Z3)

    def neg(self, P):
        return (P[0], (-P[1]) % self.p, P[2])

    def mul(self, P, k):
        """k*P for an integer k >= 0 (used for `&G1 * &FieldElement`, cofactor clearing)."""
        R = self.INF
        if k == 0 or P[2] == 0:
            return R
        for bit in bin(k)[2:]:
            R = self.dbl(R)
            if bit == "1":
                R = self.add(R, P)
        return R

    def eq(self, P, Q):
        return self.to_affine(P) == self.to_affine(Q)

    # ------------------------------------------------------------------ encodings
    def g1_to_bytes(self, P):
        """G1::to_bytes = ECP::tobytes(compress=false) after affine():
        0x04 || X || Y, MODBYTES big endian each; AMCL's identity is (0,1,0) whose
        affine() is a no-op, hence 04 || 0..0 || 0..01 (absorbed by every 1-phase
        R1CS transcript: prover.rs:429-434)."""
        a = self.to_affine(P)
        x, y = (0, 1) if a is None else a
        return b"\x04" + x.to_bytes(self.MODBYTES, "big") + y.to_bytes(self.MODBYTES, "big")

    def g1_from_bytes(self, buf):
        m = self.MODBYTES
        assert len(buf) == 2 * m + 1 and buf[0] == 4
        x = int.from_bytes(buf[1:1 + m], "big")
        y = int.from_bytes(buf[1 + m:], "big")
        if (x, y) == (0, 1):
            return self.INF
        assert self.on_curve((x, y))
        return (x, y, 1)

    def g1_xy_bytes(self, P):
        """The C-ABI point encoding of include/bpgpu.h: to_bytes() without the 0x04 tag."""
        return self.g1_to_bytes(P)[1:]

    def g1_from_xy_bytes(self, buf):
        return self.g1_from_bytes(b"\x04" + bytes(buf))

    # ------------------------------------------------------------------ hash to curve
    def sqrt_fq(self, a):
        """p = 3 (mod 4): candidate a^((p+1)/4); None when a is a non-residue."""
        s = pow(a, (self.p + 1) // 4, self.p)
        return s if s * s % self.p == a % self.p else None

    def g1_from_msg_hash(self, msg):
        """G1::from_msg_hash = ECP::mapit(SHAKE256(msg)[..MODBYTES]): x = be(h) mod p;
        try x, x+1, ... until x^3+b is a square, take the root whose canonical value is
        EVEN (new_bigint(x, 0)), then clear the cofactor with a plain scalar mult
        (utils/mod.rs:16-23 builds every generator this way)."""
        h = hashlib.shake_256(msg).digest(self.MODBYTES)
        x = int.from_bytes(h, "big") % self.p
        while True:
            y = self.sqrt_fq((x * x * x + self.b) % self.p)
            if y is not None:
                if y & 1:
                    y = self.p - y
                P = self.mul((x, y, 1), self.cof)
                x = (x + 1) % self.p
                if P[2] != 0:
                    return P
            else:
                x = (x + 1) % self.p

    def get_generators(self, prefix, n):
        """utils/mod.rs:16-23: G_i = from_msg_hash(prefix || decimal(i)), i = 1..n."""
        return [self.g1_from_msg_hash((prefix + str(i)).encode()) for i in range(1, n + 1)]

    # ------------------------------------------------------------------ MSM
    def msm(self, points, scalars):
        """sum_i scalars[i] * points[i]: the value of
        G1Vector::multi_scalar_mul_var_time / inner_product_var_time_with_ref_vecs /
        inner_product_const_time (ipp.rs:91,104,158,170,251; verifier.rs:451;
        prover.rs:347-362).  The result is a group element, so the bucket method here
        (8-bit unsigned windows) is interchangeable with AMCL's Straus."""
        if len(points) != len(scalars):
            raise ValueError("UnequalSizeVectors")
        c = 8 if len(points) > 64 else 4
        r = self.r
        scalars = [s % r for s in scalars]
        nwin = (r.bit_length() + c - 1) // c
        acc = self.INF
        for w in range(nwin - 1, -1, -1):
            for _ in range(c):
                acc = self.dbl(acc)
            buckets = {}
            for P, s in zip(points, scalars):
                d = (s >> (w * c)) & ((1 << c) - 1)
                if d and P[2] != 0:
                    buckets[d] = self.add(buckets[d], P) if d in buckets else P
            run, tot = self.INF, self.INF
            for d in range((1 << c) - 1, 0, -1):
                if d in buckets:
                    run = self.add(run, buckets[d])
                tot = self.add(tot, run)
            acc = self.add(acc, tot)
        return acc

    def binary_scalar_mul(self, g, h, r1, r2):
        """G1::binary_scalar_mul(&self=g, h, r1, r2) = r1*g + r2*h (ipp.rs:119-129,185-187;
        commit_to_field_element(g,h,v,r) = g*v + h*r, prover.rs:123)."""
        return self.add(self.mul(g, r1 % self.r), self.mul(h, r2 % self.r))

    # ------------------------------------------------------------------ synthetic inputs
    def synth_scalar(self, seed, i, tag=b"s"):
        m = seed.to_bytes(8, "little") + tag + i.to_bytes(8, "little")
        return self.fr_from_bytes(hashlib.shake_256(m).digest(self.MODBYTES))

    def synth_scalars(self, seed, n, tag=b"s"):
        return [self.synth_scalar(seed, i, tag) for i in range(n)]


_z = -0xD201000000010000
BLS12_381 = Curve(
    "BLS12_381", 0,
    p=0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    r=0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    b=4,
    gx=0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    gy=0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
    cof=(_z - 1) ** 2 // 3,
    modbytes=48,
)
assert BLS12_381.cof == 0x396C8C005555E1568C00AAAB0000AAAB
assert BLS12_381.r == _z ** 4 - _z ** 2 + 1

_u = -(2 ** 62 + 2 ** 55 + 1)
_bn_p = 36 * _u ** 4 + 36 * _u ** 3 + 24 * _u ** 2 + 6 * _u + 1
_bn_r = 36 * _u ** 4 + 36 * _u ** 3 + 18 * _u ** 2 + 6 * _u + 1
assert _bn_p == 0x2523648240000001BA344D80000000086121000000000013A700000000000013
assert _bn_r == 0x2523648240000001BA344D8000000007FF9F800000000010A10000000000000D
BN254 = Curve("BN254", 1, p=_bn_p, r=_bn_r, b=2, gx=_bn_p - 1, gy=1, cof=1, modbytes=32)

CURVES = {"BLS12_381": BLS12_381, "BN254": BN254, 0: BLS12_381, 1: BN254}
