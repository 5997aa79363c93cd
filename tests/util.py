"""Shared helpers for the parity tests: byte marshalling between the oracle's integers and the ABI."""
import random

from oracle.curves import BLS12_381, BN254


def curve_of(ctx):
    return BLS12_381 if ctx.curve == 0 else BN254


def enc_scalars(C, xs):
    return b"".join(C.fr_to_bytes(x) for x in xs)


def dec_scalars(C, buf):
    m = C.MODBYTES
    return [int.from_bytes(buf[i:i + m], "big") for i in range(0, len(buf), m)]


def enc_points(C, ps):
    return b"".join(C.g1_xy_bytes(p) for p in ps)


def dec_points(C, buf):
    m = 2 * C.MODBYTES
    return [C.g1_from_xy_bytes(buf[i:i + m]) for i in range(0, len(buf), m)]


def rand_points(C, n, seed):
    """n pseudo-random G1 points as cheap as the oracle can make them: a random walk k_i*G."""
    rnd = random.Random(seed)
    G = C.from_affine(C.g)
    base = C.mul(G, rnd.randrange(1, C.r))
    step = C.mul(G, rnd.randrange(1, C.r))
    out, cur = [], base
    for _ in range(n):
        out.append(cur)
        cur = C.add(cur, step)
        if rnd.random() < 0.1:
            step = C.dbl(step)
    return out
