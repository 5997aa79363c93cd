#!/usr/bin/env python3
"""Generates tests/golden/*.json from the oracle (oracle/*.py, the CPU restatement of the reference).

The reference holds no golden vectors and cannot be built here, so these are outputs of the oracle,
committed so that (a) the oracle itself cannot drift silently (tests/test_golden.py re-derives the cheap ones on the CPU)
and (b) the GPU path can be checked at sizes the Python oracle needs minutes for (config 2: n = 1024 multipliers).

    python tests/golden/make_golden.py            # all fixtures (several minutes)
    python tests/golden/make_golden.py --quick    # only the cheap ones
    python tests/golden/make_golden.py --config3  # only BASELINE config 3 at full size (slow)
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ipp as oipp                      # noqa: E402
from oracle import r1cs as or1cs                    # noqa: E402
from oracle.curves import BLS12_381, BN254          # noqa: E402
from oracle.merlin import Transcript                # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def ipp_bytes(C, pr):
    return (b"".join(C.g1_to_bytes(p) for p in pr.L) + b"".join(C.g1_to_bytes(p) for p in pr.R)
            + C.fr_to_bytes(pr.a) + C.fr_to_bytes(pr.b))


def gen_generators():
    out = {}
    for C in (BLS12_381, BN254):
        d = {}
        for prefix in ("G", "H", "g", "h"):
            d[prefix] = [C.g1_xy_bytes(p).hex() for p in C.get_generators(prefix, 4)]
        for msg in (b"g", b"h", b"Q", b""):
            d["msg:" + msg.decode()] = C.g1_xy_bytes(C.g1_from_msg_hash(msg)).hex()
        out[C.name] = d
    return out


def gen_ipp(n):
    out = {}
    for C in (BLS12_381, BN254):
        G, H, Q = C.get_generators("g", n), C.get_generators("h", n), C.g1_from_msg_hash(b"Q")
        a, b = C.synth_scalars(11, n, b"a"), C.synth_scalars(11, n, b"b")
        y_inv = C.fr_inv(C.synth_scalar(9, 0))
        Gf, Hf = [1] * n, C.vandermonde(y_inv, n)
        pr = oipp.create_ipp(C, Transcript(b"innerproduct", C), Q, Gf, Hf, G, H, a, b)
        bp_ = [x * y % C.r for x, y in zip(b, Hf)]
        P = C.msm(G + H + [Q], a + bp_ + [C.inner_product(a, b)])
        out[C.name] = {"n": n, "label": "innerproduct", "a_seed": 11, "y_seed": [9, 0], "proof": ipp_bytes(C, pr).hex(),
                       "P": C.g1_xy_bytes(P).hex()}
    return out


def gen_range(C, m, bits, seed, label):
    g, h = C.g1_from_msg_hash(b"g"), C.g1_from_msg_hash(b"h")
    n = m * bits
    N = 1 << max(0, (n - 1).bit_length())
    G, H = C.get_generators("G", N), C.get_generators("H", N)
    vals = [C.synth_scalar(seed, j, b"v") & ((1 << bits) - 1) for j in range(m)]
    rng = or1cs.make_rng(C, seed)
    p = or1cs.Prover(C, g, h, Transcript(label, C))
    comms = []
    for v in vals:
        com, var = p.commit(v, rng())
        comms.append(com)
        or1cs.positive_no_gadget(p, or1cs.AllocatedQuantity(var, v), bits)
    proof = p.prove(G, H, rng)
    # the oracle's own verifier accepts it
    v_ = or1cs.Verifier(C, Transcript(label, C))
    for com in comms:
        or1cs.positive_no_gadget(v_, or1cs.AllocatedQuantity(v_.commit(com), None), bits)
    v_.verify(proof, g, h, G, H, C.synth_scalar(seed, 0, b"verifier"))
    pb = proof.to_bytes(C)
    return {"curve": C.name, "m": m, "bits": bits, "seed": seed, "label": label.decode(), "values": [str(v) for v in vals],
            "commitments": b"".join(C.g1_xy_bytes(c) for c in comms).hex(), "proof": pb.hex(),
            "proof_sha256": hashlib.sha256(pb).hexdigest()}


def gen_bound(C, bits, seed):
    g, h = C.g1_from_msg_hash(b"g"), C.g1_from_msg_hash(b"h")
    G, H = C.get_generators("G", 2 * bits), C.get_generators("H", 2 * bits)
    rng = or1cs.make_rng(C, seed)
    p = or1cs.Prover(C, g, h, Transcript(b"BoundsTest", C))
    comms = or1cs.prove_bounded_num(p, 75, rng(), 10, 100, bits, rng)
    proof = p.prove(G, H, rng)
    return {"curve": C.name, "bits": bits, "seed": seed, "val": 75, "lower": 10, "upper": 100, "label": "BoundsTest",
            "commitments": b"".join(C.g1_xy_bytes(c) for c in comms).hex(), "proof": proof.to_bytes(C).hex()}


def gen_msm():
    out = {}
    for C in (BLS12_381, BN254):
        n = 257
        G = C.from_affine(C.g)
        pts, cur = [], C.mul(G, 7)
        for _ in range(n):
            pts.append(cur)
            cur = C.add(cur, G)
        s = C.synth_scalars(1, n)
        out[C.name] = {"n": n, "points": "P_i = (7+i)*G_std", "scalar_seed": 1, "result": C.g1_xy_bytes(C.msm(pts, s)).hex()}
    return out


def write(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=1)
    print("wrote", name, flush=True)


def main():
    quick = "--quick" in sys.argv
    t0 = time.time()
    if "--config3" in sys.argv:          # BASELINE config 3 at full size: 2^14 multipliers on BN254 (tens of minutes in Python)
        write("range_config3.json", {"bn_m256_b64": gen_range(BN254, 256, 64, 7, b"Range")})
        print("done in %.1f s" % (time.time() - t0))
        return
    write("generators.json", gen_generators())
    write("msm_257.json", gen_msm())
    write("ipp_n8.json", gen_ipp(8))
    write("bound_check_8bit.json", {C.name: gen_bound(C, 8, 1) for C in (BLS12_381, BN254)})
    write("range_small.json", {"bls_m2_b8": gen_range(BLS12_381, 2, 8, 3, b"Range"), "bn_m3_b5": gen_range(BN254, 3, 5, 4, b"Range")})
    if not quick:
        write("ipp_n64.json", gen_ipp(64))                                                   # BASELINE config 1
        write("range_config5_unit.json", {"bls_m1_b64": gen_range(BLS12_381, 1, 64, 5, b"Range")})   # config 5 unit
        write("range_config2.json", {"bls_m16_b64": gen_range(BLS12_381, 16, 64, 2, b"Range")})      # config 2
        write("range_config3_reduced.json", {"bn_m8_b64": gen_range(BN254, 8, 64, 6, b"Range")})     # config 3 at n = 512
    print("done in %.1f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
