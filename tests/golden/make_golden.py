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


def gen_shuffle(C, k, bits, seed):
    """two-phase circuit: k-shuffle with x[0] range-checked to `bits` bits (the statement of bph_shuffle_prove, with the
    same value and draw order as tests/test_gpu_host.py::test_two_phase_shuffle_matches_oracle)"""
    g, h = C.g1_from_msg_hash(b"g"), C.g1_from_msg_hash(b"h")
    n = bits + 2 * (k - 1)
    N = 1 << max(0, (n - 1).bit_length())
    G, H = C.get_generators("G", N), C.get_generators("H", N)
    xs = [(7 + 13 * i) % (1 << max(bits, 6)) for i in range(k)]
    if bits:
        xs[0] %= 1 << bits
    ys = xs[1:] + xs[:1]
    rng = or1cs.make_rng(C, seed)
    p = or1cs.Prover(C, g, h, Transcript(b"Shuffle", C))
    blinds = [rng() for _ in range(2 * k)]
    comms, vars_ = [], []
    for v, bl in zip(xs + ys, blinds):
        com, var = p.commit(v, bl)
        comms.append(com)
        vars_.append(var)
    if bits:
        or1cs.positive_no_gadget(p, or1cs.AllocatedQuantity(vars_[0], xs[0]), bits)
    or1cs.shuffle_gadget(p, vars_[:k], vars_[k:])
    proof = p.prove(G, H, rng)
    v_ = or1cs.Verifier(C, Transcript(b"Shuffle", C))
    vv = [v_.commit(c) for c in comms]
    if bits:
        or1cs.positive_no_gadget(v_, or1cs.AllocatedQuantity(vv[0], None), bits)
    or1cs.shuffle_gadget(v_, vv[:k], vv[k:])
    v_.verify(proof, g, h, G, H, C.synth_scalar(seed, 0, b"verifier"))
    pb = proof.to_bytes(C)
    return {"curve": C.name, "k": k, "bits": bits, "seed": seed, "label": "Shuffle", "x": xs, "y": ys,
            "commitments": b"".join(C.g1_xy_bytes(c) for c in comms).hex(), "proof": pb.hex(), "proof_sha256": hashlib.sha256(pb).hexdigest()}


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
    write("shuffle_two_phase.json", {"bls_k3_b4": gen_shuffle(BLS12_381, 3, 4, 43), "bn_k3_b4": gen_shuffle(BN254, 3, 4, 43)})
    if not quick:
        write("ipp_n64.json", gen_ipp(64))                                                   # BASELINE config 1
        write("range_config5_unit.json", {"bls_m1_b64": gen_range(BLS12_381, 1, 64, 5, b"Range")})   # config 5 unit
        write("range_config2.json", {"bls_m16_b64": gen_range(BLS12_381, 16, 64, 2, b"Range")})      # config 2
        write("range_config3_reduced.json", {"bn_m8_b64": gen_range(BN254, 8, 64, 6, b"Range")})     # config 3 at n = 512
    print("done in %.1f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
