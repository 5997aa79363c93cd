"""Golden fixtures (tests/golden/*.json, made by tests/golden/make_golden.py from the oracle).

CPU part: the oracle still reproduces the cheap fixtures (guards the checker against drift).
GPU part: the CUDA path reproduces every fixture byte for byte, including the ones the Python oracle needs minutes for
(BASELINE config 1: IPP n = 64; config 2: 16 x 64-bit range proof, n = 1024; config 3 reduced: BN254 n = 512)."""
import json
import os

import pytest

from oracle.curves import CURVES
from tests.util import enc_scalars

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    path = os.path.join(HERE, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated")
    return json.load(open(path))


# ------------------------------------------------------------------ CPU: oracle vs fixtures
def test_oracle_reproduces_generators():
    fx = load("generators.json")
    for cname, d in fx.items():
        C = CURVES[cname]
        for prefix in ("G", "H", "g", "h"):
            assert [C.g1_xy_bytes(p).hex() for p in C.get_generators(prefix, 4)] == d[prefix]
        assert C.g1_xy_bytes(C.g1_from_msg_hash(b"Q")).hex() == d["msg:Q"]


def test_oracle_reproduces_small_proofs():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    assert mg.gen_ipp(8) == load("ipp_n8.json")
    fx = load("bound_check_8bit.json")
    for cname, d in fx.items():
        assert mg.gen_bound(CURVES[cname], 8, 1) == d
    assert mg.gen_msm() == load("msm_257.json")


def test_oracle_reproduces_two_phase_fixture():
    """the two-phase (randomised constraints) shuffle proof: pure-Python oracle and C-accelerated oracle both reproduce the
    committed bytes (the GPU path is compared with the live oracle on the same statement in tests/test_gpu_host.py)."""
    import importlib.util
    from oracle import fast
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    fx = load("shuffle_two_phase.json")
    assert mg.gen_shuffle(CURVES["BN254"], 3, 4, 43) == fx["bn_k3_b4"]
    with fast.c_keccak():
        assert mg.gen_shuffle(fast.FastCurve(CURVES["BLS12_381"]), 3, 4, 43) == fx["bls_k3_b4"]


@pytest.mark.parametrize("name,key,cname,m,seed", [("range_config5_unit.json", "bls_m1_b64", "BLS12_381", 1, 5),
                                                  ("range_config3_reduced.json", "bn_m8_b64", "BN254", 8, 6),
                                                  ("range_config2.json", "bls_m16_b64", "BLS12_381", 16, 2)])
def test_c_accelerated_oracle_reproduces_config_proofs(name, key, cname, m, seed):
    """The fixtures at BASELINE's config sizes took the pure-Python oracle minutes; the C-accelerated oracle
    (oracle/fast.py: same protocol code, group operations and Keccak in oracle/c) must arrive at the same bytes.
    A Python-vs-C cross check of the whole prover and verifier, and the guard for bench.py's CPU proof baseline."""
    import importlib.util
    from oracle import fast
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    fx = load(name)[key]
    with fast.c_keccak():
        got = mg.gen_range(fast.FastCurve(CURVES[cname]), m, 64, seed, b"Range")
    assert got == fx


# ------------------------------------------------------------------ GPU: CUDA path vs fixtures
def _ctx(name, ctx_bls, ctx_bn):
    return ctx_bls if name == "BLS12_381" else ctx_bn


@pytest.mark.gpu
def test_gpu_generators_golden(ctx_bls, ctx_bn):
    fx = load("generators.json")
    for cname, d in fx.items():
        ctx = _ctx(cname, ctx_bls, ctx_bn)
        for prefix in ("G", "H", "g", "h"):
            assert ctx.get_generators(prefix, 4).download().hex() == "".join(d[prefix])
        for msg in ("g", "h", "Q", ""):
            assert ctx.g1_from_msg_hash(msg.encode()).hex() == d["msg:" + msg]


@pytest.mark.gpu
def test_gpu_msm_golden(ctx_bls, ctx_bn):
    fx = load("msm_257.json")
    for cname, d in fx.items():
        ctx = _ctx(cname, ctx_bls, ctx_bn)
        C = CURVES[cname]
        G = C.from_affine(C.g)
        pts, cur = [], C.mul(G, 7)
        for _ in range(d["n"]):
            pts.append(cur)
            cur = C.add(cur, G)
        xy = b"".join(C.g1_xy_bytes(p) for p in pts)
        assert ctx.msm_refs(xy, enc_scalars(C, C.synth_scalars(d["scalar_seed"], d["n"]))).hex() == d["result"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ipp_n8.json", "ipp_n64.json"])
def test_gpu_ipp_golden(name, ctx_bls, ctx_bn):
    fx = load(name)
    for cname, d in fx.items():
        ctx = _ctx(cname, ctx_bls, ctx_bn)
        C = CURVES[cname]
        n = d["n"]
        dG, dH = ctx.get_generators("g", n), ctx.get_generators("h", n)
        Q = ctx.g1_from_msg_hash(b"Q")
        a, b = C.synth_scalars(d["a_seed"], n, b"a"), C.synth_scalars(d["a_seed"], n, b"b")
        Gf, Hf = [1] * n, C.vandermonde(C.fr_inv(C.synth_scalar(*d["y_seed"])), n)
        proof = ctx.ipp_create(d["label"].encode(), dG, dH, Q, enc_scalars(C, Gf), enc_scalars(C, Hf), enc_scalars(C, a), enc_scalars(C, b), n)
        assert proof.hex() == d["proof"]
        assert ctx.ipp_verify(d["label"].encode(), n, enc_scalars(C, Gf), enc_scalars(C, Hf), bytes.fromhex(d["P"]), Q, dG, dH, proof)


@pytest.mark.gpu
def test_gpu_bound_check_golden(ctx_bls, ctx_bn):
    fx = load("bound_check_8bit.json")
    for cname, d in fx.items():
        ctx = _ctx(cname, ctx_bls, ctx_bn)
        bits = d["bits"]
        dG, dH = ctx.get_generators("G", 2 * bits), ctx.get_generators("H", 2 * bits)
        gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
        proof, comms = ctx.bound_check_prove(d["label"].encode(), gx, hx, dG, dH, d["val"], d["lower"], d["upper"], bits, seed=d["seed"])
        assert proof.hex() == d["proof"] and comms.hex() == d["commitments"]
        assert ctx.bound_check_verify(d["label"].encode(), gx, hx, dG, dH, d["lower"], d["upper"], bits, proof, comms)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["range_small.json", "range_config5_unit.json", "range_config2.json", "range_config3_reduced.json",
                                  "range_config3.json"])
def test_gpu_range_proof_golden(name, ctx_bls, ctx_bn):
    fx = load(name)
    for _, d in fx.items():
        ctx = _ctx(d["curve"], ctx_bls, ctx_bn)
        m, bits = d["m"], d["bits"]
        n = m * bits
        N = 1 << max(0, (n - 1).bit_length())
        dG, dH = ctx.get_generators("G", N), ctx.get_generators("H", N)
        gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
        vals = [int(v) for v in d["values"]]
        proof, comms = ctx.range_prove(d["label"].encode(), gx, hx, dG, dH, vals, bits, seed=d["seed"])
        assert comms.hex() == d["commitments"]
        assert proof.hex() == d["proof"]
        assert ctx.range_verify(d["label"].encode(), gx, hx, dG, dH, m, bits, proof, comms)


@pytest.mark.gpu
def test_gpu_two_phase_shuffle_golden(ctx_bls, ctx_bn):
    """the CUDA path reproduces the committed two-phase proof (second-phase commitments A_I2 / A_O2 / S2 included)"""
    fx = load("shuffle_two_phase.json")
    for key, d in fx.items():
        ctx = _ctx(d["curve"], ctx_bls, ctx_bn)
        k, bits = d["k"], d["bits"]
        n = bits + 2 * (k - 1)
        N = 1 << max(0, (n - 1).bit_length())
        G, H = ctx.get_generators("G", N), ctx.get_generators("H", N)
        gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
        proof, comms = ctx.shuffle_prove(d["label"].encode(), gx, hx, G, H, d["x"], d["y"], bits, seed=d["seed"])
        assert comms.hex() == d["commitments"], key
        assert proof.hex() == d["proof"], key
        assert ctx.shuffle_verify(d["label"].encode(), gx, hx, G, H, k, bits, proof, comms) is True
        G.free()
        H.free()
