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


@pytest.mark.gpu
@pytest.mark.parametrize("recorded", ["0", "1"])
@pytest.mark.parametrize("name,pre", [("range_small.json", False), ("range_config2.json", True), ("range_config3.json", False),
                                      ("range_config3.json", True)])
def test_gpu_range_proof_golden_both_prover_paths(name, pre, recorded, ctx_bls, ctx_bn, monkeypatch):
    """The committed proofs from BOTH single-proof paths of the host layer: the gadget run per proof (as the reference) and the
    circuit recorded once per context with witness and weights built on the device (BPH_RANGE_RECORDED); with window tables
    config 3 also goes through the hybrid IPP (generators materialised after 4 table rounds)."""
    monkeypatch.setenv("BPH_RANGE_RECORDED", recorded)
    fx = load(name)
    for _, d in fx.items():
        ctx = _ctx(d["curve"], ctx_bls, ctx_bn)
        m, bits = d["m"], d["bits"]
        n = m * bits
        N = 1 << max(0, (n - 1).bit_length())
        dG, dH = ctx.get_generators("G", N, precompute=pre), ctx.get_generators("H", N, precompute=pre)
        gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
        vals = [int(v) for v in d["values"]]
        for _rep in range(2):                                # the second proof finds the circuit in the context's cache
            proof, comms = ctx.range_prove(d["label"].encode(), gx, hx, dG, dH, vals, bits, seed=d["seed"])
            assert comms.hex() == d["commitments"]
            assert proof.hex() == d["proof"]
        label = d["label"].encode()
        assert ctx.range_verify(label, gx, hx, dG, dH, m, bits, proof, comms) is True
        bad = bytearray(proof)
        bad[-1] ^= 1                                         # the IPP's b
        assert ctx.range_verify(label, gx, hx, dG, dH, m, bits, bytes(bad), comms) is False
        if m > 1:                                            # commitments swapped: another statement
            mb2 = len(comms) // m
            swapped = comms[mb2:2 * mb2] + comms[:mb2] + comms[2 * mb2:]
            assert ctx.range_verify(label, gx, hx, dG, dH, m, bits, proof, swapped) is False
        dG.free()
        dH.free()


@pytest.mark.gpu
def test_gpu_range_witness_elementwise(ctx_bls, ctx_bn):
    """bpgpu_range_witness against positive_no.rs:18-24: a_L = 1 - bit, a_R = bit, a_O = 0, value-major, LSB first"""
    for ctx in (ctx_bls, ctx_bn):
        mb = 48 if ctx is ctx_bls else 32
        vals, bits = [0, 1, 0xdeadbeefcafef00d, (1 << 64) - 1, 5], 64
        w = ctx.range_witness(vals, bits)
        got = w.download()
        n = len(vals) * bits
        assert len(got) == 3 * n * mb
        for j, v in enumerate(vals):
            for k in range(bits):
                bit = (v >> k) & 1
                i = j * bits + k
                assert int.from_bytes(got[i * mb:(i + 1) * mb], "big") == 1 - bit
                assert int.from_bytes(got[(n + i) * mb:(n + i + 1) * mb], "big") == bit
                assert int.from_bytes(got[(2 * n + i) * mb:(2 * n + i + 1) * mb], "big") == 0
        w.free()
        w7 = ctx.range_witness([0x55, 0x7f], 7).download()     # bits < 64: higher bits of the value are ignored, as shift_right(i) for i < n
        assert [int.from_bytes(w7[(14 + i) * mb:(15 + i) * mb], "big") for i in range(14)] == [1, 0, 1, 0, 1, 0, 1] + [1] * 7
