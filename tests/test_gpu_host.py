"""GPU parity of the host layer (include/bphost.h): the reference's own tests replayed end to end through the
C++ mirror of its API -- ipp.rs:318-490 (test_ipp, test_ipp_non_power_of_2), gadgets/bound_check.rs:188-225
(test_bound_check_gadget) -- and compared BYTE FOR BYTE with the oracle's restatement on the same seeds."""
import pytest

from oracle import ipp as oipp
from oracle import r1cs as or1cs
from oracle.merlin import Transcript
from tests.util import curve_of, enc_points, enc_scalars

pytestmark = pytest.mark.gpu


def ipp_bytes(C, pr):
    return (b"".join(C.g1_to_bytes(p) for p in pr.L) + b"".join(C.g1_to_bytes(p) for p in pr.R)
            + C.fr_to_bytes(pr.a) + C.fr_to_bytes(pr.b))


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_get_generators_matches_oracle(which, ctx_bls, ctx_bn):
    """utils::get_generators / G1::from_msg_hash: hash on the host, map-to-curve on the device."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    for prefix, n in (("G", 40), ("h", 3), ("g", 0)):
        t = ctx.get_generators(prefix, n)
        assert t.download() == enc_points(C, C.get_generators(prefix, n))
        t.free()
    for msg in (b"g", b"h", b"Q", b"", b"a longer message that still hashes to one point"):
        assert ctx.g1_from_msg_hash(msg) == C.g1_xy_bytes(C.g1_from_msg_hash(msg))


@pytest.mark.parametrize("which", ["bls", "bn"])
@pytest.mark.parametrize("case", ["n4", "n8_padded", "n64", "n1"])
def test_ipp_like_reference_tests(which, case, ctx_bls, ctx_bn):
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    if case == "n4":
        a, b, k = [1, 2, 3, 4], [5, 6, 7, 8], 4                                   # ipp.rs:326-335
    elif case == "n8_padded":
        a, b, k = [1, 2, 3, 4, 9, 0, 0, 0], [5, 6, 7, 8, 10, 0, 0, 0], 5          # ipp.rs:395-402
    elif case == "n1":
        a, b, k = [7], [9], 1
    else:
        a, b, k = C.synth_scalars(11, 64, b"a"), C.synth_scalars(11, 64, b"b"), 64   # BASELINE config 1
    n = len(a)
    dG, dH = ctx.get_generators("g", n), ctx.get_generators("h", n)                # ipp.rs:339-341
    G, H, Q = C.get_generators("g", n), C.get_generators("h", n), C.g1_from_msg_hash(b"Q")
    y_inv = C.fr_inv(C.synth_scalar(9, 0))
    Gf, Hf = [1] * n, C.vandermonde(y_inv, n)                                      # ipp.rs:344-348
    exp = oipp.create_ipp(C, Transcript(b"innerproduct", C), Q, Gf, Hf, G, H, a, b)
    proof = ctx.ipp_create(b"innerproduct", dG, dH, C.g1_xy_bytes(Q), enc_scalars(C, Gf), enc_scalars(C, Hf), enc_scalars(C, a),
                           enc_scalars(C, b), n)
    assert proof == ipp_bytes(C, exp)
    # P = <a, G> + <b', H> + <a,b> Q with b' = b o H_factors, from the un-padded vectors (ipp.rs:354-372,458-471)
    bp_ = [x * y % C.r for x, y in zip(b[:k], Hf)]
    P = C.msm(G[:k] + H[:k] + [Q], a[:k] + bp_ + [C.inner_product(a, b)])
    args = (b"innerproduct", n, enc_scalars(C, Gf), enc_scalars(C, Hf), C.g1_xy_bytes(P), C.g1_xy_bytes(Q), dG, dH)
    assert ctx.ipp_verify(*args, proof) is True
    mb = C.MODBYTES
    bad = bytearray(proof)
    bad[-1] ^= 1                                                                   # b -> b +- 1
    assert ctx.ipp_verify(*args, bytes(bad)) is False
    if n > 1:
        bad = proof[:1] + proof[1 + 2 * mb + 1:1 + 4 * mb + 1] + proof[1 + 2 * mb:]    # L_1 replaced by L_2's bytes (shifted)
        assert len(bad) == len(proof)
        assert ctx.ipp_verify(*args, bad) is False
        # verification_scalars' guard n == 2^lg (ipp.rs:274-276)
        assert ctx.ipp_verify(b"innerproduct", 2 * n, enc_scalars(C, Gf + Gf), enc_scalars(C, Hf + Hf), C.g1_xy_bytes(P),
                              C.g1_xy_bytes(Q), ctx.get_generators("g", 2 * n), ctx.get_generators("h", 2 * n), proof) is False
    with pytest.raises(Exception):                                                 # assert!(n.is_power_of_two()) ipp.rs:48
        ctx.ipp_create(b"innerproduct", dG, dH, C.g1_xy_bytes(Q), enc_scalars(C, Gf[:3]), enc_scalars(C, Hf[:3]),
                       enc_scalars(C, a[:3]), enc_scalars(C, b[:3]), 3)


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_bound_check_matches_oracle(which, ctx_bls, ctx_bn):
    """gadgets/bound_check.rs:188-225 at 8 bits: proof bytes identical to the oracle's on the same blinding stream."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    bits, seed = 8, 1
    g, h = C.g1_from_msg_hash(b"g"), C.g1_from_msg_hash(b"h")
    G, H = C.get_generators("G", 2 * bits), C.get_generators("H", 2 * bits)
    rng = or1cs.make_rng(C, seed)
    p = or1cs.Prover(C, g, h, Transcript(b"BoundsTest", C))
    comms = or1cs.prove_bounded_num(p, 75, rng(), 10, 100, bits, rng)
    exp = p.prove(G, H, rng)
    dG, dH = ctx.get_generators("G", 2 * bits), ctx.get_generators("H", 2 * bits)
    gx, hx = C.g1_xy_bytes(g), C.g1_xy_bytes(h)
    proof, cb = ctx.bound_check_prove(b"BoundsTest", gx, hx, dG, dH, 75, 10, 100, bits, seed=seed)
    assert cb == enc_points(C, comms)
    assert proof == exp.to_bytes(C)
    r_be = C.fr_to_bytes(C.synth_scalar(77, 0))
    assert ctx.bound_check_verify(b"BoundsTest", gx, hx, dG, dH, 10, 100, bits, proof, cb, r_be) is True
    assert ctx.bound_check_verify(b"BoundsTest", gx, hx, dG, dH, 10, 100, bits, proof, cb) is True          # OS-entropy r
    # wrong statement / wrong transcript label / tampered scalar / tampered commitment -> VerificationError
    assert ctx.bound_check_verify(b"BoundsTest", gx, hx, dG, dH, 11, 100, bits, proof, cb, r_be) is False
    assert ctx.bound_check_verify(b"OtherLabel", gx, hx, dG, dH, 10, 100, bits, proof, cb, r_be) is False
    mb = C.MODBYTES
    off = 11 * (2 * mb + 1) + mb - 1                                              # last byte of t_x
    bad = bytearray(proof)
    bad[off] ^= 1
    assert ctx.bound_check_verify(b"BoundsTest", gx, hx, dG, dH, 10, 100, bits, bytes(bad), cb, r_be) is False
    assert ctx.bound_check_verify(b"BoundsTest", gx, hx, dG, dH, 10, 100, bits, proof, cb[2 * mb:4 * mb] + cb[:2 * mb] + cb[4 * mb:],
                                  r_be) is False
    # InvalidGeneratorsLength (prover.rs:332-334, verifier.rs:297-299)
    small = ctx.get_generators("G", 5)
    with pytest.raises(Exception) as e:
        ctx.bound_check_prove(b"BoundsTest", gx, hx, small, small, 75, 10, 100, bits, seed=seed)
    assert e.value.code == -3
    with pytest.raises(Exception) as e:
        ctx.bound_check_verify(b"BoundsTest", gx, hx, small, small, 10, 100, bits, proof, cb, r_be)
    assert e.value.code == -3
    # FormatError on a truncated proof
    with pytest.raises(Exception) as e:
        ctx.bound_check_verify(b"BoundsTest", gx, hx, dG, dH, 10, 100, bits, proof[:-3], cb, r_be)
    assert e.value.code == -5


def test_bound_check_reference_shape_32_bits(ctx_bls):
    """the reference's own parameters: min 10, max 100, n = 32 bits, 128 generators (bound_check.rs:195-210): completeness
    with OS randomness, and an out-of-range witness must not verify."""
    ctx = ctx_bls
    C = curve_of(ctx)
    dG, dH = ctx.get_generators("G", 128), ctx.get_generators("H", 128)
    gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
    for v in (10, 57, 100):
        proof, cb = ctx.bound_check_prove(b"BoundsTest", gx, hx, dG, dH, v, 10, 100, 32)
        assert ctx.bound_check_verify(b"BoundsTest", gx, hx, dG, dH, 10, 100, 32, proof, cb) is True
    assert len(proof) == 11 * 97 + 3 * 48 + 2 * 6 * 97 + 2 * 48                      # proof.rs:26-58 with lg = 6


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_range_proof_matches_oracle(which, ctx_bls, ctx_bn):
    """m x positive_no_gadget in one constraint system (BASELINE configs 2/3/5 at reduced size), incl. a padded circuit
    (n = 3*5 = 15 -> 16, prover.rs:527-535) and a value that is out of range."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    g, h = C.g1_from_msg_hash(b"g"), C.g1_from_msg_hash(b"h")
    gx, hx = C.g1_xy_bytes(g), C.g1_xy_bytes(h)
    for vals, bits, seed in (([5, 200], 8, 3), ([1, 30, 17], 5, 4)):
        n = len(vals) * bits
        N = 1 << (n - 1).bit_length()
        G, H = C.get_generators("G", N), C.get_generators("H", N)
        rng = or1cs.make_rng(C, seed)
        p = or1cs.Prover(C, g, h, Transcript(b"Range", C))
        comms = []
        for v in vals:
            com, var = p.commit(v, rng())
            comms.append(com)
            or1cs.positive_no_gadget(p, or1cs.AllocatedQuantity(var, v), bits)
        exp = p.prove(G, H, rng)
        dG, dH = ctx.get_generators("G", N), ctx.get_generators("H", N)
        proof, cb = ctx.range_prove(b"Range", gx, hx, dG, dH, vals, bits, seed=seed)
        assert cb == enc_points(C, comms)
        assert proof == exp.to_bytes(C)
        assert ctx.range_verify(b"Range", gx, hx, dG, dH, len(vals), bits, proof, cb) is True
    # 300 does not fit 8 bits: the prover still emits a proof (as the reference would) but it must not verify
    dG, dH = ctx.get_generators("G", 16), ctx.get_generators("H", 16)
    proof, cb = ctx.range_prove(b"Range", gx, hx, dG, dH, [5, 300], 8, seed=9)
    assert ctx.range_verify(b"Range", gx, hx, dG, dH, 2, 8, proof, cb) is False


def test_batch_prove_verify_keeps_per_proof_verdicts(bp, ctx_bls):
    """BASELINE config 5 at reduced size: independent range proofs spread over several contexts (host threads / CUDA
    streams) sharing one generator table; every proof keeps its own verdict (verifier.rs:267 has no batch API)."""
    C = curve_of(ctx_bls)
    ctxs = [ctx_bls] + [bp.Context(bp.BLS12_381, 0) for _ in range(2)]
    bits, m, count = 8, 1, 12
    dG, dH = ctx_bls.get_generators("G", m * bits), ctx_bls.get_generators("H", m * bits)
    gx, hx = ctx_bls.g1_from_msg_hash(b"g"), ctx_bls.g1_from_msg_hash(b"h")
    values = [(37 * i + 5) % 256 for i in range(count)]
    proofs, stride, comms = bp.range_prove_many(ctxs, b"Batch", gx, hx, dG, dH, values, m, bits, seed=100)
    assert stride == 11 * 97 + 3 * 48 + 2 * 3 * 97 + 2 * 48
    # proof i is exactly what a single call with seed 100+i produces
    for i in (0, 7):
        p1, c1 = ctx_bls.range_prove(b"Batch", gx, hx, dG, dH, [values[i]], bits, seed=100 + i)
        assert p1 == proofs[i * stride:(i + 1) * stride] and c1 == comms[i * 2 * 48:(i + 1) * 2 * 48]
    assert bp.range_verify_many(ctxs, b"Batch", gx, hx, dG, dH, count, m, bits, proofs, stride, comms) == [0] * count
    bad = bytearray(proofs)
    bad[5 * stride + 11 * 97 + 47] ^= 1            # t_x of proof 5
    bad[9 * stride] = 5                            # point tag of proof 9 -> FormatError
    v = bp.range_verify_many(ctxs, b"Batch", gx, hx, dG, dH, count, m, bits, bytes(bad), stride, comms)
    assert v == [0, 0, 0, 0, 0, -4, 0, 0, 0, -5, 0, 0]
    # the same verdicts from ONE device call for the whole batch (bph_range_verify_batch); the tables it needs are built on
    # first use, and the generator-length guard (verifier.rs:297-299) applies per proof
    dG2, dH2 = ctx_bls.get_generators("G", m * bits), ctx_bls.get_generators("H", m * bits)
    assert bp.range_verify_batch(ctx_bls, b"Batch", gx, hx, dG2, dH2, count, m, bits, proofs, stride, comms) == [0] * count
    assert bp.range_verify_batch(ctx_bls, b"Batch", gx, hx, dG2, dH2, count, m, bits, bytes(bad), stride, comms, nthreads=3) == v
    assert bp.range_verify_batch(ctx_bls, b"Other", gx, hx, dG2, dH2, 3, m, bits, proofs, stride, comms) == [-4] * 3
    swapped = comms[2 * 48:4 * 48] + comms[:2 * 48] + comms[4 * 48:]
    assert bp.range_verify_batch(ctx_bls, b"Batch", gx, hx, dG2, dH2, 3, m, bits, proofs, stride, swapped) == [-4, -4, 0]
    small = ctx_bls.get_generators("G", 4)
    assert bp.range_verify_batch(ctx_bls, b"Batch", gx, hx, small, small, 2, m, bits, proofs, stride, comms) == [-3, -3]
    assert bp.range_verify_batch(ctx_bls, b"Batch", gx, hx, dG2, dH2, 0, m, bits, b"", stride, b"") == []
    for c in ctxs[1:]:
        c.close()


@pytest.mark.parametrize("which", ["bls", "bn"])
@pytest.mark.parametrize("k,bits,pre", [(2, 0, False), (3, 4, False), (5, 8, True)])
def test_two_phase_shuffle_matches_oracle(which, k, bits, pre, ctx_bls, ctx_bn):
    """A circuit with RANDOMISED constraints (the shuffle example of constraint_system.rs:86-135): first-phase commitments,
    2-phase domain separator, challenge z, 2(k-1) second-phase multipliers, A_I2 / A_O2 / S2 (prover.rs:300-319,384-436;
    verifier.rs:245-264), n padded to a power of two.  Proof bytes equal the oracle's; a non-permutation does not verify."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    C = curve_of(ctx)
    g, h = C.g1_from_msg_hash(b"g"), C.g1_from_msg_hash(b"h")
    gx, hx = C.g1_xy_bytes(g), C.g1_xy_bytes(h)
    n = bits + 2 * (k - 1)
    N = 1 << max(0, (n - 1).bit_length())
    G, H = C.get_generators("G", N), C.get_generators("H", N)
    dG, dH = ctx.get_generators("G", N, precompute=pre), ctx.get_generators("H", N, precompute=pre)
    xs = [(7 + 13 * i) % (1 << max(bits, 6)) for i in range(k)]
    if bits:
        xs[0] %= 1 << bits
    ys = xs[1:] + xs[:1]                                     # a rotation: a permutation
    seed = 40 + k
    rng = or1cs.make_rng(C, seed)
    p = or1cs.Prover(C, g, h, Transcript(b"Shuffle", C))
    comms, vars_ = [], []
    blinds = [rng() for _ in range(2 * k)]
    for v, bl in zip(xs + ys, blinds):
        com, var = p.commit(v, bl)
        comms.append(com)
        vars_.append(var)
    if bits:
        or1cs.positive_no_gadget(p, or1cs.AllocatedQuantity(vars_[0], xs[0]), bits)
    or1cs.shuffle_gadget(p, vars_[:k], vars_[k:])
    exp = p.prove(G, H, rng)
    assert not C.is_inf(exp.A_I2) or k == 1                  # the second phase really committed something
    # the oracle's verifier accepts its own proof
    v = or1cs.Verifier(C, Transcript(b"Shuffle", C))
    vv = [v.commit(c) for c in comms]
    if bits:
        or1cs.positive_no_gadget(v, or1cs.AllocatedQuantity(vv[0], None), bits)
    or1cs.shuffle_gadget(v, vv[:k], vv[k:])
    v.verify(exp, g, h, G, H, C.synth_scalar(3, 0))
    proof, cb = ctx.shuffle_prove(b"Shuffle", gx, hx, dG, dH, xs, ys, bits, seed=seed)
    assert cb == enc_points(C, comms)
    assert proof == exp.to_bytes(C)
    assert ctx.shuffle_verify(b"Shuffle", gx, hx, dG, dH, k, bits, proof, cb) is True
    # not a permutation: the prover still emits a proof, the verifier rejects it
    bad_y = list(ys)
    bad_y[0] += 1
    proof2, cb2 = ctx.shuffle_prove(b"Shuffle", gx, hx, dG, dH, xs, bad_y, bits, seed=seed)
    assert ctx.shuffle_verify(b"Shuffle", gx, hx, dG, dH, k, bits, proof2, cb2) is False
    # a proof for other commitments does not transfer
    assert ctx.shuffle_verify(b"Shuffle", gx, hx, dG, dH, k, bits, proof, cb2) is False


def test_batch_verification_on_bn254_and_larger_circuits(bp, ctx_bn):
    """bph_range_verify_batch beyond the config-5 shape: AMCL BN254, two committed values per proof (n = 16, lg = 4),
    a batch that is not a multiple of anything, one tampered proof."""
    ctx = ctx_bn
    bits, m, count = 8, 2, 11
    dG, dH = ctx.get_generators("G", m * bits), ctx.get_generators("H", m * bits)
    gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
    values = [(29 * i + 3) % 256 for i in range(count * m)]
    proofs, stride, comms = bp.range_prove_many([ctx], b"BatchBn", gx, hx, dG, dH, values, m, bits, seed=7)
    assert bp.range_verify_many([ctx], b"BatchBn", gx, hx, dG, dH, count, m, bits, proofs, stride, comms) == [0] * count
    assert bp.range_verify_batch(ctx, b"BatchBn", gx, hx, dG, dH, count, m, bits, proofs, stride, comms) == [0] * count
    bad = bytearray(proofs)
    bad[3 * stride + stride - 1] ^= 1              # the IPP's b of proof 3
    assert bp.range_verify_batch(ctx, b"BatchBn", gx, hx, dG, dH, count, m, bits, bytes(bad), stride, comms, nthreads=2) == \
        [0, 0, 0, -4] + [0] * (count - 4)


@pytest.mark.parametrize("which,m,bits,count", [("bls", 1, 8, 9), ("bn", 2, 8, 5), ("bls", 3, 5, 4), ("bls", 1, 64, 3)])
def test_lock_step_batch_prover_equals_single_proofs(which, m, bits, count, bp, ctx_bls, ctx_bn):
    """bph_range_prove_batch: every stage of the prover and every IPP round as ONE device call for the whole batch.  Proof i
    must be byte-identical to the single-context proof with seed + i (which is itself oracle-checked), also for a padded
    circuit (n = 15 -> 16), and the batch verifier accepts them."""
    ctx = ctx_bls if which == "bls" else ctx_bn
    n = m * bits
    N = 1 << max(0, (n - 1).bit_length())
    dG, dH = ctx.get_generators("G", N), ctx.get_generators("H", N)
    gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
    values = [(0x9E3779B97F4A7C15 * (i + 1)) % (1 << bits) for i in range(count * m)]
    ref_p, stride, ref_c = bp.range_prove_many([ctx], b"Batch", gx, hx, dG, dH, values, m, bits, seed=300)
    for mode in (0, 1):                       # transcripts on the device / on host threads
        got_p, stride2, got_c = bp.range_prove_batch(ctx, b"Batch", gx, hx, dG, dH, values, m, bits, seed=300, nthreads=3, mode=mode)
        assert stride2 == stride
        assert got_c == ref_c, mode
        for i in range(count):
            assert got_p[i * stride:(i + 1) * stride] == ref_p[i * stride:(i + 1) * stride], (mode, i)
    assert bp.range_verify_batch(ctx, b"Batch", gx, hx, dG, dH, count, m, bits, got_p, stride, got_c) == [0] * count
    # OS-entropy blindings: proofs differ from run to run and verify
    p1, _, c1 = bp.range_prove_batch(ctx, b"Batch", gx, hx, dG, dH, values, m, bits)
    assert p1 != got_p
    assert bp.range_verify_many([ctx], b"Batch", gx, hx, dG, dH, count, m, bits, p1, stride, c1) == [0] * count


def test_lock_step_batch_prover_two_drivers(bp, ctx_bls):
    """4100 proofs: the batch is split over two interleaved drivers (two contexts, alternating slabs of 1024, a short last
    slab); still byte-identical to the single-context proofs with seed + i."""
    ctx = ctx_bls
    m, bits, count = 1, 8, 4100
    dG, dH = ctx.get_generators("G", 8), ctx.get_generators("H", 8)
    gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
    values = [(37 * i + 11) % 256 for i in range(count)]
    ctxs = [ctx] + [bp.Context(bp.BLS12_381, 0) for _ in range(3)]
    ref_p, stride, ref_c = bp.range_prove_many(ctxs, b"Batch2", gx, hx, dG, dH, values, m, bits, seed=9000)
    for mode in (0, 1):
        got_p, _, got_c = bp.range_prove_batch(ctx, b"Batch2", gx, hx, dG, dH, values, m, bits, seed=9000, mode=mode)
        assert got_c == ref_c and got_p == ref_p, mode
    assert bp.range_verify_batch(ctx, b"Batch2", gx, hx, dG, dH, count, m, bits, got_p, stride, got_c) == [0] * count
    for c in ctxs[1:]:
        c.close()


@pytest.mark.parametrize("which", ["bls", "bn"])
def test_device_transcript_batch_prover_warp_rows(which, bp, ctx_bls, ctx_bn):
    """1100 proofs of 32 multipliers: the slab is large enough for the one-warp-per-row table sums (>= 2048 rows of >= 32
    terms) in the A_I / A_O / S stage and in every IPP round; records byte-identical to single proofs, in both modes"""
    ctx = ctx_bls if which == "bls" else ctx_bn
    m, bits, count = 1, 32, 1100
    dG, dH = ctx.get_generators("G", 32), ctx.get_generators("H", 32)
    gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
    values = [(0x9E3779B97F4A7C15 * (i + 3)) % (1 << bits) for i in range(count)]
    ctxs = [ctx] + [bp.Context(ctx.curve, 0) for _ in range(3)]
    ref_p, stride, ref_c = bp.range_prove_many(ctxs, b"Warp", gx, hx, dG, dH, values, m, bits, seed=77000)
    for mode in (0, 1):
        got_p, _, got_c = bp.range_prove_batch(ctx, b"Warp", gx, hx, dG, dH, values, m, bits, seed=77000, mode=mode)
        assert got_c == ref_c, mode
        bad = [i for i in range(count) if got_p[i * stride:(i + 1) * stride] != ref_p[i * stride:(i + 1) * stride]]
        assert not bad, (mode, bad[:5])
    assert bp.range_verify_batch(ctx, b"Warp", gx, hx, dG, dH, count, m, bits, got_p, stride, got_c) == [0] * count
    for c in ctxs[1:]:
        c.close()


@pytest.mark.gpu
def test_padded_large_circuit_same_bytes_on_every_prover_path(ctx_bn, monkeypatch):
    """m = 100 values x 64 bits = 6400 multipliers, padded to N = 8192 (prover.rs:527-535) on BN254 with window tables: the
    proof bytes do not depend on the path -- gadget per proof or recorded circuit (BPH_RANGE_RECORDED), every IPP round on
    table sums or generators materialised after 4 rounds (BPGPU_IPP_HYBRID) -- and the verifier (its own MSM path) accepts
    them on both of its paths; a tampered value commitment is rejected."""
    ctx = ctx_bn
    m, bits = 100, 64
    N = 8192
    gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
    dG, dH = ctx.get_generators("G", N, precompute=True), ctx.get_generators("H", N, precompute=True)
    vals = [(0x9E3779B97F4A7C15 * (i + 1)) & ((1 << 64) - 1) for i in range(m)]
    seen = set()
    for rec in ("0", "1"):
        for hyb in ("0", "4", "2"):
            monkeypatch.setenv("BPH_RANGE_RECORDED", rec)
            monkeypatch.setenv("BPGPU_IPP_HYBRID", hyb)
            proof, comms = ctx.range_prove(b"padded", gx, hx, dG, dH, vals, bits, seed=77)
            seen.add((proof, comms))
            assert ctx.range_verify(b"padded", gx, hx, dG, dH, m, bits, proof, comms) is True
    assert len(seen) == 1
    proof, comms = next(iter(seen))
    mb2 = len(comms) // m
    swapped = comms[mb2:2 * mb2] + comms[:mb2] + comms[2 * mb2:]
    assert ctx.range_verify(b"padded", gx, hx, dG, dH, m, bits, proof, swapped) is False
    dG.free()
    dH.free()
