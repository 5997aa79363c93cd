"""C-accelerated oracle (TEST INFRASTRUCTURE): the same restatement of ipp.rs / prover.rs / verifier.rs as
oracle/ipp.py and oracle/r1cs.py, with the group operations (`msm`, `binary_scalar_mul`, `mul`) and the Keccak
permutation routed to oracle/c (Straus wNAF-5 MSM and interleaved-wNAF two-scalar multiplication: the algorithms the
reference gets from amcl_wrapper).

Two uses, both on the checker side only:
  * tests: its proofs must equal the pure-Python oracle's byte for byte (a Python-vs-C cross check of the whole
    protocol, not just of single operations);
  * bench.py's CPU baseline for the proofs/s half of the metric: what the reference's single-threaded prover and
    verifier cost on this box's host cores (an upper bound on its speed: 64-bit Montgomery limbs, var-time MSM also
    where the reference uses the slower constant-time one, prover.rs:347-362).
"""
from . import cref
from . import merlin
from .curves import Curve


class FastCurve(Curve):
    def __init__(self, base):
        self.__dict__.update(base.__dict__)
        self.base = base

    def to_affine(self, P):
        if P[2] == 1:
            return (P[0], P[1])
        return Curve.to_affine(self, P)

    def msm(self, points, scalars):
        if len(points) != len(scalars):
            raise ValueError("UnequalSizeVectors")
        if not points:
            return self.INF
        xy = b"".join(self.g1_xy_bytes(P) for P in points)
        sb = b"".join(self.fr_to_bytes(s % self.r) for s in scalars)
        return self.g1_from_xy_bytes(cref.msm(self.id, xy, sb, len(points), 1))

    def mul(self, P, k):
        if k >= self.r or k < 0 or P[2] == 0:          # cofactor clearing etc.: integers that are not field elements
            return Curve.mul(self, P, k)
        return self.g1_from_xy_bytes(cref.scalar_mul(self.id, self.g1_xy_bytes(P), self.fr_to_bytes(k)))

    def binary_scalar_mul(self, g, h, r1, r2):
        return self.g1_from_xy_bytes(cref.binary_scalar_mul(self.id, self.g1_xy_bytes(g), self.g1_xy_bytes(h),
                                                            self.fr_to_bytes(r1 % self.r), self.fr_to_bytes(r2 % self.r)))


class c_keccak:
    """context manager: oracle.merlin uses the C permutation inside the block"""

    def __enter__(self):
        self.saved = merlin.keccak_f1600
        merlin.keccak_f1600 = cref.keccak_f1600
        return self

    def __exit__(self, *a):
        merlin.keccak_f1600 = self.saved
        return False


def range_prove_verify(C, vals, bits, seed, gens=None, label=b"bench"):
    """One aggregated range proof (len(vals) x positive_no_gadget in one constraint system: BASELINE configs 1/2/3/5)
    proved and verified with the C-accelerated oracle.  Returns (prove_seconds, verify_seconds, proof_bytes, gens)."""
    import time
    from . import r1cs as or1cs
    from .merlin import Transcript
    F = C if isinstance(C, FastCurve) else FastCurve(C)
    n = len(vals) * bits
    N = 1 << max(0, (n - 1).bit_length())
    if gens is None:
        gens = (F.g1_from_msg_hash(b"g"), F.g1_from_msg_hash(b"h"), F.get_generators("G", N), F.get_generators("H", N))
    g, h, G, H = gens
    with c_keccak():
        rng = or1cs.make_rng(F, seed)
        t0 = time.perf_counter()
        p = or1cs.Prover(F, g, h, Transcript(label, F))
        comms = []
        for v in vals:
            com, var = p.commit(v, rng())
            comms.append(com)
            or1cs.positive_no_gadget(p, or1cs.AllocatedQuantity(var, v), bits)
        proof = p.prove(G, H, rng)
        t1 = time.perf_counter()
        vf = or1cs.Verifier(F, Transcript(label, F))
        for com in comms:
            or1cs.positive_no_gadget(vf, or1cs.AllocatedQuantity(vf.commit(com), None), bits)
        vf.verify(proof, g, h, G, H, F.synth_scalar(seed, 0, b"verifier"))
        t2 = time.perf_counter()
    return t1 - t0, t2 - t1, proof.to_bytes(F), gens


def ipp_prove_verify(C, n, seed, gens=None, label=b"ipp"):
    """IPP::create_ipp + IPP::verify_ipp (ipp.rs:35-260, the shape of its own tests ipp.rs:325-390 at length n) with the
    C-accelerated oracle.  Returns (prove_seconds, verify_seconds, gens)."""
    import time
    from . import ipp as oipp
    from .merlin import Transcript
    F = C if isinstance(C, FastCurve) else FastCurve(C)
    if gens is None:
        gens = (F.get_generators("g", n), F.get_generators("h", n), F.g1_from_msg_hash(b"Q"))
    G, H, Q = gens
    a, b = F.synth_scalars(seed, n, b"a"), F.synth_scalars(seed, n, b"b")
    ones = [1] * n
    P = F.msm(G + H + [Q], a + b + [F.inner_product(a, b)])
    with c_keccak():
        t0 = time.perf_counter()
        proof = oipp.create_ipp(F, Transcript(label, F), Q, ones, ones, G, H, a, b)
        t1 = time.perf_counter()
        oipp.verify_ipp(F, n, Transcript(label, F), ones, ones, P, Q, G, H, proof.a, proof.b, proof.L, proof.R)
        t2 = time.perf_counter()
    return t1 - t0, t2 - t1, gens


def _ipp_worker(args):
    cname, n, count, seed = args
    from .curves import CURVES
    F = FastCurve(CURVES[cname])
    gens = None
    tp = tv = 0.0
    for k in range(count):
        a, b, gens = ipp_prove_verify(F, n, seed + k, gens)
        tp += a
        tv += b
    return tp, tv


def _worker(args):
    cname, m, bits, count, seed = args
    import time
    from .curves import CURVES
    F = FastCurve(CURVES[cname])
    gens = None
    tp = tv = 0.0
    for k in range(count):
        vals = [F.synth_scalar(seed + k, j, b"v") & ((1 << bits) - 1) for j in range(m)]
        a, b, _, gens = range_prove_verify(F, vals, bits, seed + k, gens)
        tp += a
        tv += b
    return tp, tv


def main(argv):
    """python -m oracle.fast CURVE m bits proofs_per_process processes -> one JSON line with the reference algorithm's
    prove / verify rate on this machine's cores (generator set-up excluded, as in the GPU timing)."""
    import json
    import multiprocessing as mp
    import time
    ipp = argv[0] == "ipp"                      # python -m oracle.fast ipp CURVE n proofs_per_process processes
    if ipp:
        argv = argv[1:]
        cname, m, bits, count, procs = argv[0], int(argv[1]), 1, int(argv[2]), int(argv[3])
        jobs = [(cname, m, count, 1000 * (i + 1)) for i in range(procs)]
        fn = _ipp_worker
    else:
        cname, m, bits, count, procs = argv[0], int(argv[1]), int(argv[2]), int(argv[3]), int(argv[4])
        jobs = [(cname, m, bits, count, 1000 * (i + 1)) for i in range(procs)]
        fn = _worker
    t0 = time.perf_counter()
    if procs == 1:
        res = [fn(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(fn, jobs)
    wall = time.perf_counter() - t0
    tp, tv = sum(r[0] for r in res), sum(r[1] for r in res)
    total = count * procs
    # per-process rates summed over the processes (set-up, which runs once per process, is not in tp / tv)
    print(json.dumps({"curve": cname, "multipliers": m * bits, "proofs": total, "processes": procs,
                      "prove_per_s": procs * total / tp if tp else None, "verify_per_s": procs * total / tv if tv else None,
                      "prove_verify_per_s": procs * total / (tp + tv) if tp + tv else None, "wall_s": wall}))


if __name__ == "__main__":
    import sys
    main(sys.argv[1:])
