"""A/B of the fused small-MSM kernel (BPGPU_SMALL=0/1 in the environment): median latency of warm range verifications and of
general-path MSMs at small n.  usage: small_msm_ab.py"""
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofs_amcl_b200 as bp
import numpy as np

for curve, cname in ((bp.BLS12_381, "bls"), (bp.BN254, "bn")):
    ctx = bp.Context(curve, 0)
    gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
    for m in (1, 4, 16):
        n = m * 64
        G, H = ctx.get_generators("G", n, precompute=True), ctx.get_generators("H", n, precompute=True)
        vals = [(12345678901234567 * (i + 1)) & ((1 << 64) - 1) for i in range(m)]
        proof, comms = ctx.range_prove(b"ab", gx, hx, G, H, vals, 64, seed=1)
        ts = []
        for _ in range(30):
            t0 = time.perf_counter()
            ok = ctx.range_verify(b"ab", gx, hx, G, H, m, 64, proof, comms)
            ts.append((time.perf_counter() - t0) * 1e3)
        print(f"{cname} verify n={n}: median {statistics.median(ts[5:]):.3f} ms ok={ok}", flush=True)
    mb = ctx.modbytes
    rng = np.random.default_rng(1)
    for lg in (6, 8, 10, 12):
        n = 1 << lg
        pts = ctx.get_generators("P", n)
        sc = b"".join(int(x).to_bytes(mb, "big") for x in rng.integers(1, 1 << 62, size=n, dtype=np.uint64))
        ts = []
        for _ in range(30):
            t0 = time.perf_counter()
            ctx.msm(pts, sc)
            ts.append((time.perf_counter() - t0) * 1e3)
        print(f"{cname} msm n=2^{lg} (62-bit scalars): median {statistics.median(ts[5:]):.3f} ms", flush=True)
    ctx.close()
