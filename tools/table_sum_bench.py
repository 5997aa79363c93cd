"""Latency of a table-path MSM (bpgpu_msm_device over precomputed window tables).  usage: table_sum_bench.py [curve] [n ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofs_amcl_b200 as bp

curve = bp.BN254 if len(sys.argv) > 1 and sys.argv[1] == "bn" else bp.BLS12_381
sizes = [int(x) for x in sys.argv[2:]] or [64, 1024, 4096, 16384]
ctx = bp.Context(curve, 0)
for n in sizes:
    G = ctx.get_generators("G", n, precompute=True)
    s = ctx.fr_random(b"k", 0, n)
    ref = ctx.msm_device(G, s)
    for _ in range(3):
        ctx.msm_device(G, s)
    t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        r = ctx.msm_device(G, s)
    dt = (time.perf_counter() - t0) / reps
    assert r == ref
    print(f"n={n}: {dt * 1e6:.0f} us per table MSM ({n * 32 / dt / 1e9:.2f} G madd/s)", flush=True)
    G.free()
    s.free()
