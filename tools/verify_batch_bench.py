"""Throughput of the batched verifier in its three modes (device transcripts / host transcripts / round 1 host scalars).
usage: python tools/verify_batch_bench.py [count] [m] [bits] [curve]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofs_amcl_b200 as bp

count = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bits = int(sys.argv[3]) if len(sys.argv) > 3 else 64
curve = bp.BN254 if len(sys.argv) > 4 and sys.argv[4] == "bn" else bp.BLS12_381
ctx = bp.Context(curve, 0)
n = m * bits
G, H = ctx.get_generators("G", n, precompute=True), ctx.get_generators("H", n, precompute=True)
gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
vals = [(0x9E3779B97F4A7C15 * (i + 1)) % (1 << min(bits, 63)) for i in range(count * m)]
t0 = time.perf_counter()
proofs, stride, comms = bp.range_prove_batch(ctx, b"bench", gx, hx, G, H, vals, m, bits)
print(f"prove_batch: {count / (time.perf_counter() - t0):.0f} proofs/s", flush=True)
modes = [int(x) for x in os.environ.get('VB_MODES', '0,1,2').split(',')]
for mode in modes:
    for rep in range(3):
        t0 = time.perf_counter()
        v = bp.range_verify_batch(ctx, b"bench", gx, hx, G, H, count, m, bits, proofs, stride, comms, mode=mode)
        dt = time.perf_counter() - t0
        assert v == [0] * count, (mode, v[:8])
        print(f"mode {mode} rep {rep}: {count / dt:.0f} proofs/s ({dt * 1e3:.1f} ms)", flush=True)
