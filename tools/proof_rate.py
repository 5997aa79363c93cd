"""prove / verify rate of one range-proof configuration over 16 contexts.  usage: proof_rate.py m curve count [pre]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofs_amcl_b200 as bp

m = int(sys.argv[1])
curve = bp.BLS12_381 if sys.argv[2] == "bls" else bp.BN254
count = int(sys.argv[3])
pre = len(sys.argv) <= 4 or sys.argv[4] != "0"
nctx = int(sys.argv[5]) if len(sys.argv) > 5 else 16
ctxs = [bp.Context(curve, 0) for _ in range(nctx)]
c0 = ctxs[0]
gx, hx = c0.g1_from_msg_hash(b"g"), c0.g1_from_msg_hash(b"h")
n = m * 64
t0 = time.perf_counter()
G, H = c0.get_generators("G", n, precompute=pre), c0.get_generators("H", n, precompute=pre)
print(f"generators (+tables={pre}): {time.perf_counter() - t0:.2f} s", flush=True)
vals = [(0x9E3779B97F4A7C15 * (i + 1)) & ((1 << 63) - 1) for i in range(count * m)]
bp.range_prove_many(ctxs, b"bench", gx, hx, G, H, vals[:m * 2 * nctx], m, 64)
for rep in range(2):
    t0 = time.perf_counter()
    proofs, stride, comms = bp.range_prove_many(ctxs, b"bench", gx, hx, G, H, vals, m, 64)
    tp = time.perf_counter() - t0
    t0 = time.perf_counter()
    v = bp.range_verify_many(ctxs, b"bench", gx, hx, G, H, count, m, 64, proofs, stride, comms)
    tv = time.perf_counter() - t0
    print(f"n={n} ctxs={nctx}: prove {count / tp:.0f}/s  verify {count / tv:.0f}/s ok={v == [0] * count}", flush=True)
