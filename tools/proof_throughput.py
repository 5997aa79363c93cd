"""Throughput of independent range proofs over several contexts (host threads / streams) of one GPU."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofs_amcl_b200 as bp

curve = bp.BLS12_381 if (len(sys.argv) < 2 or sys.argv[1] == "bls") else bp.BN254
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1
count = int(sys.argv[3]) if len(sys.argv) > 3 else 256
bits = 64
pre = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
nctxs = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else (8, 16, 32, 64)
print("cpus", os.cpu_count())
for nctx in nctxs:
    ctxs = [bp.Context(curve, 0) for _ in range(nctx)]
    c0 = ctxs[0]
    gx, hx = c0.g1_from_msg_hash(b"g"), c0.g1_from_msg_hash(b"h")
    G, H = c0.get_generators("G", m * bits, precompute=pre), c0.get_generators("H", m * bits, precompute=pre)
    vals = [(0x9E3779B97F4A7C15 * (i + 1)) & ((1 << 64) - 1) for i in range(count * m)]
    bp.range_prove_many(ctxs, b"tp", gx, hx, G, H, vals[:nctx * m * 2], m, bits)      # warm-up: tables, scratch
    t0 = time.perf_counter()
    proofs, stride, comms = bp.range_prove_many(ctxs, b"tp", gx, hx, G, H, vals, m, bits)
    tp = time.perf_counter() - t0
    bp.range_verify_many(ctxs, b"tp", gx, hx, G, H, min(count, 2 * nctx), m, bits, proofs, stride, comms)
    t0 = time.perf_counter()
    v = bp.range_verify_many(ctxs, b"tp", gx, hx, G, H, count, m, bits, proofs, stride, comms)
    tv = time.perf_counter() - t0
    assert v == [0] * count
    print(f"pre={pre} nctx={nctx:2d} m={m} n={m*bits}: prove {count/tp:8.1f} proofs/s   verify {count/tv:8.1f} proofs/s", flush=True)
    for c in ctxs:
        c.close()
