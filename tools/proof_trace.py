import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofs_amcl_b200 as bp
m = int(sys.argv[1]); curve = bp.BLS12_381 if sys.argv[2] == "bls" else bp.BN254
ctx = bp.Context(curve, 0)
gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
n = m * 64
pre = len(sys.argv) <= 3 or sys.argv[3] != "0"
G, H = ctx.get_generators("G", n, precompute=pre), ctx.get_generators("H", n, precompute=pre)
vals = [(12345678901234567 * (i + 1)) & ((1 << 64) - 1) for i in range(m)]
proof, comms = ctx.range_prove(b"bench", gx, hx, G, H, vals, 64, seed=1)
print("=== PROVE", file=sys.stderr, flush=True)
t0 = time.perf_counter(); proof, comms = ctx.range_prove(b"bench", gx, hx, G, H, vals, 64, seed=1); print("prove ms", (time.perf_counter()-t0)*1e3, file=sys.stderr)
print("=== VERIFY", file=sys.stderr, flush=True)
t0 = time.perf_counter(); ok = ctx.range_verify(b"bench", gx, hx, G, H, m, 64, proof, comms); print("verify ms", (time.perf_counter()-t0)*1e3, ok, file=sys.stderr)
for _ in range(2):
    t0 = time.perf_counter(); ok = ctx.range_verify(b"bench", gx, hx, G, H, m, 64, proof, comms); print("verify again ms", (time.perf_counter()-t0)*1e3, ok, file=sys.stderr)
