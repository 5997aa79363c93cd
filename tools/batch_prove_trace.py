"""Stage times of bph_range_prove_batch (BPH_TRACE=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofs_amcl_b200 as bp
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1
count = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ctx = bp.Context(bp.BLS12_381, 0)
gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
G, H = ctx.get_generators("G", m * 64, precompute=True), ctx.get_generators("H", m * 64, precompute=True)
vals = [(0x9E3779B97F4A7C15 * (i + 1)) & ((1 << 64) - 1) for i in range(count * m)]
bp.range_prove_batch(ctx, b"tp", gx, hx, G, H, vals, m, 64)
os.environ["BPH_TRACE"] = "1"
t0 = time.perf_counter()
bp.range_prove_batch(ctx, b"tp", gx, hx, G, H, vals, m, 64)
dt = time.perf_counter() - t0
print(f"{count} proofs {dt*1e3:.1f} ms  {count/dt:.0f}/s", file=sys.stderr)
