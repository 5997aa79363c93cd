"""Host / device split of bph_range_verify_batch (BPH_TRACE=1) for `count` 64-bit range proofs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofs_amcl_b200 as bp
count = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nctx = min(16, os.cpu_count() or 1)
ctxs = [bp.Context(bp.BLS12_381, 0) for _ in range(nctx)]
c0 = ctxs[0]
gx, hx = c0.g1_from_msg_hash(b"g"), c0.g1_from_msg_hash(b"h")
G, H = c0.get_generators("G", 64, precompute=True), c0.get_generators("H", 64, precompute=True)
vals = [(0x9E3779B97F4A7C15 * (i + 1)) & ((1 << 64) - 1) for i in range(512)]
proofs, stride, comms = bp.range_prove_many(ctxs, b"tp", gx, hx, G, H, vals, 1, 64)
reps = count // 512
P, Cm = proofs * reps, comms * reps
bp.range_verify_batch(c0, b"tp", gx, hx, G, H, 512, 1, 64, proofs, stride, comms)
for th in (0, 8, 4):
    t0 = time.perf_counter()
    v = bp.range_verify_batch(c0, b"tp", gx, hx, G, H, count, 1, 64, P, stride, Cm, nthreads=th)
    dt = time.perf_counter() - t0
    print(f"threads={th or os.cpu_count()} {count} proofs {dt*1e3:.1f} ms  {count/dt:.0f}/s ok={v == [0]*count}", file=sys.stderr, flush=True)
