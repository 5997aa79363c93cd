"""Latency of prove / verify through the host layer for the BASELINE.json configurations (single proof, one ctx)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofs_amcl_b200 as bp


def t(fn, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); best = min(best, time.perf_counter() - t0)
    return best * 1e3, r


PRE = os.environ.get("PRE", "1") == "1"


def main():
    for curve, name in ((bp.BLS12_381, "bls"), (bp.BN254, "bn")):
        ctx = bp.Context(curve, 0)
        mb = ctx.modbytes
        gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
        for m, bits in ((1, 64), (16, 64), (256, 64)):
            n = m * bits
            t0 = time.perf_counter()
            G, H = ctx.get_generators("G", n, precompute=PRE and n <= 4096), ctx.get_generators("H", n, precompute=PRE and n <= 4096)
            tg = (time.perf_counter() - t0) * 1e3
            vals = [(12345678901234567 * (i + 1)) & ((1 << 64) - 1) for i in range(m)]
            l0 = ctx.launches
            tp, (proof, comms) = t(lambda: ctx.range_prove(b"bench", gx, hx, G, H, vals, bits, seed=1))
            l1 = ctx.launches
            tv, ok = t(lambda: ctx.range_verify(b"bench", gx, hx, G, H, m, bits, proof, comms))
            l2 = ctx.launches
            print(f"{name} range m={m} bits={bits} n={n}: gens {tg:.1f} ms  prove {tp:.2f} ms ({(l1-l0)//4} launches)  verify {tv:.2f} ms ({(l2-l1)//4} launches) ok={ok} proof={len(proof)}B", flush=True)
        ctx.close()


main()
