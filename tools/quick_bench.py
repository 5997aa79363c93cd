"""Quick on-box numbers: integer-pipe microbenchmarks and MSM timing sweep (not the contract bench)."""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofs_amcl_b200 as bp
from oracle.curves import BLS12_381 as C


def main():
    ctx = bp.Context(bp.BLS12_381, 0)
    for kind, it in ((0, 2000), (1, 200)):
        ops, ms = ctx.int_pipe_bench(kind, it)
        print(f"int_pipe kind={kind}: {ops:.3e} ops/s ({ms:.3f} ms)", flush=True)
    lgs = [int(a) for a in sys.argv[1:]] or [10, 14, 16, 18, 20]
    rnd = random.Random(1)
    nmax = 1 << max(lgs)
    g = C.g1_xy_bytes(C.from_affine(C.g))
    t0 = time.time()
    ks = b"".join(rnd.randrange(1, C.r).to_bytes(48, "big") for _ in range(nmax))
    xy = ctx.selftest_group(2, g * nmax, g * nmax, ks)
    print("gen points s", time.time() - t0, flush=True)
    dp = ctx.upload_points(xy)
    sc = b"".join(rnd.randrange(C.r).to_bytes(48, "big") for _ in range(nmax))
    ds = ctx.upload_scalars(sc)
    for lg in lgs:
        n = 1 << lg
        for _ in range(2):
            ctx.msm_device(dp, ds, n=n)
        t0 = time.time()
        reps = 5
        for _ in range(reps):
            ctx.msm_device(dp, ds, n=n)
        dt = (time.time() - t0) / reps
        print(f"msm 2^{lg}: {dt*1e3:.3f} ms  {n/dt:.3e} points/s  c={bp.lib().bpgpu_msm_window_bits(n)}", flush=True)


if __name__ == "__main__":
    main()
