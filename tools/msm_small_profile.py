"""Per-stage times (BPGPU_PROFILE=1 in the environment) of small general-path MSMs.  usage: msm_small_profile.py curve n..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofs_amcl_b200 as bp
curve = bp.BN254 if sys.argv[1] == "bn" else bp.BLS12_381
ctx = bp.Context(curve, 0)
for n in [int(x) for x in sys.argv[2:]]:
    G = ctx.get_generators("G", n)
    s = ctx.fr_random(b"k", 0, n)
    for _ in range(3):
        ctx.msm_device(G, s)
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.msm_device(G, s)
    print(f"n={n}: {(time.perf_counter() - t0) / 20 * 1e6:.0f} us per MSM", file=sys.stderr, flush=True)
