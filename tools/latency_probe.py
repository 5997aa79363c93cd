import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofs_amcl_b200 as bp
ctx = bp.Context(bp.BLS12_381, 0)
for kind, name, per in ((31, "BLS Fq mul, 1 warp, 2 interleaved dependent chains", 2), (32, "BN Fq mul", 2), (33, "BLS XYZZ dbl chain", 2)):
    ops, ms = ctx.int_pipe_bench(kind, 2000)
    print(f"{name}: {ms:.3f} ms for 2000 iters -> {ms*1e3/2000/per:.3f} us per op (two per iter)")
ops, ms = ctx.int_pipe_bench(1, 200); print("throughput Fq mul/s", ops)
