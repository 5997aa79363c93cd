import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofs_amcl_b200 as bp
ctx = bp.Context(bp.BLS12_381, 0)
for kind, name in ((0, "IMAD.WIDE/s"), (2, "IMAD(lo) instr/s"), (3, "mix instr/s"), (1, "Fq mul SP=0"), (11, "SP=1"), (12, "SP=2"), (13, "SP=3"), (14, "SP=4"), (16, "SP=6"), (20, "BN SP=0"), (22, "BN SP=2")):
    ops, ms = ctx.int_pipe_bench(kind, 2000 if kind in (0, 2, 3) else 200)
    print(f"{name:20s} {ops:.4e} /s  ({ms:.3f} ms)")
