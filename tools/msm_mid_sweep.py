"""Wall time per MSM (device pipeline + host finish) for mid-size n under BPGPU_C / BPGPU_S overrides."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time, os
sys.path.insert(0, %r)
import bulletproofs_amcl_b200 as bp, numpy as np
n = int(sys.argv[1]); curve = bp.BLS12_381 if sys.argv[2] == "bls" else bp.BN254
ctx = bp.Context(curve, 0); mb = ctx.modbytes
G = ctx.get_generators("G", n)
rng = np.random.default_rng(1)
raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); raw[:, 0] &= 0x1F
sc = np.zeros((n, mb), dtype=np.uint8); sc[:, mb-32:] = raw
ds = ctx.upload_scalars(sc.tobytes())
for _ in range(3): ctx.msm_device(G, ds, n=n)
t0 = time.perf_counter()
for _ in range(20): ctx.msm_device(G, ds, n=n)
print((time.perf_counter() - t0) / 20 * 1e3)
''' % ROOT
n, curve = sys.argv[1], sys.argv[2]
for cfg in [{}] + [json.loads(a) for a in sys.argv[3:]]:
    env = dict(os.environ); env.update({k: str(v) for k, v in cfg.items()})
    r = subprocess.run([sys.executable, "-c", CHILD, n, curve], capture_output=True, text=True, env=env)
    print(n, curve, cfg, r.stdout.strip() or r.stderr[-300:], flush=True)
