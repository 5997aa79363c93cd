"""Parameter sweep of the MSM pipeline (window bits, chunk size, reduction fan-in) through the BPGPU_* overrides.
Each configuration runs in a fresh process because the overrides are read once."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lg = sys.argv[1] if len(sys.argv) > 1 else "20"
configs = [{}] + [json.loads(a) for a in sys.argv[2:]]
for cfg in configs:
    env = dict(os.environ)
    env.update({k: str(v) for k, v in cfg.items()})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "4", "--warmup", "2", "--no-proofs", "--no-cpu-baseline", "--no-sweep", "--lg", lg],
                       capture_output=True, text=True, env=env)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        st = {k: round(v, 3) for k, v in d["stages_ms"].items()}
        print(cfg, "ms/step", round(d["ms_per_step"], 3), st, flush=True)
    except Exception as e:
        print(cfg, "FAILED", r.stderr[-500:], flush=True)
