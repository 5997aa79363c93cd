"""Summarise profiles/launches_r01.csv (ncu launch list) and the key metrics of the full captures into
profiles/launches_r01_summary.txt and profiles/kernels_r01_summary.txt."""
import collections
import csv
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
rows = list(csv.reader(l for l in open(os.path.join(P, "launches_r01.csv")) if not l.startswith("==")))
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki])
    v = float(r[vi].replace(",", ""))
    ms = {"ns": v / 1e6, "us": v / 1e3, "ms": v}.get(r[ui], v * 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
MSM = ("k_digits", "k_scan", "k_scatter", "k_chunk", "k_giant", "k_merge", "k_reduce")
tm = sum(a[1] for k, a in agg.items() if any(x in k for x in MSM))
out = ["# ncu launch list summary (gpu__time_duration.sum, --clock-control none), bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-proofs",
       "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes",
       "%-52s %8s %10s %9s %7s %9s" % ("kernel", "launches", "total ms", "avg ms", "share", "msm share")]
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append("%-52s %8d %10.3f %9.4f %6.1f%% %8s" % (k[:52], c, t, t / c, 100 * t / tot,
                                                      ("%.1f%%" % (100 * t / tm)) if any(x in k for x in MSM) else "-"))
open(os.path.join(P, "launches_r01_summary.txt"), "w").write("\n".join(out) + "\n")
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
lines = ["# key metrics of the ncu --set full captures (one launch each; MSM kernels: 2^20-term BLS12-381 MSM; k_batch_fixed: an IPP round of",
         "# the lock-step batch prover, 2048 rows x 129 terms); raw pages: profiles/k_*_r01_raw.csv"]
for f in ("k_chunk_acc", "k_reduce_l1", "k_scatter", "k_batch_fixed"):
    if not os.path.exists(os.path.join(P, f + "_r01_raw.csv")):
        continue
    rows = list(csv.reader(open(os.path.join(P, f + "_r01_raw.csv"))))
    d = {n: (rows[2][i], rows[1][i]) for i, n in enumerate(rows[0])}
    lines.append("== " + f)
    for k in KEYS:
        if k in d:
            lines.append("   %-86s %s %s" % (k, d[k][0], d[k][1]))
open(os.path.join(P, "kernels_r01_summary.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(out))
print("\n".join(lines))
