"""Summarise the ncu artefacts of one round under profiles/: the launch lists (gpu__time_duration.sum) into
profiles/launches*_<round>_summary.txt and the key metrics of the `--set full` captures into profiles/kernels_<round>_summary.txt.
usage: python tools/ncu_summary.py [r02]"""
import collections
import csv
import glob
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
COMMANDS = {
    "launches_%s.csv" % R: "python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-proofs --no-sweep",
    "launches_verify_batch_%s.csv" % R: "VB_MODES=0 BPH_VB_DRIVERS=1 python tools/verify_batch_bench.py 4096 (one range_prove_batch call to make the proofs, "
                                        "then bph_range_verify_batch x 3)",
    "launches_prove_batch_%s.csv" % R: "python tools/prove_batch_bench.py 2048 1 64 bls 2 0 (bph_range_prove_batch x 2, device transcripts, one verify)",
    "launches_config3_%s.csv" % R: "python tools/proof_trace.py 256 bn 1 (config 3: two bph_range_prove + one bph_range_verify at n = 2^14 on BN254, "
                                   "window tables (k_fb_*: one-off set-up), recorded circuit, hybrid IPP)",
}
MSM = ("k_digits", "k_scan", "k_scatter", "k_chunk", "k_giant", "k_merge", "k_reduce")


def launch_summary(fname, cmd):
    path = os.path.join(P, fname)
    if not os.path.exists(path):
        return
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"[<(].*", "", r[ki]).replace("void ", "").replace("bp::", "")
        v = float(r[vi].replace(",", ""))
        ms = {"ns": v / 1e6, "nsecond": v / 1e6, "us": v / 1e3, "usecond": v / 1e3, "ms": v, "msecond": v}.get(r[ui], v * 1e3)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    tm = sum(a[1] for k, a in agg.items() if any(x in k for x in MSM))
    out = ["# ncu launch list summary (gpu__time_duration.sum, --clock-control none): " + cmd,
           "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes (set-up kernels -- k_mapit, k_fb_*, k_group_op -- included)",
           "%-40s %8s %10s %9s %7s %9s" % ("kernel", "launches", "total ms", "avg ms", "share", "msm share")]
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("%-40s %8d %10.3f %9.4f %6.1f%% %8s" % (k[:40], c, t, t / c, 100 * t / tot,
                                                          ("%.1f%%" % (100 * t / tm)) if tm and any(x in k for x in MSM) else "-"))
    open(os.path.join(P, fname.replace(".csv", "_summary.txt")), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


for f, cmd in COMMANDS.items():
    launch_summary(f, cmd)

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
lines = ["# key metrics of the ncu --set full captures of round %s (one launch each); raw pages: profiles/<kernel>_%s_raw.csv" % (R, R),
         "# MSM kernels: 2^20-term BLS12-381 MSM (bench.py); k_batch_* / k_vb_*: one 4096-proof slab of bph_range_verify_batch (n = 64);",
         "# k_batch_fixed_warp_ipp / k_pd_normalise: an IPP round of a 2048-proof slab of bph_range_prove_batch (device transcripts)"]
for path in sorted(glob.glob(os.path.join(P, "*_%s_raw.csv" % R))):
    rows = list(csv.reader(open(path)))
    d = {n: (rows[2][i], rows[1][i]) for i, n in enumerate(rows[0])}
    lines.append("== " + os.path.basename(path).replace("_%s_raw.csv" % R, ""))
    for k in KEYS:
        if k in d:
            lines.append("   %-86s %s %s" % (k, d[k][0], d[k][1]))
open(os.path.join(P, "kernels_%s_summary.txt" % R), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:6]))
