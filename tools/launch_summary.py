"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_summary.py file.csv [skip_first_n]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
seq = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"[<(].*", "", row["Kernel Name"]).replace("void ", "").replace("bp::", "")
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000 if unit in ("ns", "nsecond") else (v * 1000 if unit in ("ms", "msecond") else v)
    seq.append((name, v))
seq = seq[skip:]
tot, cnt = collections.Counter(), collections.Counter()
for n, v in seq:
    tot[n] += v
    cnt[n] += 1
all_us = sum(tot.values())
print(f"{len(seq)} launches, {all_us / 1e3:.3f} ms of kernel time")
for n, v in tot.most_common(40):
    print(f"{n:36s} {cnt[n]:6d} launches {v:12.1f} us  avg {v / cnt[n]:9.1f} us  {100 * v / all_us:5.1f} %")
