"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share.
usage: python tools/launch_summary.py launches.csv [first_id last_id]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else None
hi = int(sys.argv[3]) if len(sys.argv) > 3 else None
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    i = int(r["ID"])
    if (lo is not None and i < lo) or (hi is not None and i > hi):
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^.*::", "", name)
    rows.append((i, name, us, r.get("Grid Size", ""), r.get("Block Size", "")))
agg = defaultdict(lambda: [0, 0.0])
for i, name, us, g, b in rows:
    agg[name][0] += 1
    agg[name][1] += us
total = sum(v[1] for v in agg.values())
print("launches %d  total %.1f us" % (len(rows), total))
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-44s n=%4d  %10.1f us  %5.1f%%  avg %8.1f us" % (name[:44], n, us, 100 * us / total, us / n))
if "--list" in sys.argv:
    for r in rows:
        print(r)
