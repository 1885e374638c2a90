"""Aggregate the warp-stall samples of an ncu report per CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex launch_skip [top]"""
import csv
import io
import subprocess
import sys

rep, kern, skip = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
files = {}
cur = None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = files.setdefault(r[1], {})
        continue
    if r[0] == "Line No":
        hdr = r
        ci = {}
        for i, x in enumerate(hdr):
            ci.setdefault(x, i)
        stalls = [x for x in hdr if x.startswith("stall_") and "Not Issued" not in x]
        continue
    if hdr is None or cur is None or len(r) != len(hdr):
        continue
    if r[0] != "":          # a source line: carries the aggregate of its SASS
        line = int(r[0])
        d = cur.setdefault(line, {"src": r[1], "samples": 0, "inst": 0, "stalls": {}})
        d["samples"] += int(r[ci["# Samples"]] or 0)
        d["inst"] += int(r[ci["Instructions Executed"]] or 0)
        for s in stalls:
            v = int(r[ci[s]] or 0)
            if v:
                d["stalls"][s] = d["stalls"].get(s, 0) + v
for path, lines in files.items():
    tot = sum(d["samples"] for d in lines.values())
    if tot == 0:
        continue
    print("==", path, "samples", tot)
    for line, d in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:3]
        print("%5d %6d %5.1f%% %10d  %-90s %s" % (line, d["samples"], 100.0 * d["samples"] / tot, d["inst"],
                                                   d["src"].strip()[:90], st))
