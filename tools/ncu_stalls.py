"""Top warp-stall reasons + memory numbers per profiled launch of an ncu --set full report.
usage: python tools/ncu_stalls.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
for r in rows[2:]:
    name = r[ci["Kernel Name"]].split("(")[0].split("::")[-1][:40]

    def g(k):
        try:
            return float(r[ci[k]].replace(",", ""))
        except Exception:
            return 0.0
    tops = sorted(((g(h), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:4]
    print("%-40s %8.1f us  dram %5.1f%%  l2hit %5.1f%%  occ %5.1f%%  ipc %.2f | %s" % (
        name, g("gpu__time_duration.sum") / (1e3 if units[ci["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1),
        g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g("lts__t_sector_hit_rate.pct"),
        g("sm__warps_active.avg.pct_of_peak_sustained_active"), g("sm__inst_executed.avg.per_cycle_elapsed"),
        ", ".join("%s %.1f" % (n, v) for v, n in tops)))
