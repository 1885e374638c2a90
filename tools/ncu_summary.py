"""One line per profiled launch of an ncu --set full report: duration, DRAM bytes / throughput, tensor-pipe activity.
usage: python tools/ncu_summary.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ci = {h: i for i, h in enumerate(hdr)}


def g(r, k, d=0.0):
    try:
        return float(r[ci[k]].replace(",", ""))
    except Exception:
        return d


def to_bytes(r, k):
    u = units[ci[k]]
    return g(r, k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(r, k):
    u = units[ci[k]]
    return g(r, k) * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1)


print("%-34s %-14s %9s %9s %9s %7s %7s %7s %5s %6s" % ("kernel", "grid", "us", "rd MB", "wr MB", "GB/s", "dram%", "tens%", "regs", "occ%"))
for r in rows[2:]:
    name = r[ci["Kernel Name"]].split("(")[0].split("::")[-1][:34]
    us = to_us(r, "gpu__time_duration.sum")
    rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
    print("%-34s %-14s %9.1f %9.1f %9.1f %7.0f %7.1f %7.1f %5d %6.1f" % (
        name, r[ci["Grid Size"]].replace(" ", ""), us, rd / 1e6, wr / 1e6, (rd + wr) / us / 1e3,
        g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        int(g(r, "launch__registers_per_thread")), g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")))
