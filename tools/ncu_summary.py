"""Summarise ncu CSV output into text for profiles/.
  python tools/ncu_summary.py launches <launches.csv>          per-kernel share of the step (gpu__time_duration)
  python tools/ncu_summary.py report <file.ncu-rep> [regex]    key metrics per captured launch"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_active.avg",
    "sm__cycles_elapsed.max",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        if "Metric Name" in hdr and r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
            continue  # CSVs that also carry DRAM byte counters (tools/ncu_traffic.py reads those)
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] == "ns" else (v * 1000.0 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"launches {sum(a[0] for a in agg.values())}  total {tot:.1f} us (cold-cache, serialised: compare shares)")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{a[1] / tot * 100:6.2f}%  {a[1]:12.1f} us  {a[0]:5d} launches  {a[1] / a[0]:9.1f} us/launch  {k}")


def report(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    idx = [i for i, h in enumerate(hdr) if h in KEYS]
    for r in rows[2:]:
        if pattern and not re.search(pattern, r[ki]):
            continue
        print("----", r[ki][:90])
        for i in idx:
            print(f"  {hdr[i]} = {r[i]} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
