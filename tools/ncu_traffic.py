"""Per-kernel DRAM traffic from an ncu CSV taken with
   --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
Writes a JSON summary (bytes per launch, averaged per kernel family) used by bench.py for roofline.traffic."""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ki, mi, vi, ui, ii = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
per = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    d = per.setdefault(r[ii], {"name": r[ki].split("(")[0].replace("void ", "").replace("s3b::", "")})
    d[r[mi]] = v * mult
fam = collections.OrderedDict()
for d in per.values():
    name = d["name"].split("<")[0]
    f = fam.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0})
    f["launches"] += 1
    f["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    f["time_us"] += d.get("gpu__time_duration.sum", 0.0)
out = {k: {"launches": v["launches"], "dram_bytes_per_launch": v["dram_bytes"] / v["launches"],
           "time_us_per_launch": v["time_us"] / v["launches"]} for k, v in fam.items()}
json.dump(out, open(sys.argv[2], "w"), indent=1)
for k, v in out.items():
    print(f"{k:28s} launches {v['launches']:4d}  dram/launch {v['dram_bytes_per_launch'] / 1e6:9.2f} MB  "
          f"time/launch {v['time_us_per_launch']:9.1f} us")
