"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

  python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"<.*", "", name)
    m = re.search(r"mmh::(\w+)", r["Kernel Name"])
    f = re.search(r"mmh::(\w+F)\b", r["Kernel Name"])
    if f:
        name = "mmh::" + (m.group(1) if m else "") + "<" + f.group(1) + ">"
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    agg[name][0] += 1
    agg[name][1] += ns
    total += ns
print("%-70s %8s %12s %7s" % ("kernel", "launches", "total_ms", "share"))
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %8d %12.3f %6.1f%%" % (name[:70], n, ns / 1e6, 100.0 * ns / total))
print("%-70s %8d %12.3f" % ("TOTAL", sum(v[0] for v in agg.values()), total / 1e6))
