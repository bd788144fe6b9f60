"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name:
  python tools/ncu_kernel_summary.py gpurun_out/effdet_launches.csv > profiles/..._summary.txt"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
    rows.append((name, v, r.get("Grid Size", ""), r.get("Block Size", "")))
tot = sum(r[1] for r in rows)
agg = collections.OrderedDict()
for n, v, _, _ in rows:
    a = agg.setdefault(n, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
print("# %d launches, %.1f us total (ncu durations: serialised, cold-cache; shares are what matters)" % (len(rows), tot))
print("kernel,launches,total_us,share,mean_us,max_us")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.1f,%.1f%%,%.2f,%.2f" % (n, a[0], a[1], 100 * a[1] / tot, a[1] / a[0], a[2]))
