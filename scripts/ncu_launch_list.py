"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (gpurun_out/launches.csv) by kernel.
    python scripts/ncu_launch_list.py gpurun_out/launches.csv "header text" > profiles/rN_ncu_launch_list_....txt
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
n = 0
for r in rd:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[unit]
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
    n += 1
tot = sum(a[1] for a in agg.values())
for h in sys.argv[2:]:
    print("# " + h)
print(f"# total {tot:.1f} ms over {n} launches")
for name, (k, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:10.3f} ms {100 * ms / tot:6.2f} %  n={k:5d}  {name}")
