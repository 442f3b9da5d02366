"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel name."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("nsecond", "ns") else v * 1000.0 if unit in ("msecond", "ms") else v
    agg[name][0] += 1; agg[name][1] += us; total += us
print(f"total {total/1000:.3f} ms over {sum(a[0] for a in agg.values())} launches")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us/1000:9.3f} ms {100*us/total:5.1f}%  x{n:4d}  {name}")
