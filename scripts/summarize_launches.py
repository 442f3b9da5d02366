"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel name; with --top N also
list the N longest individual launches with their grids."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
rd = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
total = 0.0
indiv = []
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("nsecond", "ns") else v * 1000.0 if unit in ("msecond", "ms") else v
    agg[name][0] += 1; agg[name][1] += us; total += us
    indiv.append((us, name, r["Grid Size"], r["Block Size"], r["ID"]))
print(f"total {total/1000:.3f} ms over {sum(a[0] for a in agg.values())} launches")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us/1000:9.3f} ms {100*us/total:5.1f}%  x{n:4d}  {name}")
if top:
    print(f"--- {top} longest launches")
    for us, name, grid, block, i in sorted(indiv, reverse=True)[:top]:
        print(f"{us:9.1f} us  id {i:>5}  grid {grid:<18} block {block:<14} {name}")
