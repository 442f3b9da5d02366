"""Whole-step ledger from an ncu --csv launch list carrying gpu__time_duration.sum, dram__bytes_read.sum and
dram__bytes_write.sum per launch: per-kernel-name time, DRAM bytes, and the step total against the algorithmic
bytes of the counting rule (scripts/algorithmic_bytes.py).  usage: summarize_traffic.py launches.csv [--top N] [--csv out.csv]"""
import csv, sys, collections, re
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
out_csv = sys.argv[sys.argv.index("--csv") + 1] if "--csv" in sys.argv else None
per = collections.OrderedDict()
for r in csv.DictReader(lines):
    i = r["ID"]
    e = per.setdefault(i, {"name": re.sub(r"\(.*", "", r["Kernel Name"]), "grid": r["Grid Size"], "block": r["Block Size"]})
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]; m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        e["us"] = v / 1000.0 if u in ("nsecond", "ns") else v * 1000.0 if u in ("msecond", "ms") else v
    else:
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        e["rd" if "read" in m else "wr"] = v * mult
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for e in per.values():
    a = agg[e["name"]]; a[0] += 1; a[1] += e.get("us", 0); a[2] += e.get("rd", 0); a[3] += e.get("wr", 0)
T = sum(a[1] for a in agg.values()); R = sum(a[2] for a in agg.values()); W = sum(a[3] for a in agg.values())
print(f"total {T/1000:.3f} ms over {len(per)} launches; DRAM read {R/1e9:.3f} GB + write {W/1e9:.3f} GB = {(R+W)/1e9:.3f} GB")
print(f"{'ms':>9} {'%':>6} {'n':>5} {'read MB':>9} {'write MB':>9} {'GB/s':>7}  kernel")
for name, (n, us, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us/1000:9.3f} {100*us/T:5.1f}% {n:5d} {rd/1e6:9.1f} {wr/1e6:9.1f} {(rd+wr)/max(us,1e-9)/1e3:7.0f}  {name}")
if top:
    print(f"--- {top} longest launches")
    for e in sorted(per.values(), key=lambda e: -e.get("us", 0))[:top]:
        print(f"{e.get('us',0):9.1f} us  rd {e.get('rd',0)/1e6:8.1f} MB wr {e.get('wr',0)/1e6:8.1f} MB  {(e.get('rd',0)+e.get('wr',0))/max(e.get('us',1),1e-9)/1e3:6.0f} GB/s grid {e['grid']:<16} {e['name']}")
if out_csv:
    with open(out_csv, "w") as f:
        f.write("id,kernel,grid,block,us,dram_read_bytes,dram_write_bytes\n")
        for i, e in per.items():
            f.write(f"{i},\"{e['name']}\",\"{e['grid']}\",\"{e['block']}\",{e.get('us',0):.2f},{e.get('rd',0):.0f},{e.get('wr',0):.0f}\n")
