"""Summarise an .ncu-rep: per launch key metrics (--raw) and per-source-line instruction / stall
samples (--lines N) using the ncu CLI's csv pages."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 0
launch = int(sys.argv[sys.argv.index("--launch") + 1]) if "--launch" in sys.argv else 0
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__grid_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_bytes.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {k: i for i, k in enumerate(hdr)}
for n, r in enumerate(rows[2:]):
    print(f"--- launch {n}: {r[idx['Kernel Name']][:60]}  grid {r[idx.get('launch__grid_size', 0)]}")
    for k in KEYS:
        if k in idx:
            print(f"   {k:70s} {r[idx[k]]:>16s} {units[idx[k]]}")
if nlines:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--launch-skip", str(launch), "--launch-count", "1"], capture_output=True, text=True).stdout
    cur, out = None, []
    for r in csv.reader(io.StringIO(src)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if len(r) > 7 and r[0].isdigit():
            try:
                out.append((int(r[7]), int(r[6]), cur, int(r[0]), r[1].strip()[:100]))
            except ValueError:
                pass
    tot, ts = sum(o[0] for o in out), sum(o[1] for o in out)
    print(f"== launch {launch}: {tot} warp instructions, {ts} stall samples; top lines by samples")
    for n, s, f, l, t in sorted(out, key=lambda o: -o[1])[:nlines]:
        print(f"  smp {s:5d} {100*s/max(ts,1):5.1f}%  inst {n:9d} {100*n/max(tot,1):5.1f}%  {f}:{l}  {t}")
