"""Timeline of ONE graph-replayed vessel step (B = 64) from torch.profiler (CUPTI): per stream the kernel time and the idle
gaps inside the step, the longest gaps, and the kernels of each stream in start order (written to gpurun_out/).
    python scripts/trace_step.py [out_prefix]"""
import json, os, sys, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cvae_oracle as O
from causal_vae_b200.vessel import models, train
from torch.profiler import profile, ProfilerActivity
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_step"
B = 64
models.CONFIG["IMG_HEIGHT"] = models.CONFIG["IMG_WIDTH"] = 256
torch.manual_seed(0)
model = models.CausalViTVAE().cuda()
tr = train.VesselTrainer(model, lr=1e-4)
x, m, t, eps = (a.cuda() for a in O.vessel_inputs(B, 256, 256, seed=0))
tr.capture(B, 256, 256, warmup=2)
tr.load_batch(x, m, t, eps)
for _ in range(5):
    tr.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.replay()
    torch.cuda.synchronize()
prof.export_chrome_trace(out + ".json")
ev = [e for e in json.load(open(out + ".json"))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
# split into the three replays by large gaps in start order: take the middle third by count
n = len(ev) // 3
step = ev[n:2 * n]
t0 = min(e["ts"] for e in step); t1 = max(e["ts"] + e["dur"] for e in step)
print(f"{len(step)} kernels in the step, span {(t1 - t0) / 1e3:.3f} ms")
by = collections.defaultdict(list)
for e in step:
    by[e["args"].get("stream")].append(e)
for s, L in sorted(by.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    busy = sum(e["dur"] for e in L)
    gaps = []
    for a, b in zip(L, L[1:]):
        g = b["ts"] - (a["ts"] + a["dur"])
        if g > 0:
            gaps.append((g, a["name"][:50], b["name"][:50]))
    print(f"stream {s}: {len(L)} kernels, busy {busy / 1e3:.3f} ms, first {(L[0]['ts'] - t0) / 1e3:.3f} ms, last end "
          f"{(L[-1]['ts'] + L[-1]['dur'] - t0) / 1e3:.3f} ms, sum of gaps {sum(g for g, _, _ in gaps) / 1e3:.3f} ms "
          f"({sum(1 for g, _, _ in gaps if g > 5)} gaps > 5 us)")
    for g, a, b in sorted(gaps, reverse=True)[:8]:
        print(f"     gap {g:8.1f} us after {a}  before {b}")
with open(out + "_kernels.txt", "w") as f:
    for e in step:
        f.write(f"{(e['ts'] - t0):10.1f} {e['dur']:8.1f} s{e['args'].get('stream')} {e['name'][:90]}\n")
agg = collections.defaultdict(lambda: [0, 0.0])
for e in step:
    k = e["name"].split("(")[0][:60]
    agg[k][0] += 1; agg[k][1] += e["dur"]
print("in-graph kernel time by name (warm caches, concurrent streams):")
for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"  {d / 1e3:7.3f} ms x{c:4d}  {k}")
