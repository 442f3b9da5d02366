"""Time the graph-replayed step of the small configs (mnist01 | cascade | latent_translator) under the CURRENT environment
switches: one line '<label> <config> <ms per step>' each (median / min of 5 x 20 replays, CUDA events).  For A/B runs inside
one gpurun call:   CVAE_SMALL_SCOPE=0 python scripts/ab_small.py off; python scripts/ab_small.py on"""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
label = sys.argv[1] if len(sys.argv) > 1 else "step"
names = sys.argv[2:] or ["mnist01", "cascade", "latent_translator"]
for name in names:
    gs, pin, loss_of, tr = bench._small_trainer(torch, name, False)
    gs.load(**{k: v.cuda() for k, v in pin.items()})
    for _ in range(5):
        out = gs.replay()
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            out = gs.replay()
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1) / 20)
    print(f"{label} {name}: median {statistics.median(ms):.3f} ms  min {min(ms):.3f} ms  loss {float(loss_of(out)):.6g}", flush=True)
    del gs, tr
