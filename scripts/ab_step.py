"""Time the graph-replayed vessel step (B = 64, 256x256) under the CURRENT environment switches: one line
'<label> <ms per step>' (median and min of 5 x 20 replays, CUDA events).  Meant for A/B runs inside one gpurun call:
    CVAE_X=0 python scripts/ab_step.py off; CVAE_X=1 python scripts/ab_step.py on"""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cvae_oracle as O
from causal_vae_b200.vessel import models, train
label = sys.argv[1] if len(sys.argv) > 1 else "step"
B = int(os.environ.get("AB_BATCH", "64"))
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
ms = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        tr.replay()
    e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1) / 20)
print(f"{label}: median {statistics.median(ms):.3f} ms  min {min(ms):.3f} ms  loss {float(tr.static_losses[0]):.6g}")
