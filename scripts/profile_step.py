"""One eager vessel training step (B=64, 256x256) between cudaProfilerStart/Stop for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
       --log-file gpurun_out/launches.csv python scripts/profile_step.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cvae_oracle as O
from causal_vae_b200.vessel import models, train
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
models.CONFIG["IMG_HEIGHT"] = models.CONFIG["IMG_WIDTH"] = 256
torch.manual_seed(0)
model = models.CausalViTVAE().cuda()
tr = train.VesselTrainer(model, lr=1e-4)
x, m, t, eps = (a.cuda() for a in O.vessel_inputs(B, 256, 256, seed=0))
for _ in range(2):
    tr.step(x, m, t, eps)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(x, m, t, eps)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
