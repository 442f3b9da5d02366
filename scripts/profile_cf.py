"""One eager counterfactual sweep (32 sources x 12 concepts, eval mode) under the CUDA profiler range:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
       --log-file gpurun_out/launches_cf.csv python scripts/profile_cf.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from causal_vae_b200 import counterfactual as CF
from causal_vae_b200.vessel import models
models.CONFIG["IMG_HEIGHT"] = models.CONFIG["IMG_WIDTH"] = 256
torch.manual_seed(0)
model = models.CausalViTVAE().cuda().eval()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = torch.randn(S, models.CONFIG["M_DIM"], device="cuda")
z = torch.randn(S, models.CONFIG["Z_DIM"], device="cuda")
with torch.no_grad():
    for _ in range(2):
        CF.counterfactual_sweep(model, m, z, delta=5.0)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    CF.counterfactual_sweep(model, m, z, delta=5.0)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok")
