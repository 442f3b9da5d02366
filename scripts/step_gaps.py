"""Where does the wall time of the graph-replayed vessel step go?  Replays the captured step with the weight-gradient
side stream on and off; together with the kernel-duration sum of an ncu launch list of the same step
(scripts/profile_step.py) this separates kernel time, overlap and launch gaps:
    serial wall - sum(kernel durations) = launch gaps;   serial wall - overlapped wall = what the side stream hides."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cvae_oracle as O
from causal_vae_b200 import _lib as L
from causal_vae_b200.vessel import models, train
B = 64
models.CONFIG["IMG_HEIGHT"] = models.CONFIG["IMG_WIDTH"] = 256
x, m, t, eps = (a.cuda() for a in O.vessel_inputs(B, 256, 256, seed=0))
for overlap in (True, False):
    torch.manual_seed(0)
    model = models.CausalViTVAE().cuda()
    tr = train.VesselTrainer(model, lr=1e-4, overlap_wgrad=overlap)
    n0 = L.launch_count
    tr.capture(B, 256, 256, warmup=2)
    calls = (L.launch_count - n0) // 3
    tr.load_batch(x, m, t, eps)
    for _ in range(5):
        tr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        tr.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"side stream {'on ' if overlap else 'off'}: {e0.elapsed_time(e1) / 20:.3f} ms per step, {calls} library calls per step", flush=True)
    del tr, model
    torch.cuda.empty_cache()
