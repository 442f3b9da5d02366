"""Run-to-run reproducibility of one eager training step (same inputs, two fresh models)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for tc in ("1", "0"):
        subprocess.run([sys.executable, __file__, "go"], env=dict(os.environ, CVAE_TC=tc))
    sys.exit(0)
import torch
from oracle import cvae_oracle as O
from causal_vae_b200.vessel import models, train
H = W = 64; B = 4
models.CONFIG["IMG_HEIGHT"], models.CONFIG["IMG_WIDTH"] = H, W
x, m, t, eps = (a.cuda() for a in O.vessel_inputs(B, H, W, seed=0))
sd = O.fill_state_dict(O.vessel_shapes(H, W), seed=0)
runs = []
for r in range(3):
    model = models.CausalViTVAE(); model.load_state_dict(sd); model = model.cuda()
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout): mod.p = 0.0
        if hasattr(mod, "in_proj_weight"): mod.dropout = 0.0
    tr = train.VesselTrainer(model, lr=1e-3)
    tr.model.train()
    losses = tr._fwd_bwd(x, m, t, eps)
    torch.cuda.synchronize()
    runs.append((float(losses[0]), {k: p.grad.detach().clone() for k, p in model.named_parameters()}))
print(f"TC={os.environ.get('CVAE_TC')} losses {[r[0] for r in runs]}")
worst = []
for k in runs[0][1]:
    a, b = runs[0][1][k], runs[1][1][k]
    d = (a - b).abs().max().item() / max(a.abs().max().item(), 1e-30)
    worst.append((d, k))
worst.sort(reverse=True)
print("  worst run-to-run grad rel diffs:", [(f"{d:.1e}", k) for d, k in worst[:6]])
