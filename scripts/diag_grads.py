"""Diagnostic: per-tensor gradient error of the native vessel step vs the fp64 oracle, next to the
oracle's own fp32 noise.  Usage: python scripts/diag_grads.py H W B"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cvae_oracle as O
from tests.test_vessel_gpu import build, rel
from causal_vae_b200.vessel import train
H, W, B = map(int, sys.argv[1:4])
model, sd = build(H, W)
x, m, t, eps = O.vessel_inputs(B, H, W, seed=0)
tr = train.VesselTrainer(model, lr=1e-4)
tr.model.train()
tr._fwd_bwd(x.cuda(), m.cuda(), t.cuda(), eps.cuda())
grads = {k: p.grad.detach().clone().cpu() for k, p in model.named_parameters()}
P64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
_, g64, _ = O.vessel_train_step(P64, {}, 1, x.double(), m.double(), t.double(), eps.double())
_, g32, _ = O.vessel_train_step({k: v.clone() for k, v in sd.items()}, {}, 1, x, m, t, eps)
rows = []
for k, g in g64.items():
    rows.append((rel(grads[k], g), rel(g32[k], g), float(g.abs().max()), k))
rows.sort(reverse=True)
for e, n, mx, k in rows[:60]:
    print(f"{e:9.2e} noise {n:9.2e} max|g| {mx:9.2e}  {k}")
k = "backbone.transformer.4.attn.in_proj_bias"
for i, nm in enumerate("qkv"):
    sl = slice(256 * i, 256 * (i + 1))
    print(nm, "err", (grads[k][sl].double() - g64[k][sl]).abs().max().item(), "noise", (g32[k][sl].double() - g64[k][sl]).abs().max().item(), "max", g64[k][sl].abs().max().item())
