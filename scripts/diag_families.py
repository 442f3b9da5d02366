"""Per-parameter gradient errors of the cascade model against the fp64 oracle (diagnostic)."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cvae_oracle as O
from causal_vae_b200.cascade import models, train

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()

g = json.load(open("tests/golden/cascade.json")); c = g["config"]
P = O.fill_state_dict(O.cascade_shapes(8, 19), seed=c["wseed"])
gen = torch.Generator().manual_seed(c["xseed"]); B = c["B"]
x = torch.randn(B, 1, 64, 64, generator=gen); m = torch.rand(B, 8, generator=gen)
t = torch.randint(0, 19, (B,), generator=gen); eps = torch.randn(B, 64, generator=gen)
model = models.CausalBioVAE(1, 8, 19, 64); model.load_state_dict(P); model = model.cuda().train()
P64 = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in P.items()}
o = O.cascade_forward(P64, x.double(), m.double(), t, eps.double(), train=True)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
def loss_of(o, x, m, F=None):
    if which == "kld":
        return -0.5 * torch.sum(1 + o[3] - o[2].pow(2) - o[3].exp())
    if which == "rec":
        return torch.sum((o[0] - x) ** 2)
    if which == "aten_mu":
        return 0.5 * (o[2] ** 2).sum()
    if which == "aten_lv":
        return (o[3].exp() - o[3]).sum() * 0.5
    return O.cascade_loss(o[0], x, o[1], m, o[2], o[3])[0]
loss_of(o, x.double(), m.double()).backward()
outs = model(x.cuda(), m.cuda(), t.cuda(), eps.cuda())
from causal_vae_b200 import functional as F
if which == "kld":
    l = F.kld_loss(outs[2], outs[3])
elif which == "rec":
    l = F.mse_sum(outs[0], x.cuda())
elif which == "aten_mu":
    l = 0.5 * (outs[2] ** 2).sum()
elif which == "aten_lv":
    l = (outs[3].exp() - outs[3]).sum() * 0.5
else:
    l = train.loss_function(outs[0], x.cuda(), outs[1], m.cuda(), outs[2], outs[3])[0]
l.backward()
for k, p in model.named_parameters():
    g64 = P64[k].grad
    if g64 is None or not k.startswith(("enc_fc", "fc_")):
        continue
    print(f"{k:28s} {'none' if p.grad is None else f'{rel(p.grad, g64):.2e}'}   max|g| {float(g64.abs().max()):.3e}")
