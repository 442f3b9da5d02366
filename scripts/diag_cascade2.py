import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cvae_oracle as O
from causal_vae_b200.cascade import models
from causal_vae_b200 import functional as F
import torch.nn.functional as TF
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
g = json.load(open("tests/golden/cascade.json")); c = g["config"]
P = O.fill_state_dict(O.cascade_shapes(8, 19), seed=c["wseed"])
gen = torch.Generator().manual_seed(c["xseed"]); B = c["B"]
x = torch.randn(B, 1, 64, 64, generator=gen); m = torch.rand(B, 8, generator=gen)
t = torch.randint(0, 19, (B,), generator=gen); eps = torch.randn(B, 64, generator=gen)
mode = sys.argv[1]
model = models.CausalBioVAE(1, 8, 19, 64); model.load_state_dict(P); model = model.cuda().train()
# ---- oracle with retained intermediates
P64 = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in P.items()}
xd, md = x.double(), m.double()
t1 = TF.one_hot(t, 19).double()
h = xd
for i in range(4):
    h = TF.relu(TF.conv2d(h, P64[f"enc_conv.{2*i}.weight"], P64[f"enc_conv.{2*i}.bias"], 2, 1))
feat_r = h.flatten(1); feat_r.retain_grad()
cat_r = torch.cat([feat_r, md, t1], 1); cat_r.retain_grad()
y0_r = TF.linear(cat_r, P64["enc_fc.0.weight"], P64["enc_fc.0.bias"]); y0_r.retain_grad()
y2_r = TF.linear(TF.relu(y0_r), P64["enc_fc.2.weight"], P64["enc_fc.2.bias"]); y2_r.retain_grad()
h2_r = TF.relu(y2_r); h2_r.retain_grad()
mu_r = TF.linear(h2_r, P64["fc_mu.weight"], P64["fc_mu.bias"]); lv_r = TF.linear(h2_r, P64["fc_logvar.weight"], P64["fc_logvar.bias"])
mu_r.retain_grad(); lv_r.retain_grad()
def kl(mu, lv): return -0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp())
kl(mu_r, lv_r).backward()
# ---- model
xc, mc = x.cuda(), m.cuda()
toh = F.one_hot(t.cuda(), 19)
feat = model.enc_conv(xc); feat.retain_grad()
cat = F.cat_pad([feat, mc, toh]); cat.retain_grad()
if mode.startswith("split"):
    y0 = model.enc_fc[0](cat); y0.retain_grad()
    a0 = model.enc_fc[1](y0)
    y2 = model.enc_fc[2](a0); y2.retain_grad()
    h2 = model.enc_fc[3](y2)
else:
    h2 = model.enc_fc(cat)
h2.retain_grad()
mu, lv = model.fc_mu(h2), model.fc_logvar(h2); mu.retain_grad(); lv.retain_grad()
l = F.kld_loss(mu, lv) if mode.endswith("native") else kl(mu, lv)
l.backward()
print(mode, "loss", float(l), float(kl(mu_r, lv_r)))
print(" d mu", rel(mu.grad, mu_r.grad), " d lv", rel(lv.grad, lv_r.grad), " d h2", rel(h2.grad, h2_r.grad))
if mode.startswith("split"):
    print(" d y2", rel(y2.grad, y2_r.grad), " d y0", rel(y0.grad, y0_r.grad))
print(" d cat", rel(cat.grad[:, :4123], cat_r.grad), " d feat", rel(feat.grad, feat_r.grad))
for k in ("enc_fc.0.weight", "enc_fc.0.bias", "enc_fc.2.weight", "fc_mu.weight"):
    print(" ", k, rel(dict(model.named_parameters())[k].grad, P64[k].grad))
if mode.startswith("split"):
    a, b = y0.detach().double().cpu(), y0_r.detach()
    flip = ((a > 0) != (b > 0)).nonzero()
    print("sign flips in y0:", flip.tolist()[:10], [(float(a[i, j]), float(b[i, j])) for i, j in flip.tolist()[:10]])
    print("y0 rel err", rel(a, b), "min |y0_r|", float(b.abs().min()))
    d = (y0.grad.double().cpu() - y0_r.grad).abs()
    idx = d.flatten().topk(5).indices
    print("worst d y0 at", [(int(i) // 512, int(i) % 512, float(d.flatten()[i]), float(b.flatten()[i]), float(a.flatten()[i])) for i in idx])
    fa, fb = feat.detach().double().cpu(), feat_r.detach()
    print("feat rel err", rel(fa, fb))
