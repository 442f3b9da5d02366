import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cvae_oracle as O
from tests.test_vessel_gpu import build, rel
from causal_vae_b200.vessel import train
from causal_vae_b200 import functional as F
H, W, B = map(int, sys.argv[1:4])
model, sd = build(H, W)
x, m, t, eps = O.vessel_inputs(B, H, W, seed=0)
model.train()
lr = torch.nn.functional.leaky_relu
def oracle(dt):
    P = {k: (v.to(dt).clone() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    X, M, T, E = x.to(dt), m.to(dt), t.to(dt), eps.to(dt)
    with torch.no_grad():
        mu_, lv_ = O.vessel_encode(P, X, M, T, True)
        z_ = O.reparameterize(mu_, lv_, E)
        h = lr(O._bn(P, "dec_adapter.1", O._lin(P, "dec_adapter.0", torch.cat([M, z_], 1)), True), 0.2)
        di_ = O._lin(P, "backbone.decoder_input", O._lin(P, "dec_adapter.3", h))
    acts = {}
    hh = di_.view(B, 256, H // 32, W // 32).clone().requires_grad_(True); acts[0] = hh
    i = 0
    for s in range(5):
        hh = O._convT(P, f"backbone.decoder.{i}", hh, 2, 1, 1)
        hh = lr(O._bn(P, f"backbone.decoder.{i+1}", hh, True), 0.01); i += 3
        hh.retain_grad(); acts[i] = hh
        if s < 3:
            hh = O._resblock(P, f"backbone.decoder.{i}", hh, True); i += 1
            hh.retain_grad(); acts[i] = hh
    rec_ = O._conv(P, f"backbone.decoder.{i}", hh, 1, 1)
    a, b, c, d = O.vessel_loss(rec_, X, None, M, torch.zeros(1, dtype=dt), torch.zeros(1, dtype=dt), M, M * 0)
    (a + 0.3 * d).backward()
    return acts
a64, a32 = oracle(torch.float64), oracle(torch.float32)
xg = x.cuda()
for j in sorted(a64):
    model.zero_grad()
    inp = a32[j].detach().cuda().requires_grad_(True)
    rec = model.backbone.decoder[j:](inp)
    r, sp = F.vessel_recon_loss(rec, xg)
    (r + 0.3 * sp).backward()
    print("decoder[%d:]" % j, tuple(inp.shape), "grad-in err", rel(inp.grad, a64[j].grad), "noise", rel(a32[j].grad, a64[j].grad))
# ---- zoom into decoder[3] (ResBlock 128 @ 4x4) -----------------------------------------------
j = 3
P = {k: v.double().clone() if v.is_floating_point() else v.clone() for k, v in sd.items()}
a_in = a64[j].detach()
y1 = O._conv(P, "backbone.decoder.3.conv.0", a_in, 1, 1)
print("y1 per-channel |mean|/std max:", (y1.mean((0,2,3)).abs() / y1.std((0,2,3))).max().item(), "min std", y1.std((0,2,3)).min().item())
h1 = lr(O._bn(P, "backbone.decoder.3.conv.1", y1, True), 0.2)
y2 = O._conv(P, "backbone.decoder.3.conv.3", h1, 1, 1)
print("y2 per-channel |mean|/std max:", (y2.mean((0,2,3)).abs() / y2.std((0,2,3))).max().item(), "min std", y2.std((0,2,3)).min().item())
model.zero_grad()
inp = a32[j].detach().cuda().requires_grad_(True)
blk = model.backbone.decoder[3:4]
out = blk(inp)
gup = a64[4].grad.float()
out.backward(gup.cuda())
ref_in = a64[3].grad - 0  # includes everything downstream via same upstream
# oracle grad of the block alone with the same upstream
ai = a64[3].detach().clone().requires_grad_(True)
P2 = {k: v.double().clone() if v.is_floating_point() else v.clone() for k, v in sd.items()}
for v in O.trainable(P2).values(): v.requires_grad_(True)
o2 = O._resblock(P2, "backbone.decoder.3", ai, True); o2.backward(a64[4].grad)
e = (inp.grad.double().cpu() - ai.grad).abs()
print("block-alone grad-in err", (e.max() / ai.grad.abs().max()).item(), "fwd err", rel(out, o2))
per_c = e.amax((0, 2, 3)); print("worst channels", per_c.topk(5))
for k in ("conv.0.weight", "conv.1.weight", "conv.1.bias", "conv.3.weight", "conv.4.weight", "conv.4.bias"):
    print(k, rel(dict(blk.named_parameters())["3." + k].grad, P2["backbone.decoder.3." + k].grad))
print("upstream grad: max", gup.abs().max().item(), "per-channel mean/std ratio max", (gup.mean((0,2,3)).abs() / gup.std((0,2,3))).max().item())
# ---- [ResBlock, ConvT+BN+LReLU] with real data, upstream = oracle grad at acts[7] ---------------
model.zero_grad()
inp = a32[3].detach().cuda().requires_grad_(True)
out = model.backbone.decoder[3:7](inp)
out.backward(a64[7].grad.float().cuda())
print("[res128, convT] real data: grad-in err", rel(inp.grad, a64[3].grad), " fwd err", rel(out, a64[7]))
for k in ("3.conv.4.weight", "3.conv.4.bias", "3.conv.3.weight", "4.weight", "5.weight"):
    pass
# random-data version of the same two-element chain
from causal_vae_b200 import nn
for C, Co, Hh in ((128, 64, 4), (64, 32, 4), (128, 64, 8)):
    seq = nn.Sequential(nn.ResBlock(C), nn.ConvTranspose2d(C, Co, 3, 2, 1, 1), nn.BatchNorm2d(Co), nn.LeakyReLU())
    sdd = O.fill_state_dict({k: tuple(v.shape) for k, v in seq.state_dict().items()}, seed=1)
    seq.load_state_dict(sdd); seq = seq.cuda().train()
    g = torch.Generator().manual_seed(C)
    xx = torch.randn(4, C, Hh, Hh, generator=g); gy = torch.randn(4, Co, 2 * Hh, 2 * Hh, generator=g) + 0.3
    xg2 = xx.cuda().requires_grad_(True); yy = seq(xg2); yy.backward(gy.cuda())
    Pq = {k: v.double().clone() if v.is_floating_point() else v.clone() for k, v in sdd.items()}
    xr = xx.double().requires_grad_(True)
    hh = O._resblock(Pq, "0", xr, True)
    yr = lr(O._bn(Pq, "2", O._convT(Pq, "1", hh, 2, 1, 1), True), 0.01); yr.backward(gy.double())
    print("random [res%d, convT->%d] @%d: fwd" % (C, Co, Hh), rel(yy, yr), "dx", rel(xg2.grad, xr.grad))
