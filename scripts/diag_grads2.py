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
xg, mg, tg, eg = x.cuda(), m.cuda(), t.cuda(), eps.cuda()
mu, lv, z = model.encode(xg, mg, tg, eg); z.retain_grad()
m_mu, m_lv = model.morph_head(tg)
zv = model.dec_adapter(F.cat_pad([mg, z])); zv.retain_grad()
di = model.backbone.decoder_input(zv); di.retain_grad()
rec = model.backbone.decoder(di.view(-1, 256, H // 32, W // 32)); rec.retain_grad()
r, k, mo, sp = train.loss_function(rec, xg, m_mu, mg, mu, lv, m_mu, m_lv)
(r + 0.5 * k + mo + 0.3 * sp).backward()
mine = dict(z=z.grad, zv=zv.grad, di=di.grad, rec=rec.grad)
mine_f = dict(z=z, zv=zv, di=di, rec=rec)
res = {}
for dt in (torch.float64, torch.float32):
    P = {k: (v.to(dt).clone() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    Wt = O.trainable(P)
    for v in Wt.values(): v.requires_grad_(True)
    X, M, T, E = x.to(dt), m.to(dt), t.to(dt), eps.to(dt)
    mu_, lv_ = O.vessel_encode(P, X, M, T, True)
    z_ = O.reparameterize(mu_, lv_, E); z_.retain_grad()
    mm, ml = O.vessel_morph_head(P, T)
    h = torch.nn.functional.leaky_relu(O._bn(P, "dec_adapter.1", O._lin(P, "dec_adapter.0", torch.cat([M, z_], 1)), True), 0.2)
    zv_ = O._lin(P, "dec_adapter.3", h); zv_.retain_grad()
    di_ = O._lin(P, "backbone.decoder_input", zv_); di_.retain_grad()
    # decoder from di
    class Q(dict):
        pass
    hh = di_.view(B, 256, H // 32, W // 32)
    i = 0
    for s in range(5):
        hh = O._convT(P, f"backbone.decoder.{i}", hh, 2, 1, 1)
        hh = torch.nn.functional.leaky_relu(O._bn(P, f"backbone.decoder.{i+1}", hh, True), 0.01); i += 3
        if s < 3:
            hh = O._resblock(P, f"backbone.decoder.{i}", hh, True); i += 1
    rec_ = O._conv(P, f"backbone.decoder.{i}", hh, 1, 1); rec_.retain_grad()
    a, b, c, d = O.vessel_loss(rec_, X, mm, M, mu_, lv_, mm, ml)
    (a + 0.5 * b + c + 0.3 * d).backward()
    res[dt] = (dict(z=z_.grad, zv=zv_.grad, di=di_.grad, rec=rec_.grad), dict(z=z_, zv=zv_, di=di_, rec=rec_))
for kx in ("rec", "di", "zv", "z"):
    g64, f64 = res[torch.float64][0][kx], res[torch.float64][1][kx]
    print(kx, "fwd err", rel(mine_f[kx], f64), "noise", rel(res[torch.float32][1][kx], f64),
          "| grad err", rel(mine[kx], g64), "noise", rel(res[torch.float32][0][kx], g64))
