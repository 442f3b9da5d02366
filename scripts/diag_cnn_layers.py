"""Per-layer forward / gradient parity of the CausalVesselVAE layer shapes (conv + BatchNorm + activation chains in
isolation, training mode) against torch fp64 on the CPU: separates kernel bugs from whole-network conditioning."""
import os, sys
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from causal_vae_b200 import nn as N

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()

def run(name, mods_ours, mods_ref, shape, B=4):
    g = torch.Generator().manual_seed(0)
    ours = N.Sequential(*mods_ours)
    ref = torch.nn.Sequential(*mods_ref).double()
    sd = {k: (torch.randn(v.shape, generator=g) * (0.05 if v.dim() > 1 else 0.3) + (1.0 if k.endswith("1.weight") or k.endswith("2.weight") and v.dim() == 1 else 0.0)
              if v.is_floating_point() else v) for k, v in ours.state_dict().items()}
    for k in sd:
        if k.endswith("running_var"):
            sd[k] = sd[k].abs() + 0.5
    ours.load_state_dict(sd); ref.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()})
    ours = ours.cuda().train(); ref.train()
    x = torch.randn(B, *shape, generator=g)
    xo = x.cuda().requires_grad_(True); xr = x.double().requires_grad_(True)
    yo = ours(xo); yr = ref(xr)
    dy = torch.randn(yr.shape, generator=g)
    yo.backward(dy.cuda()); yr.backward(dy.double())
    out = [f"y {rel(yo, yr):.1e}", f"dx {rel(xo.grad, xr.grad):.1e}"]
    for (k, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
        if q.grad.abs().max() > 1e-6 * dy.abs().max():
            out.append(f"d{k} {rel(p.grad, q.grad):.1e}")
    print(f"{name:34s}", "  ".join(out), flush=True)

T = torch.nn
enc = [(1, 32, 768, 1280), (32, 64, 384, 640), (64, 128, 192, 320), (128, 256, 96, 160), (256, 512, 48, 80), (512, 512, 24, 40), (512, 512, 12, 20)]
for a, b, h, w in enc:
    run(f"enc conv4x4s2 {a}->{b} @{h}x{w}", [N.Conv2d(a, b, 4, 2, 1), N.BatchNorm2d(b), N.LeakyReLU(0.2)],
        [T.Conv2d(a, b, 4, 2, 1), T.BatchNorm2d(b), T.LeakyReLU(0.2)], (a, h, w))
    run(f"  conv only", [N.Conv2d(a, b, 4, 2, 1)], [T.Conv2d(a, b, 4, 2, 1)], (a, h, w))
dec = [(512, 512, 6, 10), (512, 512, 12, 20), (512, 256, 24, 40), (256, 128, 48, 80), (128, 64, 96, 160), (64, 32, 192, 320)]
for a, b, h, w in dec:
    run(f"dec up+conv3x3 {a}->{b} @{h}x{w}", [N.Upsample(scale_factor=2, mode="nearest"), N.Conv2d(a, b, 3, 1, 1), N.BatchNorm2d(b), N.ReLU()],
        [T.Upsample(scale_factor=2, mode="nearest"), T.Conv2d(a, b, 3, 1, 1), T.BatchNorm2d(b), T.ReLU()], (a, h, w))
run("dec head up+conv3x3 32->1 sigmoid", [N.Upsample(scale_factor=2, mode="nearest"), N.Conv2d(32, 1, 3, 1, 1), N.Sigmoid()],
    [T.Upsample(scale_factor=2, mode="nearest"), T.Conv2d(32, 1, 3, 1, 1), T.Sigmoid()], (32, 384, 640))
run("enc_fc", [N.Linear(30751, 1024), N.BatchNorm1d(1024), N.LeakyReLU(0.2), N.Linear(1024, 256)],
    [T.Linear(30751, 1024), T.BatchNorm1d(1024), T.LeakyReLU(0.2), T.Linear(1024, 256)], (30751,), B=16)
run("dec_fc", [N.Linear(140, 1024), N.BatchNorm1d(1024), N.LeakyReLU(0.2), N.Linear(1024, 30720), N.ReLU()],
    [T.Linear(140, 1024), T.BatchNorm1d(1024), T.LeakyReLU(0.2), T.Linear(1024, 30720), T.ReLU()], (140,), B=16)
