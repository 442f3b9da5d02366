"""ResBlock chain (the failing test shape) through TC and SIMT: error sparsity against fp64."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for tc in ("1", "0"):
        for seed in ("10", "11"):
            subprocess.run([sys.executable, __file__, seed], env=dict(os.environ, CVAE_TC=tc))
    sys.exit(0)
import torch
from oracle import cvae_oracle as O
from causal_vae_b200 import nn
from tests.test_ops_gpu import gen, rel
seed = int(sys.argv[1])
seq = nn.Sequential(nn.ConvTranspose2d(64, 32, 3, 2, 1, 1), nn.BatchNorm2d(32), nn.LeakyReLU(), nn.ResBlock(32),
                    nn.ConvTranspose2d(32, 16, 3, 2, 1, 1), nn.BatchNorm2d(16), nn.LeakyReLU(),
                    nn.Conv2d(16, 1, 3, padding=1))
sd = O.fill_state_dict({k: tuple(v.shape) for k, v in seq.state_dict().items()}, seed=9)
x = gen(4, 64, 16, 16, seed=seed)
seq.load_state_dict(sd); seq = seq.cuda().train()
xg = x.cuda().requires_grad_(True)
y = seq(xg)
gy = gen(*y.shape, seed=99)
y.backward(gy.cuda())
res = {}
for dt in (torch.float64, torch.float32):
    P = {k: (v.to(dt).clone() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    W = O.trainable(P)
    for v in W.values(): v.requires_grad_(True)
    xr = x.to(dt).requires_grad_(True)
    lr = torch.nn.functional.leaky_relu
    h = lr(O._bn(P, "1", O._convT(P, "0", xr, 2, 1, 1), True), 0.01)
    h = O._resblock(P, "3", h, True)
    h = lr(O._bn(P, "5", O._convT(P, "4", h, 2, 1, 1), True), 0.01)
    yr = O._conv(P, "7", h, 1, 1)
    yr.backward(gy.to(dt))
    res[dt] = (yr, xr.grad, W)
yr, dxr, W = res[torch.float64]
d = (xg.grad.cpu().double() - dxr).abs() / dxr.abs().max()
d32 = (res[torch.float32][1].double() - dxr).abs() / dxr.abs().max()
print(f"TC={os.environ.get('CVAE_TC')} seed={seed}: fwd {rel(y, yr):.2e} dx {d.max():.2e} (#>1e-4: {(d > 1e-4).sum().item()}/{d.numel()}; fp32 ref {d32.max():.1e})", end="")
for k, p in seq.named_parameters():
    if p.dim() == 4:
        print(f" {k} {rel(p.grad, W[k].grad):.1e}/{rel(res[torch.float32][2][k].grad, W[k].grad):.0e}", end="")
print()
