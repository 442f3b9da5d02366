"""Is the tensor-core chain-gradient mismatch a kink flip (sparse, seed dependent) or a bug?
Runs ConvT-BN-LReLU(s)-ConvT-BN-LReLU(s)-Conv chains with slope s in {0.01, 0.999} through the TC
and SIMT paths and prints gradient errors against the fp64 oracle."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for tc in ("1", "0"):
        for slope in ("0.01", "0.999"):
            env = dict(os.environ, CVAE_TC=tc)
            subprocess.run([sys.executable, __file__, slope], env=env)
    sys.exit(0)
import torch
from oracle import cvae_oracle as O
from causal_vae_b200 import nn
from tests.test_ops_gpu import gen, rel
slope = float(sys.argv[1])
seq = nn.Sequential(nn.ConvTranspose2d(64, 32, 3, 2, 1, 1), nn.BatchNorm2d(32), nn.LeakyReLU(slope),
                    nn.Conv2d(32, 32, 3, 1, 1), nn.BatchNorm2d(32), nn.LeakyReLU(slope),
                    nn.ConvTranspose2d(32, 16, 3, 2, 1, 1), nn.BatchNorm2d(16), nn.LeakyReLU(slope),
                    nn.Conv2d(16, 1, 3, padding=1))
sd = O.fill_state_dict({k: tuple(v.shape) for k, v in seq.state_dict().items()}, seed=9)
x = gen(4, 64, 16, 16, seed=10)
seq.load_state_dict(sd); seq = seq.cuda().train()
xg = x.cuda().requires_grad_(True)
y = seq(xg)
gy = gen(*y.shape, seed=99)
y.backward(gy.cuda())
P = {k: (v.double().clone() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
W = O.trainable(P)
for v in W.values(): v.requires_grad_(True)
xr = x.double().requires_grad_(True)
lr = torch.nn.functional.leaky_relu
h = lr(O._bn(P, "1", O._convT(P, "0", xr, 2, 1, 1), True), slope)
h = lr(O._bn(P, "4", O._conv(P, "3", h, 1, 1), True), slope)
h = lr(O._bn(P, "7", O._convT(P, "6", h, 2, 1, 1), True), slope)
yr = O._conv(P, "9", h, 1, 1)
yr.backward(gy.double())
print(f"TC={os.environ.get('CVAE_TC')} slope={slope}: fwd {rel(y, yr):.2e}  dx {rel(xg.grad, xr.grad):.2e}", end="")
d = (xg.grad.cpu().double() - xr.grad).abs() / xr.grad.abs().max()
print(f"  dx elems > 1e-4: {(d > 1e-4).sum().item()} of {d.numel()}", end="")
for k, p in seq.named_parameters():
    if k.endswith("weight") and p.dim() == 4:
        print(f"  {k} {rel(p.grad, W[k].grad):.1e}", end="")
print()
