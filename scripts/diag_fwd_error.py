"""Forward / input-gradient error of the tensor-core conv kernels on the vessel layer shapes against torch fp64
(conv only, no BatchNorm): max |err| / max |y|.  Run once per setting of CVAE_XACC (0: cross terms in the main
accumulator, 3 MMAs per product; 1: own accumulator columns, 2 MMAs) to see what the accumulator split buys.

    CVAE_XACC=0 python scripts/diag_fwd_error.py ; CVAE_XACC=1 python scripts/diag_fwd_error.py
"""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from causal_vae_b200 import nn as N  # noqa: E402


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def run(name, ours, ref, shape, B=8):
    g = torch.Generator().manual_seed(0)
    sd = {k: torch.randn(v.shape, generator=g) * (0.05 if v.dim() > 1 else 0.3) for k, v in ours.state_dict().items()}
    ours.load_state_dict(sd)
    ref.load_state_dict({k: v.double() for k, v in sd.items()})
    ours, ref = ours.cuda(), ref.double().cuda()
    x = torch.randn(B, *shape, generator=g).cuda()
    xo, xr = x.clone().requires_grad_(True), x.double().requires_grad_(True)
    yo, yr = N.Sequential(ours)(xo), ref(xr)
    dy = torch.randn(yr.shape, generator=g).cuda()
    yo.backward(dy.float())
    yr.backward(dy.double())
    torch.backends.cudnn.allow_tf32 = False
    y32 = ref.float()(x)                              # what eager fp32 (cuDNN, TF32 off) gets on the same inputs
    print(f"{name:28s} K={shape[0] * 9:5d}  y {rel(yo, yr):.1e}  dx {rel(xo.grad, xr.grad):.1e}  "
          f"dw {rel(ours.weight.grad, ref.weight.grad.double() if False else ref.double().weight.grad):.1e}  "
          f"[eager fp32 y {rel(y32, yr):.1e}]", flush=True)


T = torch.nn
print("CVAE_XACC =", os.environ.get("CVAE_XACC", "1"))
for a, b, h in [(32, 64, 128), (64, 128, 64), (128, 256, 32), (256, 256, 16)]:
    run(f"stem conv s2 {a}->{b} @{h}", N.Conv2d(a, b, 3, 2, 1), T.Conv2d(a, b, 3, 2, 1), (a, h, h))
for c, h in [(128, 16), (64, 32), (32, 64)]:
    run(f"res conv s1 {c}->{c} @{h}", N.Conv2d(c, c, 3, 1, 1), T.Conv2d(c, c, 3, 1, 1), (c, h, h))
for a, b, h in [(256, 128, 8), (128, 64, 16), (64, 32, 32), (32, 16, 64)]:
    run(f"dec convT {a}->{b} @{h}", N.ConvTranspose2d(a, b, 3, 2, 1, output_padding=1),
        T.ConvTranspose2d(a, b, 3, 2, 1, output_padding=1), (a, h, h), B=64 if h <= 16 else 8)
for a, b in [(256, 768), (256, 512), (512, 256)]:
    g = torch.Generator().manual_seed(1)
    lo, lr = N.Linear(a, b), T.Linear(a, b)
    sd = {k: torch.randn(v.shape, generator=g) * 0.05 for k, v in lo.state_dict().items()}
    lo.load_state_dict(sd); lr.load_state_dict(sd)
    lo, lr = lo.cuda(), lr.double().cuda()
    x = torch.randn(4160, a, generator=g).cuda()
    print(f"linear {a}->{b} x4160          y {rel(lo(x), lr(x.double())):.1e}", flush=True)
