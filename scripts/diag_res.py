import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cvae_oracle as O
from causal_vae_b200 import nn
def rel(a,b):
    a,b=a.detach().double().cpu(),b.detach().double().cpu(); return ((a-b).abs().max()/b.abs().max().clamp_min(1e-30)).item()
lr = torch.nn.functional.leaky_relu
def run(kind, C, H, B):
    if kind == "res":
        seq = nn.Sequential(nn.ResBlock(C))
        ref = lambda P, xx: O._resblock(P, "0", xx, True)
    else:
        seq = nn.Sequential(nn.Conv2d(C, C, 3, 1, 1), nn.BatchNorm2d(C), nn.LeakyReLU(0.2), nn.Conv2d(C, C, 3, 1, 1))
        ref = lambda P, xx: O._conv(P, "3", lr(O._bn(P, "1", O._conv(P, "0", xx, 1, 1), True), 0.2), 1, 1)
    sd = O.fill_state_dict({k: tuple(v.shape) for k, v in seq.state_dict().items()}, seed=1)
    seq.load_state_dict(sd); seq = seq.cuda().train()
    g = torch.Generator().manual_seed(C + H + B)
    x = torch.randn(B, C, H, H, generator=g); gy = torch.randn(B, C, H, H, generator=g) + 0.5
    xg = x.cuda().requires_grad_(True); y = seq(xg); y.backward(gy.cuda())
    out = {}
    for dt in (torch.float64, torch.float32):
        P = {k: (v.to(dt).clone() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        for v in O.trainable(P).values(): v.requires_grad_(True)
        xr = x.to(dt).requires_grad_(True); yr = ref(P, xr); yr.backward(gy.to(dt)); out[dt] = (xr.grad, P, yr)
    msg = [f"fwd {rel(y, out[torch.float64][2]):.1e}", f"dx {rel(xg.grad, out[torch.float64][0]):.1e} (n {rel(out[torch.float32][0], out[torch.float64][0]):.1e})"]
    for k, p in seq.named_parameters():
        if k.endswith("weight"):
            msg.append(f"{k} {rel(p.grad, out[torch.float64][1][k].grad):.1e}")
    print(kind, C, H, B, " ".join(msg))
for cfg in [("res",128,4,4),("res",64,4,4),("res",128,8,4),("res",128,4,16),("res",32,4,4),("res",256,2,4),("chain",128,4,4),("chain",64,4,4),("chain",128,8,8)]:
    run(*cfg)
