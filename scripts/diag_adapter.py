import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cvae_oracle as O
from causal_vae_b200 import nn
def rel(a,b):
    a,b=a.detach().double().cpu(),b.detach().double().cpu(); return ((a-b).abs().max()/b.abs().max()).item()
for B in (4, 8, 16):
  for scale in (1.0, 8.0):
    seq = nn.Sequential(nn.Linear(140, 256), nn.BatchNorm1d(256), nn.LeakyReLU(0.2), nn.Linear(256, 512))
    sd = O.fill_state_dict({k: tuple(v.shape) for k, v in seq.state_dict().items()}, seed=13)
    seq.load_state_dict(sd); seq=seq.cuda().train()
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B,140,generator=g)*scale; gy = torch.randn(B,512,generator=g)
    xg = x.cuda().requires_grad_(True); y = seq(xg); y.backward(gy.cuda())
    res={}
    for dt in (torch.float64, torch.float32):
        P={k:(v.to(dt).clone() if v.is_floating_point() else v.clone()) for k,v in sd.items()}
        for v in O.trainable(P).values(): v.requires_grad_(True)
        xr=x.to(dt).requires_grad_(True)
        h = torch.nn.functional.leaky_relu(O._bn(P,"1",O._lin(P,"0",xr),True),0.2); yr=O._lin(P,"3",h)
        yr.backward(gy.to(dt)); res[dt]=(xr.grad, P["0.weight"].grad, P["1.weight"].grad, yr)
    print(B, scale, "fwd", rel(y,res[torch.float64][3]), "dx", rel(xg.grad,res[torch.float64][0]), "noise", rel(res[torch.float32][0],res[torch.float64][0]),
          "dW0", rel(seq[0].weight.grad,res[torch.float64][1]), "noise", rel(res[torch.float32][1],res[torch.float64][1]),
          "dgamma", rel(seq[1].weight.grad,res[torch.float64][2]), "noise", rel(res[torch.float32][2],res[torch.float64][2]))
