"""LayerNorm backward at the ViT shape (4160 x 256), CUDA events behind a queued spin."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from causal_vae_b200 import ops
rows, D = 4160, 256
x = torch.randn(rows, D, device="cuda"); dy = torch.randn(rows, D, device="cuda")
gamma = torch.rand(D, device="cuda") + 0.5; beta = torch.zeros(D, device="cuda")
y, mean, rstd = ops.layernorm_fwd(x, gamma, beta, rows, D, D, 1e-5)
dx = torch.empty_like(x); dg = torch.zeros(D, device="cuda"); db = torch.zeros(D, device="cuda")
fn = lambda: ops.layernorm_bwd(dy, x, gamma, mean, rstd, rows, D, D, dx, D, False, dg, db)
busy = torch.empty(64 << 20, device="cuda")
ts = []
for _ in range(20):
    busy.add_(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print(f"layernorm_bwd {rows}x{D}: {ts[len(ts)//2]*1e3:.1f} us (rows/block env {os.environ.get('CVAE_LN_RPB','16')})")
