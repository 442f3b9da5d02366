"""Attention core forward / backward at the vessel shape (B = 64, S = 65, H = 8, d = 32), CUDA events."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from causal_vae_b200 import ops
B, S, H, d = 64, 65, 8, 32
qkv = torch.randn(B, S, 3 * H * d, device="cuda")
g = torch.randn(B, S, H * d, device="cuda")
once = "--once" in sys.argv
for p in (0.1,):
    out, probs = ops.attention_fwd(qkv, B, S, H, d, p, 1, 1, None)
    dq = ops.attention_bwd(qkv, probs, g, B, S, H, d, p, 1, 1, None)
    torch.cuda.synchronize()
    if once:
        break
    for name, fn in (("fwd", lambda: ops.attention_fwd(qkv, B, S, H, d, p, 1, 1, None)),
                     ("bwd", lambda: ops.attention_bwd(qkv, probs, g, B, S, H, d, p, 1, 1, None))):
        busy = torch.empty(64 << 20, device="cuda")
        ts = []
        for _ in range(10):
            busy.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(f"attention {name} p={p}: {ts[len(ts)//2]*1e3:.1f} us")
