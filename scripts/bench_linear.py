"""The ViT blocks' Linear launches in isolation (plain input, plain epilogue, M = 65 * 64 rows): CUDA-event timing with L2
flushed, or `--once` for ncu.   python scripts/bench_linear.py [--once]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from causal_vae_b200 import _lib as L, ops
once = "--once" in sys.argv
M = 4160
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for K, N in ((256, 768), (256, 256), (256, 512), (512, 256), (768, 256)):
    x = torch.randn(M, 1, 1, K, device="cuda")
    w = torch.randn(N, K, 1, device="cuda") * 0.05
    b = torch.randn(N, device="cuda")
    wt = ops.pack_weight(w, K, K, N, 1, True, K, tc=True)
    fn = lambda: ops.conv_gather(x, wt, b, (1, 1, N), 1, 1, 0, L.MODE_GATHER, tc=True)
    fn(); torch.cuda.synchronize()
    if once:
        continue
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"linear M={M} K={K} N={N}: {ts[3]*1e3:.1f} us  ({2.0*M*K*N/ts[3]/1e9:.1f} TFLOP/s)")
