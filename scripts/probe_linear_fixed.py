"""Fixed cost of a tensor-core Linear launch (conv_halo_tc in Linear mode): 50 dependent launches captured in one CUDA graph,
time per launch for shrinking problem sizes.  What is left at M = 128, K = 32, N = 16 is launch + prologue + epilogue
latency; the ViT blocks' Linears (M = 4160) pay it 48 times per step.     python scripts/probe_linear_fixed.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from causal_vae_b200 import _lib as L, ops
NL = 50
for M, K, N in ((128, 32, 16), (4160, 32, 16), (4160, 32, 128), (4160, 32, 256), (4160, 256, 16), (4160, 512, 16), (4160, 256, 128), (128, 256, 256), (4160, 256, 256), (4160, 256, 768), (4160, 512, 256), (8320, 256, 256)):
    x = torch.randn(M, 1, 1, K, device="cuda")
    w = torch.randn(N, K, 1, device="cuda") * 0.05
    b = torch.randn(N, device="cuda")
    wt = ops.pack_weight(w, K, K, N, 1, True, K, tc=True)
    outs = [torch.empty(M, 1, 1, N, device="cuda") for _ in range(2)]
    def run():
        for i in range(NL):
            ops.conv_gather(x, wt, b, (1, 1, N), 1, 1, 0, L.MODE_GATHER, tc=True, out=outs[i & 1])
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (10 * NL)
    print(f"linear M={M} K={K} N={N}: {us:.2f} us per launch in a graph ({2.0 * M * K * N / us / 1e6:.1f} TFLOP/s)")
