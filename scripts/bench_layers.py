"""Per-layer timing of the vessel step's conv-family kernels at their real shapes (B = 64, 256x256):
forward / input-gradient (gather family) and weight-gradient, CUDA events on the launching stream,
L2 flushed between launches.  Prints one line per layer with achieved TFLOP/s (algorithmic FLOPs)
and GB/s (algorithmic bytes: every operand / result tensor once).

    python scripts/bench_layers.py [filter] [--once]     # --once: a single launch per layer (for ncu)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from causal_vae_b200 import _lib as L  # noqa: E402
from causal_vae_b200 import ops  # noqa: E402

B = int(os.environ.get("CVAE_BL_B", "64"))          # batch (CVAE_BL_B=16 + CVAE_BL_NOFLUSH=1: operands stay L2-resident)
NOFLUSH = os.environ.get("CVAE_BL_NOFLUSH") == "1"
# name, kind(conv|convT), Cin, Cout, Hin, k, stride, pad, opad
LAYERS = [
    ("stem1 32->64 @128", "conv", 32, 64, 128, 3, 2, 1, 0),
    ("stem2 64->128 @64", "conv", 64, 128, 64, 3, 2, 1, 0),
    ("stem3 128->256 @32", "conv", 128, 256, 32, 3, 2, 1, 0),
    ("stem4 256->256 @16", "conv", 256, 256, 16, 3, 2, 1, 0),
    ("dec0 T256->128 @8", "convT", 256, 128, 8, 3, 2, 1, 1),
    ("res128 @16", "conv", 128, 128, 16, 3, 1, 1, 0),
    ("dec1 T128->64 @16", "convT", 128, 64, 16, 3, 2, 1, 1),
    ("res64 @32", "conv", 64, 64, 32, 3, 1, 1, 0),
    ("dec2 T64->32 @32", "convT", 64, 32, 32, 3, 2, 1, 1),
    ("res32 @64", "conv", 32, 32, 64, 3, 1, 1, 0),
    ("dec3 T32->16 @64", "convT", 32, 16, 64, 3, 2, 1, 1),
    ("dec4 T16->16 @128", "convT", 16, 16, 128, 3, 2, 1, 1),
    ("head 16->1 @256", "conv", 16, 1, 256, 3, 1, 1, 0),
    ("stem0 1->32 @256", "conv", 1, 32, 256, 3, 2, 1, 0),
    ("qkv 256->768 x4160", "linear", 256, 768, 4160, 1, 1, 0, 0),
    ("mlp1 256->512 x4160", "linear", 256, 512, 4160, 1, 1, 0, 0),
    ("mlp2 512->256 x4160", "linear", 512, 256, 4160, 1, 1, 0, 0),
]


def role_times(fn):
    """timing builds (CVAE_TIMING=1): per-role wait / busy cycles of one launch, averaged over CTAs."""
    import ctypes
    buf = (ctypes.c_ulonglong * 16)()
    if L.lib.cvae_debug_read(ctypes.cast(buf, ctypes.c_void_p), 1) != 1:
        return ""
    fn(); torch.cuda.synchronize()
    L.lib.cvae_debug_read(ctypes.cast(buf, ctypes.c_void_p), 1)
    v = [x / 148 / 1e3 for x in buf]
    return (f"  [kclk/CTA] prod wait {v[0]:.0f}/{v[1]:.0f}  mma wT {v[2]:.0f} wA {v[3]:.0f} wB {v[4]:.0f} /{v[5]:.0f}"
            f"  bload wait {v[6]:.0f}/{v[7]:.0f}  epi wait {v[8]:.0f}/{v[9]:.0f}")


def timeit(fn, once, flush):
    if once:
        fn(); torch.cuda.synchronize()
        return float("nan")
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        if not NOFLUSH:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    filt = [a for a in sys.argv[1:] if not a.startswith("--")]
    once = "--once" in sys.argv
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    print(f"{'layer':24s} {'pass':6s} {'ms':>8s} {'TFLOP/s':>8s} {'GB/s':>8s}")
    tot = 0.0
    for name, kind, Ci, Co, Hin, k, st, pad, opad in LAYERS:
        if filt and not any(f in name for f in filt):
            continue
        taps = k * k
        if kind == "linear":
            N, Hs, Ws = 1, 1, Hin
            Hd, Wd = 1, Hin
        else:
            N, Hs, Ws = B, Hin, Hin
            if kind == "conv":
                Hd = Wd = (Hin + 2 * pad - k) // st + 1
            else:
                Hd = Wd = (Hin - 1) * st - 2 * pad + k + opad
        x = torch.randn(N, Hs, Ws, Ci, device="cuda")
        dy = torch.randn(N, Hd, Wd, Co, device="cuda")
        scale, shift, cen = (torch.rand(Ci, device="cuda") + 0.5, torch.randn(Ci, device="cuda"), torch.randn(Ci, device="cuda"))
        xf = ops.XF(scale, shift, 0.01, cen)
        escale, eshift, ecen = (torch.rand(Ci, device="cuda") + 0.5, torch.randn(Ci, device="cuda"), torch.randn(Ci, device="cuda"))
        exf = ops.XF(escale, eshift, 0.01, ecen)
        if kind == "convT":
            w = torch.randn(Ci, Co, taps, device="cuda") * 0.05
            mode_f, mode_b = L.MODE_SCATTER, L.MODE_GATHER
            Mf = N * Hs * Ws          # rows of the smallest phase GEMM
            Mb = N * Hs * Ws
            flops = 2.0 * N * Hs * Ws * Ci * Co * taps
        else:
            w = torch.randn(Co, Ci, taps, device="cuda") * 0.05
            mode_f, mode_b = L.MODE_GATHER, L.MODE_SCATTER
            Mf = N * Hd * Wd
            Mb = N * (Hs // st) * (Ws // st)
            flops = 2.0 * N * Hd * Wd * Ci * Co * taps
        # ---- forward (BN+LReLU of the producer on load, statistics epilogue) ----
        epi_f = L.EPI_STATS if Co > 1 else L.EPI_PLAIN      # the image head feeds the loss directly (no BatchNorm)
        tcf = ops.tc_eligible(Ci, Co, Mf) and not ops.few_eligible(Ci, Co, k, st, pad, mode_f, N, Hs, Ws, Hd, Wd, epi_f)
        if kind == "convT":
            wt_f = ops.pack_weight(w, Ci, Ci, Co, taps, False, Co, tc=tcf)
        else:
            wt_f = ops.pack_weight(w, Ci, Ci, Co, taps, True, Ci, tc=tcf)
        stats = torch.zeros(2 * Co, dtype=torch.float64, device="cuda")
        in_x = xf if (Ci > 1 and kind != "linear") else ops.IDENT      # the ViT Linears read plain matrices (LayerNorm / GELU outputs)
        fwd = lambda: ops.conv_gather(x, wt_f, None, (Hd, Wd, Co), k, st, pad, mode_f, in_x=in_x, epi=epi_f,
                                      stats=stats if Co > 1 else None, tc=tcf)
        ms = timeit(fwd, once, flush)
        by = 4.0 * (x.numel() + dy.numel())
        print(f"{name:24s} {'fwd' + ('*' if tcf else ''):6s} {ms:8.3f} {flops / ms / 1e9:8.1f} {by / ms / 1e6:8.0f}" + (role_times(fwd) if tcf else ""))
        tot += ms
        # ---- input gradient (DACT epilogue: reads the producer's raw output) ----
        if Ci > 1:
            tcb = ops.tc_eligible(Co, Ci, Mb) and not ops.few_eligible(Co, Ci, k, st, pad, mode_b, N, Hd, Wd, Hs, Ws, L.EPI_DACT)
            if kind == "convT":
                wt_b = ops.pack_weight(w, Co, Co, Ci, taps, True, Co, tc=tcb)
            else:
                wt_b = ops.pack_weight(w, Co, Co, Ci, taps, False, Ci, tc=tcb)
            stats_b = torch.zeros(2 * Ci, dtype=torch.float64, device="cuda")
            bwd = lambda: ops.conv_gather(dy, wt_b, None, (Hs, Ws, Ci), k, st, pad, mode_b, epi=L.EPI_DACT, epi_ref=x,
                                          epi_x=exf, stats=stats_b, tc=tcb)
            ms = timeit(bwd, once, flush)
            by = 4.0 * (2 * x.numel() + dy.numel())
            print(f"{name:24s} {'dgrad' + ('*' if tcb else ''):6s} {ms:8.3f} {flops / ms / 1e9:8.1f} {by / ms / 1e6:8.0f}" + (role_times(bwd) if tcb else ""))
            tot += ms
        # ---- weight gradient (+ split-K reduce) ----
        gw = torch.empty_like(w)
        if kind == "convT":
            wg = lambda: ops.conv_wgrad(dy, x, ops.IDENT, in_x, k, st, pad, gw)
        else:
            wg = lambda: ops.conv_wgrad(x, dy, in_x, ops.IDENT, k, st, pad, gw)
        ms = timeit(wg, once, flush)
        by = 4.0 * (x.numel() + dy.numel())
        print(f"{name:24s} {'wgrad':6s} {ms:8.3f} {flops / ms / 1e9:8.1f} {by / ms / 1e6:8.0f}")
        tot += ms
    print(f"total {tot:.3f} ms")


if __name__ == "__main__":
    main()
