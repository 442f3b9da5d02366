"""Role timers of ONE tensor-core Linear launch (timing build: python causal_vae_b200/_build.py --timing, then
CVAE_LIB=causal_vae_b200/libcvae_b200_timing.so python scripts/probe_linear_roles.py): kclk per CTA spent by the
producer (mbarrier wait, cp.async wait, tcgen05.wait::st, total), the MMA issuer, the weight loader and the epilogue."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from causal_vae_b200 import _lib as L, ops
buf = (ctypes.c_ulonglong * 16)()
for M, K, N in ((4160, 512, 16), (4160, 256, 128), (4160, 256, 256), (4160, 256, 768), (4160, 768, 256)):
    x = torch.randn(M, 1, 1, K, device="cuda")
    w = torch.randn(N, K, 1, device="cuda") * 0.05
    b = torch.randn(N, device="cuda")
    wt = ops.pack_weight(w, K, K, N, 1, True, K, tc=True)
    fn = lambda: ops.conv_gather(x, wt, b, (1, 1, N), 1, 1, 0, L.MODE_GATHER, tc=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if L.lib.cvae_debug_read(ctypes.cast(buf, ctypes.c_void_p), 1) != 1:
        print("not a timing build"); break
    fn(); torch.cuda.synchronize()
    L.lib.cvae_debug_read(ctypes.cast(buf, ctypes.c_void_p), 1)
    bn = 128 if N % 128 == 0 else (64 if N % 64 == 0 else (32 if N % 32 == 0 else 16))
    ctas = min(148, ((M + 127) // 128) * (N // bn))
    v = [t / ctas / 1e3 for t in buf]
    print(f"M={M} K={K} N={N} ({ctas} CTAs, {(K + 31) // 32} k-blocks): producer mbar-wait {v[0]:.1f} cp.async-wait {v[10]:.1f} st-wait {v[11]:.1f} "
          f"total {v[1]:.1f} | mma wT {v[2]:.1f} wA {v[3]:.1f} wB {v[4]:.1f} total {v[5]:.1f} | loader wait {v[6]:.1f}/{v[7]:.1f} | "
          f"epilogue wait {v[8]:.1f}/{v[9]:.1f}  [kclk per CTA]")
