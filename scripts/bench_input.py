"""Device input pipeline timing: B raw images Hin x Win -> {0,1} masks H x W.

    python scripts/bench_input.py [B Hin Win H W] [--cpu]

CUDA events on the launching stream, L2 flushed between iterations (a 256 MB write), per-kernel times from
separate event pairs.  Algorithmic bytes per image = 4*(Hin*Win + H*W) (raw read once, mask written once).
`--cpu` also times the oracle's CPU restatement of the per-image reference arithmetic on a few images.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from causal_vae_b200.vessel.dataset import VesselBatchTransform  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    B, Hin, Win, H, W = (int(a) for a in args) if len(args) == 5 else (64, 512, 512, 256, 256)
    dev = torch.device("cuda", 0)
    raw = torch.rand(B, Hin, Win, device=dev) * 1000
    aug = torch.arange(B, device=dev, dtype=torch.int32) % 4
    tf = VesselBatchTransform(H, W, 19)
    out = torch.empty(B, 1, H, W, device=dev)
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    for _ in range(3):
        tf.transform(raw, aug, out=out)
    torch.cuda.synchronize()
    times = []
    for _ in range(20):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tf.transform(raw, aug, out=out)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    ms = times[len(times) // 2]
    alg = 4.0 * B * (Hin * Win + H * W)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    res = {"workload": f"{B} x {Hin}x{Win} -> {H}x{W}", "ms": ms, "ms_min": times[0], "images_per_s": B / ms * 1e3,
           "algorithmic_GBps": alg / ms / 1e6, "algorithmic_bytes": alg, "peaks": peaks}
    if "--cpu" in sys.argv:
        from oracle import input_oracle as IO
        r = raw[:2].cpu().numpy()
        t0 = time.perf_counter()
        for i in range(2):
            IO.preprocess_image(r[i], H, W, i)
        res["oracle_numpy_images_per_s"] = 2 / (time.perf_counter() - t0)
        # the reference's own per-sample arithmetic (torchvision Resize + torch ops), one DataLoader worker = 1 thread
        from torchvision import transforms
        torch.set_num_threads(1)
        rs = transforms.Resize((H, W), antialias=True)
        rc = raw[:8].cpu()
        t0 = time.perf_counter()
        for i in range(8):
            im = rs(rc[i:i + 1])
            im = (im - im.min()) / (im.max() - im.min())
            _ = (im > im.mean()).float()
        res["torch_cpu_1thread_images_per_s"] = 8 / (time.perf_counter() - t0)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
