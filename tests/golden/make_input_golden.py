"""Golden fixtures for the vessel input pipeline, from the LIVE reference.

    python tests/golden/make_input_golden.py            # writes tests/golden/input_pipeline.json
    python tests/golden/make_input_golden.py --probe    # re-derives the fused / unfused pattern of ATen's loop

Runs the reference's own `VesselDataset.__getitem__` (`vessel_analysis/00_core/dataset.py:193-249`, unmodified)
on synthetic raw images: the dataset object is created without its CSV / TIFF scan (`object.__new__`), given
the attributes `__init__` would have set, and `tifffile.imread` is replaced by a function returning the
synthetic array.  Raw images come from numpy's PCG64 (`raw_image`, machine independent), so only checksums
travel: sha256 of the resized fp32 image (torchvision `Resize(antialias=True)`, dataset.py:186) and of the final
{0,1} mask, the fp32 threshold the reference used, and the packed mask for the small cases.
"""
import hashlib
import importlib.util
import itertools
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
REF = "/root/reference/vessel_analysis/00_core"
sys.path.insert(0, ROOT)
from oracle import input_oracle as IO  # noqa: E402

CASES, raw_image = IO.CASES, IO.raw_image


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_dataset_module():
    for n in ("tifffile",):
        m = types.ModuleType(n)
        m.imread = None
        sys.modules[n] = m
    for p in ("config", "dataset"):
        sys.modules.pop(p, None)
    sys.path.insert(0, REF)
    try:
        spec = importlib.util.spec_from_file_location("ref_vessel_dataset", os.path.join(REF, "dataset.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(REF)
    return mod


def reference_item(mod, raw, H, W, aug_mode, t_idx=3, m_norm=None):
    from torchvision import transforms
    ds = object.__new__(mod.VesselDataset)
    ds.train = True
    ds.data_source = [{"path": "synthetic", "m_norm": np.zeros(12) if m_norm is None else m_norm, "t": t_idx}]
    ds.img_h, ds.img_w = H, W
    ds.base_transform = transforms.Resize((H, W), antialias=True)          # dataset.py:186
    mod.tifffile.imread = lambda path: raw
    x, m, t = ds[aug_mode]                                                 # idx // 4 = 0, idx % 4 = aug_mode
    resized = ds.base_transform(torch.from_numpy(raw).float().unsqueeze(0))[0].numpy()
    return x.numpy(), m.numpy(), t.numpy(), resized


def probe():
    """Which steps of ATen's `t += src*w` loop are fused: exhaustive search per output column."""
    import torch.nn.functional as F
    for win, W in [(1280, 256), (420, 96), (640, 256), (130, 64), (900, 256), (2000, 256)]:
        raw = torch.rand(1, 1, 64, win, generator=torch.Generator().manual_seed(0)) * 1000
        ref = F.interpolate(raw, size=(64, W), mode="bilinear", antialias=True, align_corners=False)[0, 0].numpy()
        src = raw[0, 0].numpy()
        xmin, xsize, w = IO.aa_weights(win, W)
        seen = {}
        for i in range(W):
            n, lo = int(xsize[i]), int(xmin[i])
            if n - 1 > 12:
                continue
            ok = []
            for pat in itertools.product([0, 1], repeat=n - 1):
                t = src[:, lo] * w[i, 0]
                for j in range(1, n):
                    if pat[j - 1]:
                        t = (src[:, lo + j].astype(np.float64) * float(w[i, j]) + t.astype(np.float64)).astype(np.float32)
                    else:
                        t = t + src[:, lo + j] * w[i, j]
                if np.array_equal(t, ref[:, i]):
                    ok.append("".join(map(str, pat)))
            seen[(n, tuple(ok))] = seen.get((n, tuple(ok)), 0) + 1
        print((win, W), torch.backends.cpu.get_cpu_capability())
        for k, v in sorted(seen.items()):
            print("   window", k[0], "fused-step patterns that match", k[1][:4], "columns", v)


def main():
    mod = load_dataset_module()
    mod.CONFIG["T_DIM"] = 19
    out = {"torch": torch.__version__, "cpu_capability": torch.backends.cpu.get_cpu_capability(), "cases": []}
    for name, hin, win, H, W, seed in CASES:
        raw = raw_image(hin, win, seed, constant=name == "constant_image")
        for aug in range(4):
            x, m, t, resized = reference_item(mod, raw, H, W, aug, t_idx=(seed * 5 + aug) % 19)
            o_res = IO.resize_aa(raw, H, W)
            o_mask, o_thr, band = IO.preprocess_image(raw, H, W, aug)
            # the reference's own threshold (fp32 cascade mean) next to the correctly rounded one
            flipped = torch.from_numpy(IO.flip(resized, aug))
            if flipped.max() > flipped.min():
                norm = (flipped - flipped.min()) / (flipped.max() - flipped.min())
            else:
                norm = torch.zeros_like(flipped)
            case = {
                "name": name, "Hin": hin, "Win": win, "H": H, "W": W, "seed": seed, "aug_mode": aug,
                "t_idx": (seed * 5 + aug) % 19, "t_onehot_argmax": int(t.argmax()), "t_onehot_sum": float(t.sum()),
                "resized_sha256": sha(resized), "mask_sha256": sha(x), "mask_sum": float(x.sum()),
                "ref_threshold_hex": np.float32(norm.mean().item()).tobytes().hex(),
                "oracle_threshold_hex": np.float32(o_thr).tobytes().hex(),
                "band_pixels": int(band.sum()),
                "oracle_resize_bit_exact": bool(np.array_equal(o_res, resized)),
                "oracle_mask_mismatch_outside_band": int(((o_mask != x) & ~band).sum()),
                "oracle_mask_mismatch": int((o_mask != x).sum()),
            }
            if H * W <= 64 * 80:
                case["mask_packed_hex"] = np.packbits(x.astype(np.uint8).ravel()).tobytes().hex()
            out["cases"].append(case)
            print(name, aug, "resize exact", case["oracle_resize_bit_exact"], "mask mismatch", case["oracle_mask_mismatch"],
                  "band", case["band_pixels"], "thr ref/oracle", case["ref_threshold_hex"], case["oracle_threshold_hex"])
    # StandardScaler (dataset.py:113-116): the reference calls scikit-learn; record its transform of a seeded table
    from sklearn.preprocessing import StandardScaler
    rng = np.random.Generator(np.random.PCG64(11))
    feats = rng.normal(size=(37, 12)) * rng.uniform(0.1, 50, size=12) + rng.uniform(-5, 5, size=12)
    feats[:, 7] = 2.5                                                     # constant column: scale_ := 1
    sc = StandardScaler().fit(feats)
    tm = torch.tensor(sc.transform(feats)[5], dtype=torch.float32).numpy()  # dataset.py:240
    out["scaler"] = {"seed": 11, "rows": 37, "cols": 12, "row": 5, "m_norm_hex": tm.tobytes().hex(),
                     "mean_hex": sc.mean_.tobytes().hex(), "scale_hex": sc.scale_.tobytes().hex()}
    with open(os.path.join(os.path.dirname(__file__), "input_pipeline.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    probe() if "--probe" in sys.argv else main()
