"""CPU oracle for the vessel input pipeline.

TEST INFRASTRUCTURE ONLY (same rules as `cvae_oracle.py`): imported by `tests/`, `smoke()` and the CPU legs
of `bench.py`, never by the product package.

Restates, in numpy fp32 with explicit rounding points, what the reference's dataset does to one raw image
(`vessel_analysis/00_core/dataset.py:216-249`):

    Resize((H, W), antialias=True)            dataset.py:186,216  (torchvision -> ATen `_upsample_bilinear2d_aa`,
                                                                     CPU kernel: separable, width pass first)
    hflip / vflip / both by `idx % 4`         dataset.py:219-226
    per-image min-max to [0, 1]               dataset.py:229-232  (all zeros when max == min)
    threshold at the image mean, -> {0,1}     dataset.py:236-237
    one-hot treatment                          dataset.py:243-245
    StandardScaler.transform of the features   dataset.py:113-116  (scikit-learn, fp64: (m - mean_) / scale_)

Third-party arithmetic: the resize is ATen's (`aten/src/ATen/native/cpu/UpSampleKernel.cpp`, torch 2.11.0 in this
image; not under /root/reference).  Its published algorithm is restated in `aa_weights` / `resize_aa`; the
restatement is pinned bit-for-bit against the live `torchvision.transforms.Resize` by
`tests/golden/make_input_golden.py` (fixtures `tests/golden/input_pipeline.json`).

The image mean is the one place where the reference is not reproducible across machines: `Tensor.mean()` is a
vectorised fp32 cascade sum whose association order depends on the CPU's vector width.  The oracle uses the
correctly rounded mean (fp64 sum of the fp32 normalised values, rounded once) and reports the *ambiguity band*:
pixels whose normalised value lies within `BAND_ULPS` fp32 ulps of the threshold, the only pixels whose bit can
depend on the summation order.  Parity = identical masks outside the band; the committed goldens have empty bands.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
BAND_ULPS = 16

# deterministic synthetic raw images shared by the golden generator, the tests, smoke() and the bench
CASES = [  # (name, Hin, Win, H, W, seed)
    ("down2_256", 512, 512, 256, 256, 1),
    ("ragged_128x96", 300, 420, 128, 96, 2),
    ("ref_aspect_768x1280_to_256", 768, 1280, 256, 256, 3),
    ("up_96_to_128", 96, 96, 128, 128, 4),
    ("odd_77x53_to_64", 77, 53, 64, 64, 5),
    ("same_size_64", 64, 64, 64, 64, 6),
    ("width_only_64x200_to_64x80", 64, 200, 64, 80, 7),
    ("constant_image", 40, 40, 32, 32, 8),
]


def raw_image(hin, win, seed, constant=False):
    """Synthetic 'vessel MIP': smooth blobs + noise, float32, arbitrary intensity range (like the TIFFs).
    Only IEEE basic operations and PCG64 (no exp / sin: their SIMD implementations differ between CPUs), so
    every machine generates the same bits and the goldens can carry checksums instead of pixels."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if constant:
        return np.full((hin, win), 3.25, np.float32)
    yy, xx = np.mgrid[0:hin, 0:win].astype(np.float64)
    img = np.zeros((hin, win), np.float64)
    for _ in range(6):
        cy, cx = rng.uniform(0, hin), rng.uniform(0, win)
        s = rng.uniform(0.05, 0.3) * min(hin, win)
        img += rng.uniform(200, 900) / (1.0 + ((yy - cy) ** 2 + (xx - cx) ** 2) / (s * s)) ** 2
    img += rng.uniform(0, 120, size=(hin, win))
    return img.astype(np.float32)


def aa_weights(in_size: int, out_size: int):
    """Index ranges and normalised triangle weights of one antialiased-bilinear axis.

    Follows ATen `HelperInterpBase::_compute_indices_min_size_weights_aa` for `float` input, including which
    intermediate is double and which is rounded to float (`center`, `invscale`, `center -/+ support`, every weight
    and the running total are float; `i + 0.5`, the `+ 0.5` before truncation and the filter argument are double —
    checked against the live kernel on 400 random (in, out) pairs).
    Returns (xmin[int64 out], xsize[int64 out], w[float32 out, max_interp]).
    """
    scale = f32(f32(in_size) / f32(out_size))                 # area_pixel_compute_scale<float>, align_corners=False
    support = f32(scale) if scale >= 1.0 else f32(1.0)        # (interp_size * 0.5) * scale, interp_size = 2
    max_interp = int(np.ceil(support)) * 2 + 1
    invscale = f32(1.0 / float(scale)) if scale >= 1.0 else f32(1.0)
    xmin = np.zeros(out_size, np.int64)
    xsize = np.zeros(out_size, np.int64)
    w = np.zeros((out_size, max_interp), f32)
    for i in range(out_size):
        center = f32(float(scale) * (i + 0.5))
        lo = max(int(float(f32(center - support)) + 0.5), 0)
        n = min(int(float(f32(center + support)) + 0.5), in_size) - lo
        n = min(max(n, 0), max_interp)
        total = f32(0.0)
        for j in range(n):
            arg = f32((float(f32(f32(j + lo) - center)) + 0.5) * float(invscale))
            a = abs(arg)
            wj = f32(1.0) - a if a < 1.0 else f32(0.0)
            w[i, j] = wj
            total = f32(total + wj)
        if total != 0.0:
            w[i, :n] = w[i, :n] / total
        xmin[i], xsize[i] = lo, n
    return xmin, xsize, w


def _pass(src: np.ndarray, xmin, xsize, w, fma) -> np.ndarray:
    """One separable pass along the LAST axis: t = s0*w0; then t += sj*wj for j = 1..n-1, sequentially in fp32.

    Which of those multiply-adds are fused is a property of the compiled ATen loop (`interpolate_aa_single_dim`,
    gcc, x86-64 AVX2 / AVX512 builds of torch 2.11.0 — established by exhaustive search over the 2^(n-1)
    fused / unfused patterns against the live kernel, `tests/golden/make_input_golden.py --probe`): the compiler
    vectorises the products four at a time with an in-order reduction (product rounded, then added) and
    contracts the scalar remainder loop to FMA.  `fma="aten"` restates that: the first 4*floor((n-1)/4) steps
    are unfused, the last (n-1) mod 4 are fused.  `fma=True / False` fuse all / none (the ATEN_CPU_CAPABILITY=
    default build has no FMA and is the all-unfused pattern).
    """
    out = np.empty(src.shape[:-1] + (len(xmin),), f32)
    for i in range(len(xmin)):
        lo, n = int(xmin[i]), int(xsize[i])
        unfused = 4 * ((n - 1) // 4) if fma == "aten" else (0 if fma else n)
        t = src[..., lo] * w[i, 0]
        for j in range(1, n):
            if j - 1 < unfused:
                t = t + src[..., lo + j] * w[i, j]
            else:
                t = (src[..., lo + j].astype(np.float64) * float(w[i, j]) + t.astype(np.float64)).astype(f32)
        out[..., i] = t
    return out


def resize_aa(img: np.ndarray, H: int, W: int, fma="aten") -> np.ndarray:
    """Antialiased bilinear resize of [..., Hin, Win] fp32, width pass first, then height
    (ATen `separable_upsample_generic_Nd_kernel_impl`; an axis whose size does not change is skipped)."""
    x = np.ascontiguousarray(img, f32)
    if x.shape[-1] != W:
        x = _pass(x, *aa_weights(x.shape[-1], W), fma)
    if x.shape[-2] != H:
        xt = np.swapaxes(x, -1, -2)
        x = np.swapaxes(_pass(xt, *aa_weights(xt.shape[-1], H), fma), -1, -2)
    return np.ascontiguousarray(x)


def flip(img: np.ndarray, aug_mode: int) -> np.ndarray:
    """dataset.py:219-226: 1 = hflip, 2 = vflip, 3 = both."""
    if aug_mode & 1:
        img = img[..., ::-1]
    if aug_mode & 2:
        img = img[..., ::-1, :]
    return np.ascontiguousarray(img)


def binarise(img: np.ndarray):
    """dataset.py:229-237 on one [H, W] image.  Returns (mask fp32 {0,1}, threshold fp32, band bool[H, W])."""
    lo, hi = img.min(), img.max()
    if hi > lo:
        norm = ((img - lo) / f32(hi - lo)).astype(f32)
    else:
        norm = np.zeros_like(img)
    thr = f32(norm.astype(np.float64).sum() / norm.size)
    band = np.abs(norm - thr) <= BAND_ULPS * np.spacing(max(thr, f32(2.0 ** -20)))
    return (norm > thr).astype(f32), thr, band


def preprocess_image(raw: np.ndarray, H: int, W: int, aug_mode: int, fma="aten"):
    """raw [Hin, Win] -> (mask [1, H, W], threshold, band)."""
    x = flip(resize_aa(raw, H, W, fma), aug_mode)
    mask, thr, band = binarise(x)
    return mask[None], thr, band[None]


def one_hot(t_idx: np.ndarray, T: int) -> np.ndarray:
    out = np.zeros((len(t_idx), T), f32)
    out[np.arange(len(t_idx)), t_idx] = 1.0
    return out


def scaler_fit(m: np.ndarray):
    """StandardScaler.fit: mean_, scale_ = sqrt(population variance), zero scale -> 1 (fp64)."""
    m = np.asarray(m, np.float64)
    mean = m.mean(0)
    scale = np.sqrt(m.var(0))
    scale[scale < 10 * np.finfo(np.float64).eps] = 1.0
    return mean, scale


def scaler_transform(m: np.ndarray, mean: np.ndarray, scale: np.ndarray) -> np.ndarray:
    """StandardScaler.transform in fp64, then `torch.tensor(..., dtype=float32)` (dataset.py:116,240)."""
    return ((np.asarray(m, np.float64) - mean) / scale).astype(f32)
