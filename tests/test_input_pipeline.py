"""Vessel input pipeline: resize(antialias) -> flip -> min-max -> mean threshold -> {0,1}.

`not gpu`: the numpy oracle against the goldens made from the live reference `VesselDataset.__getitem__`
(tests/golden/make_input_golden.py).  `gpu`: the CUDA path through the C ABI against the oracle and the goldens —
bit-exact (integer / mask work), plus size-independent properties at the full batch size.
"""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import input_oracle as IO

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "input_pipeline.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hex32(v):
    return np.float32(v).tobytes().hex()


def by_name(name):
    return [c for c in G["cases"] if c["name"] == name]


CASE_IDS = [c[0] for c in IO.CASES]


# ------------------------------------------------------------------------------------------ CPU: oracle pinned
@pytest.mark.parametrize("case", IO.CASES, ids=CASE_IDS)
def test_oracle_matches_live_reference_goldens(case):
    name, hin, win, H, W, seed = case
    raw = IO.raw_image(hin, win, seed, constant=name == "constant_image")
    resized = IO.resize_aa(raw, H, W)
    for g in by_name(name):
        assert g["oracle_resize_bit_exact"] and g["band_pixels"] == 0       # what the generator saw
        assert sha(resized) == g["resized_sha256"], "resize differs from torchvision Resize(antialias=True)"
        mask, thr, band = IO.preprocess_image(raw, H, W, g["aug_mode"])
        assert int(band.sum()) == 0
        assert sha(mask) == g["mask_sha256"], "mask differs from VesselDataset.__getitem__"
        assert float(mask.sum()) == g["mask_sum"]
        assert hex32(thr) == g["oracle_threshold_hex"]
        # the reference's own fp32 cascade mean is within 2 ulp of the correctly rounded mean
        ref_thr = np.frombuffer(bytes.fromhex(g["ref_threshold_hex"]), np.float32)[0]
        assert abs(float(ref_thr) - float(thr)) <= 2 * np.spacing(np.float32(thr))
        if "mask_packed_hex" in g:
            assert np.packbits(mask.astype(np.uint8).ravel()).tobytes().hex() == g["mask_packed_hex"]
        oh = IO.one_hot(np.array([g["t_idx"]]), 19)
        assert int(oh.argmax()) == g["t_onehot_argmax"] and float(oh.sum()) == g["t_onehot_sum"] == 1.0


def _scaler_table():
    s = G["scaler"]
    rng = np.random.Generator(np.random.PCG64(s["seed"]))
    feats = rng.normal(size=(s["rows"], s["cols"])) * rng.uniform(0.1, 50, size=s["cols"]) + rng.uniform(-5, 5, size=s["cols"])
    feats[:, 7] = 2.5
    mean = np.frombuffer(bytes.fromhex(s["mean_hex"]), np.float64)
    scale = np.frombuffer(bytes.fromhex(s["scale_hex"]), np.float64)
    return s, feats, mean, scale


def test_oracle_scaler_matches_sklearn_golden():
    s, feats, mean, scale = _scaler_table()
    m, sc = IO.scaler_fit(feats)
    np.testing.assert_allclose(m, mean, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(sc, scale, rtol=1e-13)
    assert scale[7] == 1.0 and sc[7] == 1.0
    assert IO.scaler_transform(feats, mean, scale)[s["row"]].tobytes().hex() == s["m_norm_hex"]


def test_aa_weights_properties():
    for n_in, n_out in [(512, 256), (300, 128), (96, 128), (1280, 256), (53, 64)]:
        xmin, xsize, w = IO.aa_weights(n_in, n_out)
        assert (xmin >= 0).all() and (xmin + xsize <= n_in).all() and (xsize >= 1).all()
        assert (np.diff(xmin) >= 0).all() and (np.diff(xmin + xsize) >= 0).all()
        np.testing.assert_allclose(w.sum(1), 1.0, atol=3e-7)
        assert w.shape[1] % 2 == 1


# ------------------------------------------------------------------------------------------ GPU: parity
def _transform(H, W):
    from causal_vae_b200.vessel.dataset import VesselBatchTransform
    return VesselBatchTransform(H, W, 19)


@pytest.mark.gpu
@pytest.mark.parametrize("case", IO.CASES, ids=CASE_IDS)
def test_gpu_matches_oracle_and_goldens_bit_exact(case):
    name, hin, win, H, W, seed = case
    raw = IO.raw_image(hin, win, seed, constant=name == "constant_image")
    tf = _transform(H, W)
    batch = torch.from_numpy(np.stack([raw] * 4)).cuda()
    aug = torch.arange(4, dtype=torch.int32, device="cuda")
    x, thr = tf.transform(batch, aug, return_threshold=True)
    torch.cuda.synchronize()
    assert x.shape == (4, 1, H, W) and x.dtype == torch.float32
    resized = tf._ws[(4, batch.device)][0].cpu().numpy()
    o_res = IO.resize_aa(raw, H, W)
    for g in by_name(name):
        k = g["aug_mode"]
        assert np.array_equal(resized[k], IO.flip(o_res, k)), "resized image is not bit-exact"
        mask, o_thr, band = IO.preprocess_image(raw, H, W, k)
        got = x[k].cpu().numpy()
        assert hex32(thr[k].item()) == hex32(o_thr)
        assert np.array_equal(got, mask)
        assert sha(got) == g["mask_sha256"], "mask differs from the live reference's"


@pytest.mark.gpu
def test_gpu_aa_weights_bit_exact():
    from causal_vae_b200.vessel.dataset import _Axis
    rng = np.random.default_rng(3)
    pairs = [(512, 256), (256, 256), (1280, 256), (768, 256), (96, 128), (2000, 101), (379, 289), (2230, 250)]
    pairs += [(int(rng.integers(20, 3000)), int(rng.integers(16, 400))) for _ in range(40)]
    for n_in, n_out in pairs:
        ax = _Axis(n_in, n_out, torch.device("cuda", 0))
        if n_in == n_out:
            assert ax.taps == 1 and ax.xmin.cpu().tolist() == list(range(n_out))
            continue
        xmin, xsize, w = IO.aa_weights(n_in, n_out)
        assert ax.taps == w.shape[1]
        assert np.array_equal(ax.xmin.cpu().numpy(), xmin) and np.array_equal(ax.xsize.cpu().numpy(), xsize)
        assert np.array_equal(ax.w.cpu().numpy().view(np.uint32), w.view(np.uint32)), (n_in, n_out)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(70, 90, 33, 35), (128, 64, 17, 130), (31, 29, 64, 200), (900, 40, 20, 64)])
def test_gpu_ragged_output_sizes(shape):
    """Output widths that are not a multiple of the 64-column tile, pixel counts not a multiple of 4 (scalar
    path), up- and down-scaling mixed per axis, an 45x vertical reduction (smaller row tiles)."""
    hin, win, H, W = shape
    raws = np.stack([IO.raw_image(hin, win, 20 + i) for i in range(3)])
    aug = [3, 0, 1]
    x = _transform(H, W).transform(torch.from_numpy(raws).cuda(), torch.tensor(aug)).cpu().numpy()
    for i in range(3):
        mask, _, band = IO.preprocess_image(raws[i], H, W, aug[i])
        assert not ((x[i] != mask) & ~band).any()
        assert int(((x[i] != mask)).sum()) == 0 or int(band.sum()) > 0


@pytest.mark.gpu
def test_gpu_full_batch_properties():
    """B = 64 raw 512x512 -> 256x256 (the bench workload): properties that need no oracle at this size, plus a
    direct oracle check of three images."""
    g = torch.Generator(device="cuda").manual_seed(5)
    B = 64
    raw = torch.rand(B, 512, 512, device="cuda", generator=g) * 300
    blobs = torch.nn.functional.interpolate(torch.rand(B, 1, 16, 16, device="cuda", generator=g), size=(512, 512),
                                            mode="bilinear")[:, 0] * 900
    raw = (raw + blobs).contiguous()
    tf = _transform(256, 256)
    base = tf.transform(raw).clone()
    assert set(torch.unique(base).tolist()) == {0.0, 1.0}
    frac = base.mean(dim=(1, 2, 3))
    assert (frac > 0.02).all() and (frac < 0.98).all()
    # flips commute with the whole pipeline (dataset.py:219-226 flips before the order-free statistics)
    aug = torch.arange(B, device="cuda", dtype=torch.int32) % 4
    fl = tf.transform(raw, aug).clone()
    for k in range(4):
        ref = base[k::4]
        if k & 1:
            ref = ref.flip(-1)
        if k & 2:
            ref = ref.flip(-2)
        assert torch.equal(fl[k::4], ref)
    # exact invariance under intensity scaling by a power of two and under batch permutation
    assert torch.equal(tf.transform(raw * 4.0), base)
    perm = torch.randperm(B, device="cuda", generator=g)
    assert torch.equal(tf.transform(raw[perm].contiguous()), base[perm])
    # running it again gives the same bits (fp64 atomics: order may differ, the rounded threshold does not)
    assert torch.equal(tf.transform(raw), base)
    for i in (0, 31, 63):
        mask, _, band = IO.preprocess_image(raw[i].cpu().numpy(), 256, 256, 0)
        got = base[i].cpu().numpy()
        assert not ((got != mask) & ~band).any() and int((got != mask).sum()) <= int(band.sum())


@pytest.mark.gpu
def test_gpu_scaler_one_hot_and_edges():
    s, feats, mean, scale = _scaler_table()
    tf = _transform(32, 32).set_scaler(mean, scale)
    out = tf.transform_m(torch.from_numpy(feats)).cpu().numpy()
    assert out[s["row"]].tobytes().hex() == s["m_norm_hex"]
    assert np.array_equal(out, IO.scaler_transform(feats, mean, scale))
    tf2 = _transform(32, 32).fit_scaler(feats)
    np.testing.assert_allclose(tf2.mean_.cpu().numpy(), mean, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(tf2.scale_.cpu().numpy(), scale, rtol=1e-12)
    idx = torch.tensor([0, 18, 7, 7], device="cuda")
    assert np.array_equal(tf.one_hot(idx).cpu().numpy(), IO.one_hot(idx.cpu().numpy(), 19))
    # empty batch, CPU input
    assert tf.transform(torch.empty(0, 40, 40, device="cuda")).shape == (0, 1, 32, 32)
    with pytest.raises(RuntimeError, match="CUDA"):
        tf.transform(torch.zeros(1, 40, 40))


def test_oracle_resize_matches_this_machines_aten_kernel():
    """Live pin of the resize restatement wherever the tests run: ATen's antialiased bilinear kernel (a third-party
    library, not the reference tree) on random (in, out) pairs.  The fused / unfused tap pattern is a property of the
    compiled loop: the AVX2 / AVX512 builds follow the 'aten' pattern, the DEFAULT (no-FMA) build is all-unfused."""
    import torch.nn.functional as TF
    cap = torch.backends.cpu.get_cpu_capability()
    if cap == "DEFAULT":
        fma = False
    elif cap in ("AVX2", "AVX512"):
        fma = "aten"
    else:                      # other ISAs (VSX, ZVECTOR, SVE): pattern not established here
        pytest.skip(f"tap-fusion pattern of the {cap} build of ATen not established")
    rng = np.random.default_rng(11)
    pairs = [(512, 256), (1280, 256), (96, 128), (300, 128), (379, 289), (2230, 250)]
    pairs += [(int(rng.integers(20, 2000)), int(rng.integers(16, 300))) for _ in range(24)]
    for n_in, n_out in pairs:
        raw = rng.uniform(-500, 1000, size=(3, n_in)).astype(np.float32)
        want = TF.interpolate(torch.from_numpy(raw)[None, None], size=(3, n_out), mode="bilinear", antialias=True,
                              align_corners=False)[0, 0].numpy()
        got = IO.resize_aa(raw, 3, n_out, fma)
        assert np.array_equal(got, want), (n_in, n_out, int((got != want).sum()))
    # both axes, the reference's aspect ratio
    raw = rng.uniform(0, 4000, size=(192, 320)).astype(np.float32)
    want = TF.interpolate(torch.from_numpy(raw)[None, None], size=(64, 64), mode="bilinear", antialias=True,
                          align_corners=False)[0, 0].numpy()
    assert np.array_equal(IO.resize_aa(raw, 64, 64, fma), want)
