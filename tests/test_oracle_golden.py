"""Pins the oracle restatement (oracle/cvae_oracle.py) to the golden vectors produced by the
live reference (tests/golden/make_golden.py).  CPU only."""
import json
import os

import pytest
import torch

from oracle import cvae_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 2e-5   # fp32 restatement vs fp32 reference; different op order only


def load(name):
    with open(os.path.join(G, name + ".json")) as f:
        return json.load(f)


def check_summary(t, s, rtol=RTOL, what=""):
    t = t.detach().double().flatten()
    assert t.numel() == s["numel"], what
    scale = max(s["absmax"], 1e-30)
    idx = torch.tensor(s["idx"])
    err = (t[idx] - torch.tensor(s["val"], dtype=torch.float64)).abs().max().item()
    assert err <= rtol * scale + 1e-7, f"{what}: sampled err {err} vs scale {scale}"
    assert abs(t.norm().item() - s["l2"]) <= rtol * max(s["l2"], 1e-6) * 4 + 1e-7, f"{what}: l2"
    if "s101_sum" in s:        # strided ~1 % sample: position-sensitive checksum (sum bounded by sqrt(n) * rtol * scale)
        q = t[::101]
        assert abs(q.sum().item() - s["s101_sum"]) <= rtol * scale * max(q.numel(), 1) ** 0.5 * 4 + 1e-7, f"{what}: s101 sum"
        assert abs(q.norm().item() - s["s101_l2"]) <= rtol * max(s["s101_l2"], scale) * 4 + 1e-7, f"{what}: s101 l2"


def close(a, b, rtol=RTOL):
    assert abs(float(torch.as_tensor(a).detach()) - b) <= rtol * max(abs(b), 1e-6), (float(a), b)


@pytest.mark.parametrize("tag", ["vessel_64x64_b4", "vessel_128x96_b8", "vessel_256x256_b8", "vessel_256x256_b64"])
def test_vessel_oracle_matches_reference(tag):
    g = load(tag)
    c = g["config"]
    H, W, B = c["H"], c["W"], c["B"]
    assert {k: tuple(v) for k, v in g["state_dict_shapes"].items()} == O.vessel_shapes(H, W)
    P = O.fill_state_dict(O.vessel_shapes(H, W), seed=0)
    x, m, t, eps = O.vessel_inputs(B, H, W, seed=0)
    with torch.no_grad():
        outs = O.vessel_forward(P, x, m, t, eps, train=False)
        z = O.reparameterize(outs[2], outs[3], eps)
        xcf = O.vessel_decode(P, O.counterfactual_do(m, 5, delta=5.0), z, (H // 32, W // 32), False)
    for n, o in zip(["recon_x", "m_hat", "mu", "logvar", "m_mu", "m_logvar"], outs):
        check_summary(o, g["eval"][n], what="eval." + n)
    check_summary(xcf, g["eval"]["x_cf_k5_plus5"], what="x_cf")
    l2 = (xcf - outs[0]).flatten(1).norm(dim=1)
    for a, b in zip(l2.tolist(), g["eval"]["cf_l2_per_sample"]):
        close(a, b, 1e-4)

    state = {}
    losses, grads, total = O.vessel_train_step(P, state, 1, x, m, t, eps)
    tr = g["train"]
    for k in ("loss", "recon", "kld", "morph", "sparsity"):
        close(losses[k], tr[k])
    assert set(grads) == set(tr["grads"])
    assert sorted(tr["no_grad_params"]) == sorted(
        k for k in O.trainable(P) if k.startswith(("backbone.fc_mu", "backbone.fc_var")))
    # whole-network gradients are ill-conditioned in fp32 (BN backward cancels the dominant
    # component; the reference's own fp32-vs-fp64 discrepancy is recorded per tensor), so the
    # tolerance is max(1e-4, 4 x that noise floor) of each tensor's max |g|.
    noise = tr["grad_noise_fp32_vs_fp64"]
    for k, s in tr["grads"].items():
        check_summary(grads[k], s, rtol=max(1e-4, 4 * noise[k]), what="grad." + k)
    close(total, tr["grad_total_norm"], max(1e-4, 4 * max(min(v, 1.0) for v in noise.values())))
    for k, s in tr["after_step"].items():
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            check_summary(P[k], s, rtol=1e-4, what="after." + k)


@pytest.mark.parametrize("tag", ["latent_translator", "latent_translator_b128"])
def test_latent_translator_oracle(tag):
    g = load(tag)
    c = g["config"]
    assert {k: tuple(v) for k, v in g["state_dict_shapes"].items()} == O.lt_shapes(c["H"], c["W"])
    P = O.fill_state_dict(O.lt_shapes(c["H"], c["W"]), seed=c["wseed"])
    gen = torch.Generator().manual_seed(c["xseed"])
    x = torch.rand(c["B"], 1, c["H"], c["W"], generator=gen)
    eps = torch.randn(c["B"], 512, generator=gen)
    W = O.trainable(P)
    for v in W.values():
        v.requires_grad_(True)
    rec, _, mu, lv = O.lt_forward(P, x, eps, train=True)
    loss, rl, kl = O.lt_loss(rec, x, mu, lv)
    loss.backward()
    close(loss, g["loss"]); close(rl, g["recon"]); close(kl, g["kld"])
    check_summary(rec, g["outputs"]["recons"]); check_summary(mu, g["outputs"]["mu"])
    for k, s in g["grads"].items():
        check_summary(W[k].grad, s, rtol=max(1e-4, 4 * g["grad_noise_fp32_vs_fp64"][k]), what="grad." + k)


@pytest.mark.parametrize("tag", ["cascade", "cascade_b256"])
def test_cascade_oracle(tag):
    g = load(tag)
    c = g["config"]
    P = O.fill_state_dict(O.cascade_shapes(8, 19), seed=c["wseed"])
    gen = torch.Generator().manual_seed(c["xseed"])
    B = c["B"]
    x = torch.randn(B, 1, 64, 64, generator=gen)
    m = torch.rand(B, 8, generator=gen)
    t = torch.randint(0, 19, (B,), generator=gen)
    eps = torch.randn(B, 64, generator=gen)
    W = O.trainable(P)
    for v in W.values():
        v.requires_grad_(True)
    outs = O.cascade_forward(P, x, m, t, eps, train=True)
    loss, rl, ml = O.cascade_loss(outs[0], x, outs[1], m, outs[2], outs[3])
    loss.backward()
    close(loss, g["loss"]); close(rl, g["recon"]); close(ml, g["m_loss"])
    for n, o in zip(["recon_x", "m_hat", "mu", "logvar"], outs):
        check_summary(o, g["outputs"][n], what=n)
    for k, s in g["grads"].items():
        check_summary(W[k].grad, s, rtol=max(1e-4, 4 * g["grad_noise_fp32_vs_fp64"][k]), what="grad." + k)


@pytest.mark.parametrize("tag", ["mnist01_M4", "mnist01_M12", "mnist06_M12", "mnist01_M4_b64"])
def test_mnist_oracle(tag):
    g = load(tag)
    c = g["config"]
    v, M, B = c["variant"], c["M"], c["B"]
    P = O.fill_state_dict(O.mnist_shapes(M, 10, 10, v), seed=c["wseed"])
    D = O.fill_state_dict(O.disc_shapes(), seed=c["dseed"])
    gen = torch.Generator().manual_seed(c["xseed"])
    x = torch.rand(B, 1, 28, 28, generator=gen)
    m = torch.rand(B, M, generator=gen)
    t = torch.eye(10)[torch.randint(0, 10, (B,), generator=gen)]
    eps = torch.randn(B, 10, generator=gen)
    eps_adv = torch.randn(B, 10, generator=gen)
    for d in (P, D):
        for w in d.values():
            w.requires_grad_(True)
    loss, lr, lk, lm, la = O.mnist_vae_loss(P, D, x, m, t, eps, eps_adv, variant=v)
    loss.backward()
    close(loss, g["loss"]); close(lr, g["recon"]); close(lk, g["kld"]); close(lm, g["morph"]); close(la, g["adv"])
    for k, s in g["grads"].items():
        check_summary(P[k].grad, s, rtol=1e-4, what="grad." + k)
    for w in D.values():
        w.grad = None
    ld = O.mnist_disc_loss(P, D, x, m, t, eps, variant=v)
    ld.backward()
    close(ld, g["loss_d"])
    for k, s in g["disc_grads"].items():
        check_summary(D[k].grad, s, rtol=1e-4, what="dgrad." + k)


def test_ridge_loocv_translator_matches_live_reference():
    """latent_translator/analysis.py::fit_translator_ridge (sklearn Ridge + LeaveOneOut) replayed by the batched
    closed-form solve, against outputs of the live reference (tests/golden/make_ridge_golden.py).  The module is
    loaded by path: it is pure torch and needs no CUDA library."""
    import importlib.util
    import json
    import os
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "ridge_loocv.json")) as f:
        gold = json.load(f)
    spec = importlib.util.spec_from_file_location(
        "lt_analysis", os.path.join(here, "..", "causal_vae_b200", "latent_translator", "analysis.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(gold["seed"])
    N, D, Fm = gold["N"], gold["D"], gold["F"]
    Z = rng.standard_normal((N, D)).astype(np.float32)
    Wtrue = rng.standard_normal((D, Fm)) * 0.05
    M = (Z @ Wtrue + 0.1 * rng.standard_normal((N, Fm)) + np.array([1.0, -2.0, 0.5, 0.0, 3.0, -1.0])).astype(np.float32)
    names = [f"f{j}" for j in range(Fm)]
    model, metrics, Mhat, W = mod.fit_translator_ridge(Z, M, feature_names=names, alpha=gold["alpha"])
    want = {r["feature"]: r for r in gold["metrics"]}
    for r in metrics.to_dict(orient="records"):
        assert abs(r["r2"] - want[r["feature"]]["r2"]) <= 1e-4, r
        assert abs(r["corr"] - want[r["feature"]]["corr"]) <= 1e-4, r
    assert list(metrics["feature"]) == [r["feature"] for r in gold["metrics"]]          # same ranking
    assert np.abs(Mhat - np.array(gold["Mhat"])).max() <= 1e-4 * np.abs(np.array(gold["Mhat"])).max()
    assert np.abs(W[:, ::37] - np.array(gold["W_sample"])).max() <= 1e-4 * gold["W_absmax"]
    assert np.abs(model.intercept_ - np.array(gold["intercept"])).max() <= 1e-4
    assert np.abs(model.predict(Z) - Mhat).max() <= 1e-9


def test_vessel_cnn_oracle_matches_reference():
    """CausalVesselVAE (vessel_analysis/00_core/models.py:9-166) at its fixed 768x1280 input, B = 4."""
    g = load("vessel_cnn_768x1280_b4")
    c = g["config"]
    shapes = O.vessel_cnn_shapes(c["z_dim"], c["m_dim"], c["t_dim"])
    assert {k: tuple(v) for k, v in g["state_dict_shapes"].items()} == shapes
    assert list(g["state_dict_shapes"]) == list(shapes), "key order"
    P = O.fill_state_dict(shapes, seed=0)
    x, m, t, eps = O.vessel_inputs(c["B"], c["H"], c["W"], c["m_dim"], c["t_dim"], c["z_dim"], seed=0)
    names = ["recon_x", "m_hat", "mu", "logvar", "m_mu", "m_logvar"]
    with torch.no_grad():
        outs = O.vessel_cnn_forward(P, x, m, t, eps, train=False)
    for n, o in zip(names, outs):
        check_summary(o, g["eval"][n], what="eval." + n)
    outs, losses, grads = O.vessel_cnn_loss_and_grads(P, x, m, t, eps)
    tr = g["train"]
    for n, o in zip(names, outs):
        check_summary(o, tr["outputs"][n], what="train." + n)
    for k in ("loss", "recon", "kld", "morph", "sparsity"):
        close(losses[k], tr[k])
    assert set(grads) == set(tr["grads"])
    # The reference's own fp32 gradients of this network differ from its fp64 gradients by 1.5-2 % of max |g| for
    # almost every tensor (37 % for dec_fc.3; recorded per tensor) and biases in front of a BatchNorm have an exactly
    # zero gradient, so their fp32 values are pure rounding noise: tolerance = max(1e-4, 4 x that discrepancy),
    # noise-only tensors are bounded against their layer's weight gradient instead.
    noise = tr["grad_noise_fp32_vs_fp64"]
    for k, s in tr["grads"].items():
        if noise[k] > 1.0:
            wk = k[:-len("bias")] + "weight"
            assert grads[k].abs().max().item() <= 1e-3 * tr["grads"][wk]["absmax"], k
        else:
            check_summary(grads[k], s, rtol=min(max(1e-4, 4 * noise[k]), 2.0), what="grad." + k)
    for k, s in tr["running"].items():
        check_summary(P[k], s, rtol=1e-4, what="running." + k)
