"""Parity of the MNIST (01 / 06 + discriminator), causal_cascade and latent_translator model families
on the native kernels against the oracle (fp64 ground truth) and the live-reference goldens.
Tolerances (north star): 1e-5 relative on losses; 1e-4 of max |g| on gradients (widened only by
the oracle's own fp32-vs-fp64 discrepancy where that is larger); bit-exact on the integer work
(argmax, one-hot)."""
import json
import os

import pytest
import torch

import kinks as K
from oracle import cvae_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(G, name + ".json")) as f:
        return json.load(f)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def closef(a, b, tol=1e-5):
    a, b = float(a), float(b)
    assert abs(a - b) <= tol * max(abs(b), 1e-6), (a, b, abs(a - b) / max(abs(b), 1e-30))


def d64(P):
    return {k: (v.detach().double().clone() if v.is_floating_point() else v.clone()) for k, v in P.items()}


def req(P):
    for v in P.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    return P


def check_grads(model, ref64, ref32=None, floor=1e-4, skip=()):
    """1e-4 of each tensor's max |g| against the fp64 oracle, widened only where the reference's own
    fp32 arithmetic does not reproduce to that level (4 x its fp32-vs-fp64 discrepancy).  The deep families
    (cascade, latent_translator) call it with an oracle evaluated on the native run's kink sides, see tests/kinks.py."""
    worst = []
    for k, p in model.named_parameters():
        if k in skip:
            continue
        g64 = ref64[k].grad
        if g64 is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, f"no grad for {k}"
        noise = rel(ref32[k].grad, g64) if ref32 is not None else 0.0
        tol = max(floor, 4 * noise)
        e = rel(p.grad, g64)
        worst.append((e / tol, k, e, noise))
    worst.sort(reverse=True)
    print("gradient error / tolerance:", [(round(r, 2), k, f"{e:.1e}", f"{n:.1e}") for r, k, e, n in worst[:6]])
    assert worst and worst[0][0] <= 1.0, [(round(r, 2), k, f"{e:.1e}", f"{n:.1e}") for r, k, e, n in worst[:8]]


# ------------------------------------------------------------------------------------------------
# integer work: bit-exact
# ------------------------------------------------------------------------------------------------
def test_argmax_and_one_hot_bit_exact():
    from causal_vae_b200 import functional as F
    gen = torch.Generator().manual_seed(3)
    for rows, T in ((1, 10), (64, 10), (257, 19), (1000, 37)):
        idx = torch.randint(0, T, (rows,), generator=gen)
        oh = F.one_hot(idx.cuda(), T)
        assert torch.equal(oh.cpu(), torch.nn.functional.one_hot(idx, T).float())
        assert torch.equal(F.argmax_rows(oh).cpu(), idx)
        t = torch.randn(rows, T, generator=gen)
        t[::3, 1] = t[::3].max(dim=1).values          # ties: the first maximum wins, as in torch.argmax
        assert torch.equal(F.argmax_rows(t.cuda()).cpu(), t.argmax(1))
    with pytest.raises(RuntimeError):
        F.one_hot(torch.zeros(4, dtype=torch.int32, device="cuda"), 10)


def test_softmax_losses_match_torch():
    from causal_vae_b200 import functional as F
    gen = torch.Generator().manual_seed(5)
    for rows, T in ((64, 10), (300, 19)):
        lg = (torch.randn(rows, T, generator=gen) * 3).double()
        tg = torch.randint(0, T, (rows,), generator=gen)
        a = lg.clone().requires_grad_(True)
        ce = torch.nn.functional.cross_entropy(a, tg)
        ce.backward()
        b = lg.float().cuda().requires_grad_(True)
        got = F.cross_entropy(b, tg.cuda())
        got.backward()
        closef(got, ce)
        assert rel(b.grad, a.grad) <= 1e-5
        a = lg.clone().requires_grad_(True)
        kl = torch.nn.functional.kl_div(torch.log_softmax(a, 1), torch.full_like(a, 1.0 / T), reduction="batchmean") * 1000
        kl.backward()
        b = lg.float().cuda().requires_grad_(True)
        got = F.uniform_kl_batchmean(b, 1000.0)
        got.backward()
        closef(got, kl)
        assert rel(b.grad, a.grad) <= 1e-5


# ------------------------------------------------------------------------------------------------
# MNIST 01 / 06 + discriminator
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["mnist01_M4", "mnist01_M12", "mnist06_M12", "mnist01_M4_b64"])
def test_mnist_adversarial_losses_and_grads(tag):
    from causal_vae_b200.mnist import models, train
    g = load(tag)
    c = g["config"]
    v, M, B = c["variant"], c["M"], c["B"]
    models.CONFIG["M_DIM"], models.CONFIG["T_DIM"], models.CONFIG["Z_DIM"] = M, 10, 10
    P = O.fill_state_dict(O.mnist_shapes(M, 10, 10, v), seed=c["wseed"])
    D = O.fill_state_dict(O.disc_shapes(), seed=c["dseed"])
    gen = torch.Generator().manual_seed(c["xseed"])
    x = torch.rand(B, 1, 28, 28, generator=gen)
    m = torch.rand(B, M, generator=gen)
    t = torch.eye(10)[torch.randint(0, 10, (B,), generator=gen)]
    eps = torch.randn(B, 10, generator=gen)
    eps_adv = torch.randn(B, 10, generator=gen)

    vae = (models.CausalMorphVAE12 if v == "01" else models.CausalMorphVAE12Prob)()
    disc = models.LatentDiscriminator()
    assert {k: tuple(s.shape) for k, s in vae.state_dict().items()} == O.mnist_shapes(M, 10, 10, v)
    assert {k: tuple(s.shape) for k, s in disc.state_dict().items()} == O.disc_shapes()
    vae.load_state_dict(P); disc.load_state_dict(D)
    vae, disc = vae.cuda().train(), disc.cuda().train()
    xc, mc, tc, ec, eac = (a.cuda() for a in (x, m, t, eps, eps_adv))

    P64, D64 = req(d64(P)), req(d64(D))
    ref = O.mnist_vae_loss(P64, D64, x.double(), m.double(), t.double(), eps.double(), eps_adv.double(), variant=v)
    ref[0].backward()
    P32, D32 = req({k: w.clone() for k, w in P.items()}), req({k: w.clone() for k, w in D.items()})
    O.mnist_vae_loss(P32, D32, x, m, t, eps, eps_adv, variant=v)[0].backward()

    got = train.vae_loss(vae, disc, xc, mc, tc, ec, eac, beta=1.0, lambda_adv=10.0)
    got[0].backward()
    for n, a, b in zip(["loss", "recon", "kld", "morph", "adv"], got, ref):
        closef(a, b)
        closef(a, g[n], 2e-5)                         # live-reference golden
    check_grads(vae, P64, P32)
    check_grads(disc, D64, D32)

    # discriminator half: CE(D(z), argmax t), z from a no-grad VAE pass
    for w in D64.values():
        w.grad = None
    ld64 = O.mnist_disc_loss(P64, D64, x.double(), m.double(), t.double(), eps.double(), variant=v)
    ld64.backward()
    disc.zero_grad(set_to_none=True)
    ld = train.disc_loss(vae, disc, xc, mc, tc, ec)
    ld.backward()
    closef(ld, ld64)
    closef(ld, g["loss_d"], 2e-5)
    check_grads(disc, D64)

    # forward tuple arity and shapes (models.py:72 / 06 models.py:85)
    out = vae(xc, mc, tc, ec)
    assert len(out) == (4 if v == "01" else 6)
    assert out[0].shape == (B, 1, 28, 28) and out[1].shape == (B, M)
    ref_out = O.mnist_forward(d64(P), x.double(), m.double(), t.double(), eps.double(), v)
    for a, b in zip(out, ref_out):
        assert rel(a, b) <= 2e-5


def test_mnist_trainer_steps_and_submodule_access():
    from causal_vae_b200.mnist import models, train
    models.CONFIG["M_DIM"], models.CONFIG["T_DIM"], models.CONFIG["Z_DIM"] = 12, 10, 10
    torch.manual_seed(0)
    vae, disc = models.CausalMorphVAE12().cuda(), models.LatentDiscriminator().cuda()
    tr = train.AdversarialTrainer(vae, disc, lr=1e-3)
    gen = torch.Generator().manual_seed(0)
    B = 64
    x = torch.rand(B, 1, 28, 28, generator=gen).cuda()
    m = torch.rand(B, 12, generator=gen).cuda()
    t = torch.eye(10)[torch.randint(0, 10, (B,), generator=gen)].cuda()
    first = last = None
    for i in range(8):
        ld, losses = tr.step(x, m, t)
        last = float(losses[0])
        first = last if first is None else first
    assert last < first
    # callers poke submodules with any batch size, incl. torch.eye(T) (visualize.py:26-35,85-89)
    vae.eval()
    with torch.no_grad():
        m_hat = vae.morph_predictor(torch.eye(10, device="cuda"))
        assert m_hat.shape == (10, 12)
        z = torch.zeros(10, 10, device="cuda")
        img = vae.dec_conv(vae.dec_fc(torch.cat([m_hat, z], dim=1)).view(-1, 64, 7, 7))
        assert img.shape == (10, 1, 28, 28) and float(img.min()) >= 0 and float(img.max()) <= 1
        one = vae.morph_predictor(torch.eye(10, device="cuda")[:1])
        assert rel(one, m_hat[:1]) <= 1e-6


def test_mnist_fused_step_equals_autograd_step():
    """The scoped halves of the adversarial step (in-place parameter gradients, one-launch weight packing, side-stream
    weight gradients) against the plain autograd-accumulation step on identical weights, inputs and noises."""
    from causal_vae_b200.mnist import models, train
    models.CONFIG["M_DIM"], models.CONFIG["T_DIM"], models.CONFIG["Z_DIM"] = 4, 10, 10
    gen = torch.Generator().manual_seed(3)
    B = 64
    x = torch.rand(B, 1, 28, 28, generator=gen).cuda()
    m = torch.rand(B, 4, generator=gen).cuda()
    t = torch.eye(10)[torch.randint(0, 10, (B,), generator=gen)].cuda()
    eps = [torch.randn(B, 10, generator=gen).cuda() for _ in range(3)]
    runs = []
    for fused in (False, True):
        torch.manual_seed(11)
        vae, disc = models.CausalMorphVAE12().cuda(), models.LatentDiscriminator().cuda()
        tr = train.AdversarialTrainer(vae, disc, lr=1e-3, fused=fused)
        out = [tr.step(x, m, t, *eps)]
        gv = tr.opt_vae.flat.grad.clone()                  # the VAE half's gradients of step 1
        out += [tr.step(x, m, t, *eps) for _ in range(2)]
        runs.append((gv, [float(o[0]) for o in out], [[float(v) for v in o[1]] for o in out]))
    (g0, ld0, lv0), (g1, ld1, lv1) = runs
    assert (g0 - g1).abs().max().item() <= 1e-5 * g0.abs().max().item()
    # Adam turns a rounding-level gradient into a +-lr step, so parameters are not compared; the loss trajectory is
    # insensitive to exactly those parameters
    for a, b in zip(ld0, ld1):
        closef(a, b, 1e-4)
    for a, b in zip(sum(lv0, []), sum(lv1, [])):
        closef(a, b, 1e-4)


# ------------------------------------------------------------------------------------------------
# causal_cascade CausalBioVAE
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["cascade", "cascade_b256"])
def test_cascade_forward_loss_grads(tag):
    from causal_vae_b200.cascade import models, train
    g = load(tag)
    c = g["config"]
    P = O.fill_state_dict(O.cascade_shapes(8, 19), seed=c["wseed"])
    gen = torch.Generator().manual_seed(c["xseed"])
    B = c["B"]
    x = torch.randn(B, 1, 64, 64, generator=gen)
    m = torch.rand(B, 8, generator=gen)
    t = torch.randint(0, 19, (B,), generator=gen)
    eps = torch.randn(B, 64, generator=gen)
    model = models.CausalBioVAE(img_channels=1, m_dim=8, t_dim=19, latent_dim=64)
    assert {k: tuple(s.shape) for k, s in model.state_dict().items()} == O.cascade_shapes(8, 19)
    model.load_state_dict(P)
    model = model.cuda().train()

    P64 = req(d64(P))
    outs64 = O.cascade_forward(P64, x.double(), m.double(), t, eps.double(), train=True)
    ref = O.cascade_loss(outs64[0], x.double(), outs64[1], m.double(), outs64[2], outs64[3])
    ref[0].backward()
    P32 = req({k: w.clone() for k, w in P.items()})
    o32 = O.cascade_forward(P32, x, m, t, eps, train=True)
    O.cascade_loss(o32[0], x, o32[1], m, o32[2], o32[3])[0].backward()

    with K.NativeTrace(model) as tr:
        outs = model(x.cuda(), m.cuda(), t.cuda(), eps.cuda())
    assert len(outs) == 4 and outs[0].shape == (B, 1, 64, 64)
    for n, a, b in zip(["recon_x", "m_hat", "mu", "logvar"], outs, outs64):
        assert rel(a, b) <= 2e-5, (n, rel(a, b))
    got = train.loss_function(outs[0], x.cuda(), outs[1], m.cuda(), outs[2], outs[3])
    got[0].backward()
    for n, a, b in zip(["loss", "recon", "m_loss"], got, ref):
        closef(a, b)
        closef(a, g[n], 2e-5)

    # gradients: derivative sides may differ from the fp64 oracle's only inside the forward-tolerance band around a
    # ReLU kink; with the same sides the gradients agree to max(1e-4, 4 x oracle fp32-vs-fp64) -- tests/kinks.py
    pre64 = K.oracle_sides(lambda: O.cascade_forward(d64(P), x.double(), m.double(), t, eps.double(), train=True))
    print("cascade kink sides (differing, units, worst |z|/max):", K.check_sides(tr.masks, pre64, band=2e-5))

    def og(Pp, dt):
        Pp = req(Pp)
        o = O.cascade_forward(Pp, x.to(dt), m.to(dt), t, eps.to(dt), train=True)
        O.cascade_loss(o[0], x.to(dt), o[1], m.to(dt), o[2], o[3])[0].backward()
        return Pp
    with K.with_masks(tr.masks):
        P64m, P32m = og(d64(P), torch.float64), og({k: w.clone() for k, w in P.items()}, torch.float32)
    check_grads(model, P64m, P32m)
    # BatchNorm1d running statistics of mechanism_net.1 after one training forward
    sd = model.state_dict()
    for k in ("mechanism_net.1.running_mean", "mechanism_net.1.running_var"):
        assert rel(sd[k], P64[k]) <= 2e-5, k
    assert int(sd["mechanism_net.1.num_batches_tracked"]) == 1

    # a few optimizer steps reduce the loss; other image sizes are refused loudly (no silent resize)
    tr = train.CascadeTrainer(model, lr=1e-3)
    l0 = float(tr.step(x.cuda(), m.cuda(), t.cuda(), eps.cuda())[0])
    for _ in range(4):
        l1 = float(tr.step(x.cuda(), m.cuda(), t.cuda(), eps.cuda())[0])
    assert l1 < l0
    with pytest.raises(RuntimeError):
        model(torch.randn(2, 1, 96, 96, device="cuda"), m[:2].cuda(), t[:2].cuda())


# ------------------------------------------------------------------------------------------------
# latent_translator ViTVAE
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["latent_translator", "latent_translator_b128"])
def test_latent_translator_forward_loss_grads(tag):
    from causal_vae_b200.latent_translator import engine, models
    g = load(tag)
    c = g["config"]
    H, W, B = c["H"], c["W"], c["B"]
    P = O.fill_state_dict(O.lt_shapes(H, W), seed=c["wseed"])
    gen = torch.Generator().manual_seed(c["xseed"])
    x = torch.rand(B, 1, H, W, generator=gen)
    eps = torch.randn(B, 512, generator=gen)
    model = models.ViTVAE(img_size=(H, W))
    assert {k: tuple(s.shape) for k, s in model.state_dict().items()} == O.lt_shapes(H, W)
    model.load_state_dict(P)
    model = model.cuda().train()
    for mod in model.modules():                      # parity runs: dropout off
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if hasattr(mod, "in_proj_weight"):
            mod.dropout = 0.0

    P64 = req(d64(P))
    rec64, _, mu64, lv64 = O.lt_forward(P64, x.double(), eps.double(), train=True)
    ref = O.lt_loss(rec64, x.double(), mu64, lv64)
    ref[0].backward()
    P32 = req({k: w.clone() for k, w in P.items()})
    r32, _, m32, l32 = O.lt_forward(P32, x, eps, train=True)
    O.lt_loss(r32, x, m32, l32)[0].backward()

    with K.NativeTrace(model, prefix="b.") as tr:
        recons, inp, mu, lv = model(x.cuda(), eps.cuda())
    assert recons.shape == (B, 1, H, W) and mu.shape == (B, 512)
    assert rel(recons, rec64) <= 2e-5 and rel(mu, mu64) <= 2e-5 and rel(lv, lv64) <= 2e-5
    got = engine.loss_function(recons, x.cuda(), mu, lv, beta=1.0)
    got[0].backward()
    for n, a, b in zip(["loss", "recon", "kld"], got, ref):
        closef(a, b)
        closef(a, g[n], 2e-5)

    pre64 = K.oracle_sides(lambda: O.lt_forward(d64(P), x.double(), eps.double(), train=True))
    print("latent_translator kink sides (differing, units, worst |z|/max):", K.check_sides(tr.masks, pre64, band=2e-5))

    def og(Pp, dt):
        Pp = req(Pp)
        r, _, mm, ll = O.lt_forward(Pp, x.to(dt), eps.to(dt), train=True)
        O.lt_loss(r, x.to(dt), mm, ll)[0].backward()
        return Pp
    with K.with_masks(tr.masks):
        P64m, P32m = og(d64(P), torch.float64), og({k: w.clone() for k, w in P.items()}, torch.float32)
    check_grads(model, P64m, P32m)

    # encode-only path used by extract_vit_latents (engine.py:46-50), eval mode
    model.eval()
    with torch.no_grad():
        mu_e, _ = model.encode(x.cuda())
    Pe = {k: v.detach().cpu().double() if v.is_floating_point() else v.cpu() for k, v in model.state_dict().items()}
    mu_ref, _ = O.lt_encode(Pe, x.double(), train=False)
    assert rel(mu_e, mu_ref) <= 2e-5


# ------------------------------------------------------------------------------------------------
# latent translator Ridge + LOOCV on the device
# ------------------------------------------------------------------------------------------------
def test_translator_ridge_loocv_on_device():
    """latent_translator/analysis.py:11-82 with Z, M resident on the GPU: the N + 1 ridge problems as one batched fp64
    solve on the device, against the golden generated from the live reference function (sklearn Ridge under
    LeaveOneOut, tests/golden/make_ridge_golden.py)."""
    import numpy as np
    from causal_vae_b200.latent_translator import analysis
    gold = load("ridge_loocv")
    rng = np.random.default_rng(gold["seed"])
    N, D, Fm = gold["N"], gold["D"], gold["F"]
    Z = rng.standard_normal((N, D)).astype(np.float32)
    Wtrue = rng.standard_normal((D, Fm)) * 0.05
    M = (Z @ Wtrue + 0.1 * rng.standard_normal((N, Fm)) + np.array([1.0, -2.0, 0.5, 0.0, 3.0, -1.0])).astype(np.float32)
    names = [f"f{j}" for j in range(Fm)]
    Zd, Md = torch.from_numpy(Z).cuda(), torch.from_numpy(M).cuda()
    model, metrics, Mhat, W = analysis.fit_translator_ridge(Zd, Md, feature_names=names, alpha=gold["alpha"])
    assert model._w.is_cuda                                              # solved where the latents live
    want = {r["feature"]: r for r in gold["metrics"]}
    for r in metrics.to_dict(orient="records"):
        assert abs(r["r2"] - want[r["feature"]]["r2"]) <= 1e-4, r
        assert abs(r["corr"] - want[r["feature"]]["corr"]) <= 1e-4, r
    assert list(metrics["feature"]) == [r["feature"] for r in gold["metrics"]]
    assert np.abs(Mhat - np.array(gold["Mhat"])).max() <= 1e-4 * np.abs(np.array(gold["Mhat"])).max()
    assert np.abs(W[:, ::37] - np.array(gold["W_sample"])).max() <= 1e-4 * gold["W_absmax"]
    assert np.abs(model.intercept_ - np.array(gold["intercept"])).max() <= 1e-4
    # the encode -> translate chain of latent_translator/main.py: latents straight from the device model
    from causal_vae_b200.latent_translator import engine, models
    vit = models.ViTVAE(img_size=(64, 64)).cuda()
    loader = [{"x": torch.rand(4, 1, 64, 64)} for _ in range(4)]
    Zl = engine.extract_vit_latents(vit, loader, "cuda")
    assert Zl.shape == (16, 512)
    _, met, Mh, _ = analysis.fit_translator_ridge(Zl, M, feature_names=names, device="cuda")
    assert Mh.shape == (16, Fm) and len(met) == Fm
