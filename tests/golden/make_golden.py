"""Generate golden fixtures by running the LIVE reference (`/root/reference`) in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/*.json

The reference cannot travel to the GPU box, so the numbers it produces on deterministic
weights (`oracle.cvae_oracle.fill_state_dict`) and deterministic inputs are committed as small
JSON summaries: loss scalars, output checksums, sampled output entries, and per-parameter
gradient norms / sampled entries.  `tests/test_oracle_golden.py` replays them against the
oracle restatement on any machine (that is what pins the oracle); the CUDA path is then compared
with the oracle elementwise.

It also records the reference's own `state_dict` key -> shape tables, which the drop-in modules
must reproduce exactly.
"""
import importlib.util
import json
import os
import sys
import types

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import cvae_oracle as O  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)


def _stub(names):
    for n in names:
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__path__ = []
            sys.modules[n] = m


def load_ref(path, alias, extra_path=None, purge=("models", "config", "vit_backbone", "train", "dataset")):
    """Import a reference file under a unique alias (bare module names collide across dirs)."""
    for p in purge:
        sys.modules.pop(p, None)
    d = os.path.dirname(path)
    sys.path.insert(0, d)
    if extra_path:
        sys.path.insert(0, extra_path)
    try:
        spec = importlib.util.spec_from_file_location(alias, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(d)
        if extra_path:
            sys.path.remove(extra_path)
    return mod


class EpsInjector:
    """Replace torch.randn_like by a queue of prepared tensors (reference draws eps internally)."""
    def __init__(self, *eps):
        self.q = list(eps)
    def __enter__(self):
        self.orig = torch.randn_like
        torch.randn_like = lambda t, **kw: self.q.pop(0).to(t.dtype)
        return self
    def __exit__(self, *a):
        torch.randn_like = self.orig


def summarize(t: torch.Tensor, n=8):
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, min(n, t.numel()), dtype=torch.float64).long().clamp_(max=t.numel() - 1)
    s101 = t[::101]                                  # strided ~1 % sample: a second, position-sensitive checksum
    return {"numel": t.numel(), "sum": t.sum().item(), "l2": t.norm().item(),
            "absmax": t.abs().max().item(), "idx": idx.tolist(), "val": t[idx].tolist(),
            "s101_sum": s101.sum().item(), "s101_l2": s101.norm().item()}


def noise_floor(g32, g64):
    """Reference fp32-vs-fp64 discrepancy per gradient tensor (max abs err / max abs value):
    the conditioning of the reference's own arithmetic, used to calibrate tolerances."""
    out = {}
    for k in g32:
        d = g64[k].detach().double()
        out[k] = ((g32[k].detach().double() - d).abs().max() / d.abs().max().clamp_min(1e-300)).item()
    return out


def set_dropout_zero(model):
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0


def shapes_of(model):
    return {k: list(v.shape) for k, v in model.state_dict().items()}


def golden_vessel(H, W, B, tag, out):
    _stub(["matplotlib", "matplotlib.pyplot", "tifffile", "skimage", "skimage.measure",
           "skimage.morphology", "seaborn", "tqdm"])
    sys.modules["tqdm"].tqdm = lambda x, **k: x
    core = os.path.join(REF, "vessel_analysis/00_core")
    sys.modules.pop("config", None)
    sys.path.insert(0, core)
    import config as vcfg  # noqa
    vcfg.CONFIG["IMG_HEIGHT"], vcfg.CONFIG["IMG_WIDTH"] = H, W
    vcfg.CONFIG["DEVICE"] = torch.device("cpu")
    models = load_ref(os.path.join(core, "models.py"), "ref_vessel_models", purge=("models", "vit_backbone"))
    sys.modules.setdefault("dataset", types.ModuleType("dataset")).VesselDataset = object
    sys.modules["models"] = models
    train = load_ref(os.path.join(REF, "vessel_analysis/01_train/train.py"), "ref_vessel_train", purge=())
    sys.path.remove(core)

    model = models.CausalViTVAE()
    set_dropout_zero(model)
    shp = shapes_of(model)
    assert {k: tuple(v) for k, v in shp.items()} == O.vessel_shapes(H, W), "oracle shape table drifted"
    P = O.fill_state_dict(O.vessel_shapes(H, W), seed=0)
    model.load_state_dict(P, strict=True)
    x, m, t, eps = O.vessel_inputs(B, H, W, seed=0)

    rec = {"config": {"H": H, "W": W, "B": B, "seed": 0, "beta": 0.5}, "state_dict_shapes": shp}

    # --- eval-mode forward + counterfactual decode ---------------------------------------
    model.eval()
    with torch.no_grad(), EpsInjector(eps):
        outs = model(x, m, t)
        z = O.reparameterize(outs[2], outs[3], eps)
        mp = m.clone(); mp[:, 5] = mp[:, 5] + 5.0
        xcf = model.backbone.decode(model.dec_adapter(torch.cat([mp, z], dim=1)))
    rec["eval"] = {n: summarize(o) for n, o in zip(
        ["recon_x", "m_hat", "mu", "logvar", "m_mu", "m_logvar"], outs)}
    rec["eval"]["x_cf_k5_plus5"] = summarize(xcf)
    rec["eval"]["cf_l2_per_sample"] = (xcf - outs[0]).flatten(1).norm(dim=1).double().tolist()

    # --- train-mode step: losses, grads, clip, Adam ---------------------------------------
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    opt.zero_grad()
    with EpsInjector(eps):
        outs = model(x, m, t)
    recon, kld, morph, sp = train.loss_function(outs[0], x, outs[1], m, outs[2], outs[3], outs[4], outs[5])
    loss = recon + 0.5 * kld + morph + 0.3 * sp
    loss.backward()
    rec["train"] = {"loss": loss.item(), "recon": recon.item(), "kld": kld.item(),
                    "morph": morph.item(), "sparsity": sp.item(),
                    "outputs": {n: summarize(o) for n, o in zip(
                        ["recon_x", "m_hat", "mu", "logvar", "m_mu", "m_logvar"], outs)},
                    "grads": {k: summarize(p.grad) for k, p in model.named_parameters() if p.grad is not None},
                    "no_grad_params": [k for k, p in model.named_parameters() if p.grad is None]}
    g32 = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    m64 = models.CausalViTVAE().double()
    set_dropout_zero(m64)
    m64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in P.items()})
    m64.train()
    with EpsInjector(eps.double()):
        o64 = m64(x.double(), m.double(), t.double())
    r64 = train.loss_function(o64[0], x.double(), o64[1], m.double(), o64[2], o64[3], o64[4], o64[5])
    (r64[0] + 0.5 * r64[1] + r64[2] + 0.3 * r64[3]).backward()
    rec["train"]["grad_noise_fp32_vs_fp64"] = noise_floor(
        g32, {k: p.grad for k, p in m64.named_parameters() if p.grad is not None})
    rec["train"]["loss_fp64"] = (r64[0] + 0.5 * r64[1] + r64[2] + 0.3 * r64[3]).item()
    total = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=5.0)
    opt.step()
    rec["train"]["grad_total_norm"] = float(total)
    sd = model.state_dict()
    rec["train"]["after_step"] = {k: summarize(v) for k, v in sd.items()
                                  if k.endswith(("running_mean", "running_var", "num_batches_tracked"))
                                  or k in ("backbone.stem.0.weight", "backbone.decoder_input.weight",
                                           "backbone.decoder.18.weight", "enc_adapter.3.bias",
                                           "morph_predictor_mu.weight", "backbone.cls_token",
                                           "backbone.transformer.3.attn.in_proj_weight")}
    out[tag] = rec


def golden_lt(H, W, B, out, tag="latent_translator"):
    mod = load_ref(os.path.join(REF, "latent_translator/models.py"), "ref_lt_models")
    model = mod.ViTVAE(img_size=(H, W))
    set_dropout_zero(model)
    shp = shapes_of(model)
    assert {k: tuple(v) for k, v in shp.items()} == O.lt_shapes(H, W), "lt shape table drifted"
    P = O.fill_state_dict(O.lt_shapes(H, W), seed=3)
    model.load_state_dict(P, strict=True)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, 1, H, W, generator=g)
    eps = torch.randn(B, 512, generator=g)
    model.train()
    with EpsInjector(eps):
        rec_x, _, mu, lv = model(x)
    rl = torch.nn.functional.mse_loss(rec_x, x, reduction="mean")
    kl = -0.5 * torch.mean(1 + lv - mu.pow(2) - lv.exp())
    loss = rl + 1.0 * kl
    loss.backward()
    m64 = mod.ViTVAE(img_size=(H, W)).double()
    set_dropout_zero(m64)
    m64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in P.items()})
    m64.train()
    with EpsInjector(eps.double()):
        r64, _, mu64, lv64 = m64(x.double())
    (torch.nn.functional.mse_loss(r64, x.double()) - 0.5 * torch.mean(1 + lv64 - mu64.pow(2) - lv64.exp())).backward()
    noise = noise_floor({k: p.grad for k, p in model.named_parameters() if p.grad is not None},
                        {k: p.grad for k, p in m64.named_parameters() if p.grad is not None})
    out[tag] = {
        "grad_noise_fp32_vs_fp64": noise,
        "config": {"H": H, "W": W, "B": B, "wseed": 3, "xseed": 5}, "state_dict_shapes": shp,
        "loss": loss.item(), "recon": rl.item(), "kld": kl.item(),
        "outputs": {"recons": summarize(rec_x), "mu": summarize(mu), "log_var": summarize(lv)},
        "grads": {k: summarize(p.grad) for k, p in model.named_parameters() if p.grad is not None}}


def golden_cascade(B, out, tag="cascade"):
    mod = load_ref(os.path.join(REF, "causal_cascade/models.py"), "ref_cascade_models")
    _stub(["tqdm"]); sys.modules["tqdm"].tqdm = lambda x, **k: x
    tr = load_ref(os.path.join(REF, "causal_cascade/train.py"), "ref_cascade_train")
    model = mod.CausalBioVAE(img_channels=1, m_dim=8, t_dim=19)
    shp = shapes_of(model)
    assert {k: tuple(v) for k, v in shp.items()} == O.cascade_shapes(8, 19), "cascade shape table drifted"
    model.load_state_dict(O.fill_state_dict(O.cascade_shapes(8, 19), seed=7), strict=True)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 1, 64, 64, generator=g)
    m = torch.rand(B, 8, generator=g)
    t = torch.randint(0, 19, (B,), generator=g)
    eps = torch.randn(B, 64, generator=g)
    model.train()
    with EpsInjector(eps):
        outs = model(x, m, t)
    loss, rl, ml = tr.loss_function(outs[0], x, outs[1], m, outs[2], outs[3])
    loss.backward()
    # the reference's forward casts one_hot to fp32, so its fp64 copy cannot run; the noise
    # floor is taken from the (forward-pinned) oracle restatement in fp32 vs fp64 instead.
    def ograds(dt):
        Pq = {k: (v.to(dt) if v.is_floating_point() else v)
              for k, v in O.fill_state_dict(O.cascade_shapes(8, 19), seed=7).items()}
        Wq = O.trainable(Pq)
        for v in Wq.values():
            v.requires_grad_(True)
        o = O.cascade_forward(Pq, x.to(dt), m.to(dt), t, eps.to(dt), True)
        O.cascade_loss(o[0], x.to(dt), o[1], m.to(dt), o[2], o[3])[0].backward()
        return {k: v.grad for k, v in Wq.items()}
    noise = noise_floor(ograds(torch.float32), ograds(torch.float64))
    out[tag] = {
        "grad_noise_fp32_vs_fp64": noise,
        "config": {"B": B, "wseed": 7, "xseed": 11}, "state_dict_shapes": shp,
        "loss": loss.item(), "recon": rl.item(), "m_loss": ml.item(),
        "outputs": {n: summarize(o) for n, o in zip(["recon_x", "m_hat", "mu", "logvar"], outs)},
        "grads": {k: summarize(p.grad) for k, p in model.named_parameters() if p.grad is not None}}


def golden_mnist(variant, M, B, out, tag=None):
    d = os.path.join(REF, "mnist_test", "01_baseline_causal_vae" if variant == "01" else "06_model_experiment")
    sys.modules.pop("config", None)
    sys.path.insert(0, d)
    import config as mcfg  # noqa
    sys.path.remove(d)
    os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    mcfg.CONFIG["M_DIM"] = M
    mcfg.CONFIG["DEVICE"] = torch.device("cpu")
    mod = load_ref(os.path.join(d, "models.py"), f"ref_mnist{variant}_models", purge=("models",))
    vae, disc = mod.CausalMorphVAE12(), mod.LatentDiscriminator()
    shp, dshp = shapes_of(vae), shapes_of(disc)
    assert {k: tuple(v) for k, v in shp.items()} == O.mnist_shapes(M, 10, 10, variant)
    assert {k: tuple(v) for k, v in dshp.items()} == O.disc_shapes()
    vae.load_state_dict(O.fill_state_dict(O.mnist_shapes(M, 10, 10, variant), seed=13))
    disc.load_state_dict(O.fill_state_dict(O.disc_shapes(), seed=17))
    g = torch.Generator().manual_seed(19)
    x = torch.rand(B, 1, 28, 28, generator=g)
    m = torch.rand(B, M, generator=g)
    t = torch.eye(10)[torch.randint(0, 10, (B,), generator=g)]
    eps = torch.randn(B, 10, generator=g)
    eps_adv = torch.randn(B, 10, generator=g)
    F = torch.nn.functional
    # literal replay of the VAE half of train.py:65-87 (06: :67-94)
    with EpsInjector(eps, eps_adv):
        outs = vae(x, m, t)
        recon_x, m_hat, mu, logvar = outs[:4]
        loss_recon = F.binary_cross_entropy(recon_x.view(-1, 784), x.view(-1, 784), reduction="sum")
        loss_kld = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp()) * mcfg.CONFIG["BETA"]
        if variant == "01":
            loss_morph = F.mse_loss(m_hat, m, reduction="sum") * 100
        else:
            loss_morph = 0.5 * torch.sum(outs[5] + (m - outs[4]) ** 2 / outs[5].exp())
        z_sample = vae.reparameterize(mu, logvar)
        logits = disc(z_sample)
        target_uniform = torch.full_like(logits, 1.0 / 10)
        loss_adv = F.kl_div(F.log_softmax(logits, dim=1), target_uniform, reduction="batchmean") \
            * mcfg.CONFIG["LAMBDA_ADV"] * 100
        loss = loss_recon + loss_kld + loss_morph + loss_adv
    loss.backward()
    vae_grads = {k: summarize(p.grad) for k, p in vae.named_parameters() if p.grad is not None}
    # discriminator half (train.py:41-60)
    disc.zero_grad()
    with torch.no_grad():
        z = O.reparameterize(mu, logvar, eps)
    loss_d = F.cross_entropy(disc(z.detach()), t.argmax(1))
    loss_d.backward()
    out[tag or f"mnist{variant}_M{M}"] = {
        "config": {"B": B, "M": M, "wseed": 13, "dseed": 17, "xseed": 19, "variant": variant},
        "state_dict_shapes": shp, "disc_shapes": dshp,
        "loss": loss.item(), "recon": loss_recon.item(), "kld": loss_kld.item(),
        "morph": loss_morph.item(), "adv": loss_adv.item(), "loss_d": loss_d.item(),
        "outputs": {n: summarize(o) for n, o in zip(
            ["recon_x", "m_hat", "mu", "logvar", "m_mu", "m_logvar"], outs)},
        "grads": vae_grads,
        "disc_grads": {k: summarize(p.grad) for k, p in disc.named_parameters()}}


def main():
    out = {}
    golden_vessel(64, 64, 4, "vessel_64x64_b4", out)
    golden_vessel(128, 96, 8, "vessel_128x96_b8", out)
    golden_vessel(256, 256, 8, "vessel_256x256_b8", out)
    golden_lt(128, 128, 2, out)
    golden_cascade(8, out)
    golden_mnist("01", 4, 8, out)
    golden_mnist("01", 12, 8, out)
    golden_mnist("06", 12, 8, out)
    # the BASELINE.json batch sizes (configs[3], [2], [1], [0]): B = 64 / 128 / 256 / 64
    golden_vessel(256, 256, 64, "vessel_256x256_b64", out)
    golden_lt(128, 128, 128, out, tag="latent_translator_b128")
    golden_cascade(256, out, tag="cascade_b256")
    golden_mnist("01", 4, 64, out, tag="mnist01_M4_b64")
    for k, v in out.items():
        with open(os.path.join(os.path.dirname(__file__), k + ".json"), "w") as f:
            json.dump(v, f, indent=1)
        print("wrote", k, "loss =", v.get("loss", v.get("train", {}).get("loss")))


if __name__ == "__main__":
    main()
