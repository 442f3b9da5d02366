"""Drive the UNMODIFIED reference (baseline/_ref, installed by baseline/install_reference.py) on the host cores
or on a CUDA device: the comparator of `bench.py --impl reference`, of the `cpu_baseline` and `eager_gpu_baseline`
legs, and of tests/test_reference_crosscheck_gpu.py.  COMPARATOR ONLY — nothing in causal_vae_b200/ imports this.

The reference's own epoch loops are called as they are wherever they are callable functions:
  * vessel   `train_one_epoch(epoch, vae, train_loader, opt_vae)`   vessel_analysis/01_train/train.py:62-98
  * cascade  `train_one_epoch(model, loader, optimizer, device)`    causal_cascade/train.py:19-40
  * latent_translator `train_vit_vae(model, loader, optimizer, device, epochs, beta)`  latent_translator/engine.py:6-36
The MNIST adversarial step lives inside `train_model()` behind dataset construction
(mnist_test/01_baseline_causal_vae/train.py:13-24), so its loop body (train.py:34-89) is repeated here call for
call on the reference's own `CausalMorphVAE12` / `LatentDiscriminator` modules.
The counterfactual job repeats generate_counterfactual.py:54-55,83-99 + analyze_vessel.py:101-115 on the
reference model's own `dec_adapter` / `backbone.decode`.
"""
import contextlib
import importlib.util
import io
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "vessel_analysis", "00_core", "models.py"))


def _stub(names):
    for n in names:
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__path__ = []
            sys.modules[n] = m


def _load(path, alias, extra_paths=(), purge=("models", "config", "vit_backbone", "train", "dataset", "engine")):
    """Import one reference file under a unique alias (bare names like `models` collide across its directories)."""
    for p in purge:
        sys.modules.pop(p, None)
    dirs = [os.path.dirname(path), *extra_paths]
    for d in dirs:
        sys.path.insert(0, d)
    try:
        spec = importlib.util.spec_from_file_location(alias, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for d in dirs:
            sys.path.remove(d)
    return mod


def _quiet_tqdm():
    _stub(["tqdm"])
    sys.modules["tqdm"].tqdm = _PassThrough


class _PassThrough:
    """tqdm stand-in (the reference wraps its loaders in a progress bar)."""
    def __init__(self, it, **kw):
        self.it = it
    def __iter__(self):
        return iter(self.it)
    def set_postfix(self, *a, **k):
        pass


class Loader(list):
    """A DataLoader stand-in: a list of prepared batches + the `.dataset` the epoch loops take len() of."""
    def __init__(self, batches, n_samples):
        super().__init__(batches)
        self.dataset = range(n_samples)


def set_tf32(torch, on):
    torch.backends.cudnn.allow_tf32 = bool(on)
    torch.backends.cuda.matmul.allow_tf32 = bool(on)


# ------------------------------------------------------------------------------------------------------------
def load_vessel(device, H=256, W=256):
    import torch
    _stub(["matplotlib", "matplotlib.pyplot", "tifffile", "skimage", "skimage.measure", "skimage.morphology", "seaborn"])
    _quiet_tqdm()
    core = os.path.join(REF, "vessel_analysis", "00_core")
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES")
    sys.modules.pop("config", None)
    sys.path.insert(0, core)
    try:
        import config as vcfg
        vcfg.CONFIG["IMG_HEIGHT"], vcfg.CONFIG["IMG_WIDTH"] = H, W
        vcfg.CONFIG["DEVICE"] = torch.device(device)
        models = _load(os.path.join(core, "models.py"), "ref_vessel_models", purge=("models", "vit_backbone"))
        ds = types.ModuleType("dataset")
        ds.VesselDataset = object
        sys.modules["dataset"] = ds
        sys.modules["models"] = models
        train = _load(os.path.join(REF, "vessel_analysis", "01_train", "train.py"), "ref_vessel_train", purge=())
    finally:
        sys.path.remove(core)
        sys.modules.pop("dataset", None)
        if cvd is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = cvd
    return models, train, vcfg.CONFIG


def vessel_trainer(device, B, H=256, W=256, seed=0, state_dict=None, dropout=None):
    """Returns (run(k) -> seconds for k reference steps, model, opt).  One "step" = one iteration of the
    reference's train_one_epoch loop body (train.py:70-92) including its four .item() reads."""
    import torch
    models, train, cfg = load_vessel(device, H, W)
    torch.manual_seed(seed)
    model = models.CausalViTVAE().to(device)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    if dropout is not None:
        for mod in model.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = dropout
            if isinstance(mod, torch.nn.MultiheadAttention):
                mod.dropout = dropout
    opt = torch.optim.Adam(model.parameters(), lr=cfg["LEARNING_RATE"])
    return models, train, model, opt


def time_epoch(fn, sync):
    sync()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        fn()
    sync()
    return time.perf_counter() - t0


def _sync_fn(device):
    import torch
    if str(device).startswith("cuda"):
        return torch.cuda.synchronize
    return lambda: None


def vessel_rate(device, B, steps, warmup, inputs, H=256, W=256, threads=None, tf32=None):
    """samples/s of the unmodified reference step on `device`; `inputs` = (x, m, t) CPU tensors of one batch (the
    loop moves them to the device itself, train.py:71-73)."""
    import torch
    if threads:
        torch.set_num_threads(threads)
    if tf32 is not None:
        set_tf32(torch, tf32)
    _, train, model, opt = vessel_trainer(device, B, H, W)
    x, m, t = inputs
    if str(device).startswith("cuda"):
        x, m, t = x.pin_memory(), m.pin_memory(), t.pin_memory()
    sync = _sync_fn(device)
    time_epoch(lambda: train.train_one_epoch(0, model, Loader([(x, m, t)] * warmup, warmup * B), opt), sync)
    dt = time_epoch(lambda: train.train_one_epoch(1, model, Loader([(x, m, t)] * steps, steps * B), opt), sync)
    return B * steps / dt, dt / steps * 1e3


def vessel_counterfactual_rate(device, sources, chunk, delta=5.0, H=256, W=256, tf32=None, max_seconds=30.0):
    """do(M_k += delta) for every concept of every source, decode, per-image L2 effect
    (generate_counterfactual.py:83-99, analyze_vessel.py:101-115) on the unmodified reference model, eval mode."""
    import torch
    if tf32 is not None:
        set_tf32(torch, tf32)
    models, _, _ = load_vessel(device, H, W)
    torch.manual_seed(0)
    model = models.CausalViTVAE().to(device).eval()
    K, Z = 12, 128
    g = torch.Generator().manual_seed(5)
    m_all = torch.randn(sources, K, generator=g).to(device)
    z_all = torch.randn(sources, Z, generator=g).to(device)
    sync = _sync_fn(device)
    done = 0
    acc = 0.0
    with torch.no_grad():
        def sweep(m, z):
            base = model.backbone.decode(model.dec_adapter(torch.cat([m, z], dim=1)))
            out = []
            for k in range(K):
                mp = m.clone()
                mp[:, k] += delta
                xcf = model.backbone.decode(model.dec_adapter(torch.cat([mp, z], dim=1)))
                out.append(torch.norm((xcf - base).reshape(xcf.shape[0], -1), dim=1))
            return torch.stack(out, 1)
        sweep(m_all[:chunk], z_all[:chunk])
        sync()
        t0 = time.perf_counter()
        for s in range(0, sources, chunk):
            acc += float(sweep(m_all[s:s + chunk], z_all[s:s + chunk]).sum())
            done += min(chunk, sources - s)
            if time.perf_counter() - t0 > max_seconds:
                break
        sync()
        dt = time.perf_counter() - t0
    return done * K / dt, dt * 1e3, done


# ------------------------------------------------------------------------------------------------------------
def load_cascade():
    _quiet_tqdm()
    d = os.path.join(REF, "causal_cascade")
    return _load(os.path.join(d, "models.py"), "ref_cascade_models"), _load(os.path.join(d, "train.py"), "ref_cascade_train")


def cascade_rate(device, B, steps, warmup, threads=None, tf32=None):
    """causal_cascade/train.py:19-40 + main.py:50 (Adam lr 1e-3) on synthetic 64x64 batches."""
    import torch
    if threads:
        torch.set_num_threads(threads)
    if tf32 is not None:
        set_tf32(torch, tf32)
    mod, tr = load_cascade()
    torch.manual_seed(0)
    model = mod.CausalBioVAE(img_channels=1, m_dim=8, t_dim=19).to(device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(11)
    batch = (torch.randn(B, 1, 64, 64, generator=g), torch.rand(B, 8, generator=g), torch.randint(0, 19, (B,), generator=g))
    sync = _sync_fn(device)
    time_epoch(lambda: tr.train_one_epoch(model, Loader([batch] * warmup, warmup * B), opt, device), sync)
    dt = time_epoch(lambda: tr.train_one_epoch(model, Loader([batch] * steps, steps * B), opt, device), sync)
    return B * steps / dt, dt / steps * 1e3


def load_lt():
    d = os.path.join(REF, "latent_translator")
    return _load(os.path.join(d, "models.py"), "ref_lt_models"), _load(os.path.join(d, "engine.py"), "ref_lt_engine")


def lt_rate(device, B, steps, warmup, H=128, W=128, threads=None, tf32=None):
    """latent_translator/engine.py:6-36 (train_vit_vae) + main.py (Adam lr 1e-4) on synthetic 128x128 batches."""
    import torch
    if threads:
        torch.set_num_threads(threads)
    if tf32 is not None:
        set_tf32(torch, tf32)
    mod, eng = load_lt()
    torch.manual_seed(0)
    model = mod.ViTVAE(img_size=(H, W)).to(device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(5)
    batch = {"x": torch.rand(B, 1, H, W, generator=g)}
    sync = _sync_fn(device)
    time_epoch(lambda: eng.train_vit_vae(model, [batch] * warmup, opt, device, 1), sync)
    dt = time_epoch(lambda: eng.train_vit_vae(model, [batch] * steps, opt, device, 1), sync)
    return B * steps / dt, dt / steps * 1e3


def load_mnist(variant="01", M=4, device="cpu"):
    import torch
    d = os.path.join(REF, "mnist_test", "01_baseline_causal_vae" if variant == "01" else "06_model_experiment")
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES")
    sys.modules.pop("config", None)
    sys.path.insert(0, d)
    try:
        import config as mcfg                      # sets CUDA_VISIBLE_DEVICES="0" at import (config.py:4): undone below
    finally:
        sys.path.remove(d)
        if cvd is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = cvd
    mcfg.CONFIG["M_DIM"] = M
    mcfg.CONFIG["DEVICE"] = torch.device(device)
    mod = _load(os.path.join(d, "models.py"), f"ref_mnist{variant}_models", purge=("models",))
    return mod, mcfg.CONFIG


def mnist_step(vae, disc, opt_vae, opt_d, CONFIG, x, m, t):
    """Loop body of mnist_test/01_baseline_causal_vae/train.py:34-89, call for call."""
    import torch
    import torch.nn.functional as F
    x, m, t = x.to(CONFIG["DEVICE"]), m.to(CONFIG["DEVICE"]), t.to(CONFIG["DEVICE"])
    t_indices = torch.argmax(t, dim=1)
    opt_d.zero_grad()
    with torch.no_grad():
        _, _, mu, logvar = vae(x, m, t)
        z = vae.reparameterize(mu, logvar).detach()
        _, _, mu, logvar = vae(x, m, t)
        std = torch.exp(0.5 * logvar)
        eps = torch.randn_like(std)
        z = mu + eps * std
    d_logits = disc(z)
    loss_d = F.cross_entropy(d_logits, t_indices)
    loss_d.backward()
    opt_d.step()
    d_item = loss_d.item()
    opt_vae.zero_grad()
    recon_x, m_hat, mu, logvar = vae(x, m, t)
    loss_recon = F.binary_cross_entropy(recon_x.view(-1, 784), x.view(-1, 784), reduction="sum")
    kld_element = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())
    loss_kld = kld_element * CONFIG["BETA"]
    loss_morph = F.mse_loss(m_hat, m, reduction="sum") * 100
    z_sample = vae.reparameterize(mu, logvar)
    d_logits_fake = disc(z_sample)
    target_uniform = torch.full_like(d_logits_fake, 1.0 / CONFIG["T_DIM"])
    log_probs = F.log_softmax(d_logits_fake, dim=1)
    loss_adv = F.kl_div(log_probs, target_uniform, reduction="batchmean") * CONFIG["LAMBDA_ADV"] * 100
    loss = loss_recon + loss_kld + loss_morph + loss_adv
    loss.backward()
    opt_vae.step()
    return loss.item(), loss_morph.item(), loss_adv.item(), d_item


def mnist_rate(device, B, steps, warmup, M=4, threads=None, tf32=None):
    import torch
    if threads:
        torch.set_num_threads(threads)
    if tf32 is not None:
        set_tf32(torch, tf32)
    mod, CONFIG = load_mnist("01", M, device)
    torch.manual_seed(0)
    vae, disc = mod.CausalMorphVAE12().to(device), mod.LatentDiscriminator().to(device)
    opt_vae = torch.optim.Adam(vae.parameters(), lr=CONFIG["LR"])
    opt_d = torch.optim.Adam(disc.parameters(), lr=CONFIG["LR"])
    g = torch.Generator().manual_seed(19)
    x = torch.rand(B, 1, 28, 28, generator=g)
    m = torch.rand(B, M, generator=g)
    t = torch.eye(10)[torch.randint(0, 10, (B,), generator=g)]
    vae.train(); disc.train()
    sync = _sync_fn(device)
    for _ in range(warmup):
        mnist_step(vae, disc, opt_vae, opt_d, CONFIG, x, m, t)
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        mnist_step(vae, disc, opt_vae, opt_d, CONFIG, x, m, t)
    sync()
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps * 1e3
