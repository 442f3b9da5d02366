"""CPU oracle for the CausalVAE training / counterfactual hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`causal_vae_b200/`) may import
this file; it is used by `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` as the checker and the timed CPU baseline.

It is a *functional restatement* (plain `torch` on CPU, fp32 or fp64, pure functions over a
`state_dict`-shaped dict of tensors) of the arithmetic of the reference repository
bjo5029/causal-vae.  Every function cites the reference file:line it follows.  The
restatement is pinned against the live reference modules by `tests/golden/make_golden.py`
(run in the build container, where `/root/reference` is importable) whose outputs are
committed under `tests/golden/*.json`; `tests/test_oracle_golden.py` replays them on any box.

Third-party arithmetic: PyTorch itself (conv / batch-norm / layer-norm / softmax / erf-GELU /
Adam semantics).  The reference pins no version; torch 2.11.0 of this image is the pin.

Randomness: the reference draws `eps = randn_like(std)` and dropout masks internally.  The
oracle takes `eps` as an explicit argument and only models dropout with p = 0 (parity runs
use `eval()` or p = 0; training-mode dropout is validated statistically elsewhere).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
LN_EPS = 1e-5


# ----------------------------------------------------------------------------------------
# deterministic weights / inputs (shared by the golden generator, the tests and the bench)
# ----------------------------------------------------------------------------------------
def fill_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0, dtype=torch.float32) -> SD:
    """Deterministic, machine-independent parameter values for a name->shape table.

    Keys are visited in sorted order with one CPU generator, so any implementation that
    exposes the same `state_dict` keys/shapes can be loaded with identical bits without
    shipping a checkpoint.  Scales are chosen so activations stay O(1) through the stacks.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    out: SD = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        leaf = k.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[k] = torch.zeros((), dtype=torch.int64)
            continue
        r = torch.randn(shp, generator=g, dtype=torch.float32)
        if leaf == "running_mean":
            v = 0.1 * r
        elif leaf == "running_var":
            v = 1.0 + 0.2 * r.abs()
        elif leaf in ("pos_embedding", "cls_token"):
            v = 0.5 * r
        elif leaf == "in_proj_weight":
            v = r / math.sqrt(shp[1])
        elif leaf in ("in_proj_bias",):
            v = 0.05 * r
        elif leaf == "weight" and len(shp) == 1:      # BN / LN gain
            v = 1.0 + 0.1 * r
        elif leaf == "bias":
            v = 0.05 * r
        elif leaf == "weight" and len(shp) == 2:      # Linear (out, in)
            v = r / math.sqrt(shp[1])
        elif leaf == "weight" and len(shp) == 4:      # Conv (Cout,Cin,kh,kw) / ConvT (Cin,Cout,kh,kw)
            v = r / math.sqrt(shp[1] * shp[2] * shp[3])
        else:
            v = 0.1 * r
        out[k] = v.to(dtype)
    return out


def vessel_inputs(B: int, H: int, W: int, m_dim: int = 12, t_dim: int = 19, z_dim: int = 128,
                  seed: int = 0):
    """Synthetic vessel batch: binary x (~20 % foreground,
    vessel_analysis/00_core/dataset.py:236-237), standardised m (dataset.py:113-116),
    one-hot t, and the reparameterisation noise eps."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    x = (torch.rand(B, 1, H, W, generator=g) > 0.8).float()
    m = torch.randn(B, m_dim, generator=g)
    ti = torch.randint(0, t_dim, (B,), generator=g)
    t = torch.eye(t_dim)[ti]
    g2 = torch.Generator(device="cpu")
    g2.manual_seed(seed + 1)
    eps = torch.randn(B, z_dim, generator=g2)
    return x, m, t, eps


# ----------------------------------------------------------------------------------------
# small building blocks
# ----------------------------------------------------------------------------------------
def _bn(P: SD, pre: str, x: Tensor, train: bool) -> Tensor:
    """nn.BatchNorm{1,2}d forward incl. running-stat EMA (momentum 0.1, unbiased var) —
    torch semantics used at vit_backbone.py:76-89, models.py:227,237."""
    rm, rv = P[pre + ".running_mean"], P[pre + ".running_var"]
    if train:
        dims = [0] + list(range(2, x.dim()))
        n = x.numel() // x.shape[1]
        mean = x.mean(dims)
        var = x.var(dims, unbiased=False)
        with torch.no_grad():
            rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach())
            rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var.detach() * (n / max(n - 1, 1)))
            if pre + ".num_batches_tracked" in P:
                P[pre + ".num_batches_tracked"] += 1
    else:
        mean, var = rm, rv
    shape = [1, -1] + [1] * (x.dim() - 2)
    xh = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + BN_EPS)
    return xh * P[pre + ".weight"].view(shape) + P[pre + ".bias"].view(shape)


def _lin(P: SD, pre: str, x: Tensor) -> Tensor:
    return x @ P[pre + ".weight"].t() + P[pre + ".bias"]


def _conv(P: SD, pre: str, x: Tensor, stride: int, pad: int) -> Tensor:
    return F.conv2d(x, P[pre + ".weight"], P[pre + ".bias"], stride=stride, padding=pad)


def _convT(P: SD, pre: str, x: Tensor, stride: int, pad: int, opad: int) -> Tensor:
    return F.conv_transpose2d(x, P[pre + ".weight"], P[pre + ".bias"], stride=stride,
                              padding=pad, output_padding=opad)


def _ln(P: SD, pre: str, x: Tensor) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = x.var(-1, unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * P[pre + ".weight"] + P[pre + ".bias"]


def _gelu(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


# Kink hook (tests only).  End-to-end gradients are discontinuous where a pre-activation crosses a LeakyReLU / ReLU /
# |.| kink; two correct fp32 implementations whose forward values differ by one rounding can pick different sides for a
# unit that sits within rounding distance of 0.  KINKS["masks"] = {key: bool tensor} makes the activation with that
# key use the GIVEN derivative side per unit (y = x where mask else slope * x) instead of sign(x), so that gradients
# can be compared at a tight tolerance for the SAME choice of sides; KINKS["log"] = {} records every pre-activation
# so a test can check that the two implementations differ only on units inside the forward-tolerance band.
# `key` is the state_dict prefix of the module that produced the pre-activation ("recon_x" for the |recon| term).
KINKS = {"masks": None, "log": None}


def _act(x: Tensor, slope: float, key: str) -> Tensor:
    if KINKS["log"] is not None:
        KINKS["log"][key] = x.detach()
    M = KINKS["masks"]
    if M is not None and key in M:
        mask = M[key].to(x.device).reshape(x.shape)
        return x * torch.where(mask, torch.ones((), dtype=x.dtype), torch.full((), slope, dtype=x.dtype))
    return F.leaky_relu(x, slope) if slope != 0.0 else F.relu(x)


def _abs(x: Tensor, key: str) -> Tensor:
    return _act(x, -1.0, key) if (KINKS["masks"] is not None and key in KINKS["masks"]) or KINKS["log"] is not None \
        else x.abs()


def reparameterize(mu: Tensor, logvar: Tensor, eps: Tensor) -> Tensor:
    """z = mu + eps * exp(0.5*logvar)  (vessel_analysis/00_core/models.py:252-255)."""
    return mu + eps * torch.exp(0.5 * logvar)


# ----------------------------------------------------------------------------------------
# hybrid-ViT backbone (vessel_analysis/00_core/vit_backbone.py, latent_translator/models.py)
# ----------------------------------------------------------------------------------------
def _mha(P: SD, pre: str, x: Tensor, heads: int) -> Tensor:
    """nn.MultiheadAttention(batch_first=True) self-attention, dropout 0
    (vit_backbone.py:28-30,43): packed in_proj, softmax(QK^T/sqrt(d)) V, out_proj."""
    B, S, D = x.shape
    d = D // heads
    qkv = x @ P[pre + ".in_proj_weight"].t() + P[pre + ".in_proj_bias"]
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, S, heads, d).transpose(1, 2)
    k = k.view(B, S, heads, d).transpose(1, 2)
    v = v.view(B, S, heads, d).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(d), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, S, D)
    return _lin(P, pre + ".out_proj", o)


def vit_block(P: SD, pre: str, x: Tensor, heads: int = 8) -> Tensor:
    """ViTBlock.forward (vit_backbone.py:40-47; the latent_translator variant calls norm1
    three times on the same input, models.py:36 — same value)."""
    x = x + _mha(P, pre + ".attn", _ln(P, pre + ".norm1", x), heads)
    h = _gelu(_lin(P, pre + ".mlp.0", _ln(P, pre + ".norm2", x)))
    return x + _lin(P, pre + ".mlp.3", h)


def vit_stem(P: SD, pre: str, x: Tensor, train: bool) -> Tensor:
    """5 x (Conv3x3 s2 p1 + BN + LeakyReLU(0.01))  (vit_backbone.py:74-90)."""
    for i in range(5):
        x = _conv(P, f"{pre}.{3 * i}", x, 2, 1)
        x = _act(_bn(P, f"{pre}.{3 * i + 1}", x, train), 0.01, f"{pre}.{3 * i + 1}")
    return x


def vit_tokens(P: SD, pre: str, feat: Tensor) -> Tensor:
    """b c h w -> b (h w) c, prepend cls, add pos-emb[:, :n+1]  (vit_backbone.py:164-170)."""
    B, C, H, W = feat.shape
    tok = feat.flatten(2).transpose(1, 2)
    cls = P[pre + ".cls_token"].expand(B, 1, C)
    tok = torch.cat([cls, tok], dim=1)
    return tok + P[pre + ".pos_embedding"][:, : tok.shape[1]]


def vit_encode_cls(P: SD, pre: str, x: Tensor, train: bool, depth: int = 6) -> Tensor:
    tok = vit_tokens(P, pre, vit_stem(P, pre + ".stem", x, train))
    for i in range(depth):
        tok = vit_block(P, f"{pre}.transformer.{i}", tok)
    return _ln(P, pre + ".to_latent", tok[:, 0])


def _resblock(P: SD, pre: str, x: Tensor, train: bool) -> Tensor:
    """x + BN(conv(LReLU0.2(BN(conv x))))  (vit_backbone.py:7-19)."""
    h = _act(_bn(P, pre + ".conv.1", _conv(P, pre + ".conv.0", x, 1, 1), train), 0.2, pre + ".conv.1")
    return x + _bn(P, pre + ".conv.4", _conv(P, pre + ".conv.3", h, 1, 1), train)


def vit_decode(P: SD, pre: str, z: Tensor, grid_hw: Tuple[int, int], train: bool,
               res_after: int = 3) -> Tensor:
    """ViTVAE.decode: decoder_input Linear -> view(B,256,gh,gw) -> 5 up-stages
    (ConvT3x3 s2 p1 op1 + BN + LReLU(0.01)), a ResBlock after the first `res_after` stages
    (3 in vit_backbone.py:124-156, 4 in latent_translator/models.py:85-92), final Conv3x3."""
    h = _lin(P, pre + ".decoder_input", z).view(z.shape[0], 256, grid_hw[0], grid_hw[1])
    i = 0
    for s in range(5):
        h = _convT(P, f"{pre}.decoder.{i}", h, 2, 1, 1)
        h = _act(_bn(P, f"{pre}.decoder.{i + 1}", h, train), 0.01, f"{pre}.decoder.{i + 1}")
        i += 3
        if s < res_after:
            h = _resblock(P, f"{pre}.decoder.{i}", h, train)
            i += 1
    return _conv(P, f"{pre}.decoder.{i}", h, 1, 1)


# ----------------------------------------------------------------------------------------
# CausalViTVAE (vessel_analysis/00_core/models.py:181-307)
# ----------------------------------------------------------------------------------------
def vessel_shapes(H: int, W: int, z_dim=128, m_dim=12, t_dim=19) -> Dict[str, Tuple[int, ...]]:
    """state_dict key -> shape of CausalViTVAE (checked against the reference in make_golden)."""
    s: Dict[str, Tuple[int, ...]] = {}
    gh, gw = H // 32, W // 32

    def bn(p, c):
        s[p + ".weight"] = (c,); s[p + ".bias"] = (c,)
        s[p + ".running_mean"] = (c,); s[p + ".running_var"] = (c,)
        s[p + ".num_batches_tracked"] = ()

    def lin(p, o, i):
        s[p + ".weight"] = (o, i); s[p + ".bias"] = (o,)

    def conv(p, o, i, k=3):
        s[p + ".weight"] = (o, i, k, k); s[p + ".bias"] = (o,)

    def convT(p, i, o, k=3):
        s[p + ".weight"] = (i, o, k, k); s[p + ".bias"] = (o,)

    b = "backbone"
    chans = [1, 32, 64, 128, 256, 256]
    for i in range(5):
        conv(f"{b}.stem.{3 * i}", chans[i + 1], chans[i]); bn(f"{b}.stem.{3 * i + 1}", chans[i + 1])
    s[b + ".pos_embedding"] = (1, gh * gw + 1, 256)
    s[b + ".cls_token"] = (1, 1, 256)
    for i in range(6):
        p = f"{b}.transformer.{i}"
        s[p + ".norm1.weight"] = (256,); s[p + ".norm1.bias"] = (256,)
        s[p + ".norm2.weight"] = (256,); s[p + ".norm2.bias"] = (256,)
        s[p + ".attn.in_proj_weight"] = (768, 256); s[p + ".attn.in_proj_bias"] = (768,)
        lin(p + ".attn.out_proj", 256, 256)
        lin(p + ".mlp.0", 512, 256); lin(p + ".mlp.3", 256, 512)
    s[b + ".to_latent.weight"] = (256,); s[b + ".to_latent.bias"] = (256,)
    lin(b + ".fc_mu", 512, 256); lin(b + ".fc_var", 512, 256)
    lin(b + ".decoder_input", 256 * gh * gw, 512)
    dch = [256, 128, 64, 32, 16, 16]
    i = 0
    for st in range(5):
        convT(f"{b}.decoder.{i}", dch[st], dch[st + 1]); bn(f"{b}.decoder.{i + 1}", dch[st + 1])
        i += 3
        if st < 3:
            c = dch[st + 1]
            conv(f"{b}.decoder.{i}.conv.0", c, c); bn(f"{b}.decoder.{i}.conv.1", c)
            conv(f"{b}.decoder.{i}.conv.3", c, c); bn(f"{b}.decoder.{i}.conv.4", c)
            i += 1
    conv(f"{b}.decoder.{i}", 1, 16)
    lin("enc_adapter.0", 512, 256 + m_dim + t_dim); bn("enc_adapter.1", 512)
    lin("enc_adapter.3", 2 * z_dim, 512)
    lin("dec_adapter.0", 256, z_dim + m_dim); bn("dec_adapter.1", 256)
    lin("dec_adapter.3", 512, 256)
    lin("morph_predictor_shared.0", 64, t_dim); lin("morph_predictor_shared.2", 64, 64)
    lin("morph_predictor_mu", m_dim, 64); lin("morph_predictor_logvar", m_dim, 64)
    return s


def vessel_morph_head(P: SD, t: Tensor):
    """P(M|T) Gaussian head (models.py:243-250,291-295)."""
    h = _act(_lin(P, "morph_predictor_shared.0", t), 0.2, "morph_predictor_shared.0")
    h = _act(_lin(P, "morph_predictor_shared.2", h), 0.2, "morph_predictor_shared.2")
    return _lin(P, "morph_predictor_mu", h), torch.clamp(_lin(P, "morph_predictor_logvar", h), -10, 10)


def vessel_decode(P: SD, m: Tensor, z: Tensor, grid_hw, train: bool) -> Tensor:
    """backbone.decode(dec_adapter(cat[m, z]))  (models.py:299-305;
    generate_counterfactual.py:97-99) — m first."""
    h = _lin(P, "dec_adapter.0", torch.cat([m, z], dim=1))
    h = _act(_bn(P, "dec_adapter.1", h, train), 0.2, "dec_adapter.1")
    return vit_decode(P, "backbone", _lin(P, "dec_adapter.3", h), grid_hw, train)


def vessel_encode(P: SD, x: Tensor, m: Tensor, t: Tensor, train: bool):
    """stem -> tokens -> 6 blocks -> to_latent(CLS) -> enc_adapter -> chunk -> clamps
    (models.py:257-286)."""
    cls = vit_encode_cls(P, "backbone", x, train)
    h = _lin(P, "enc_adapter.0", torch.cat([cls, m, t], dim=1))
    h = _act(_bn(P, "enc_adapter.1", h, train), 0.2, "enc_adapter.1")
    mu, logvar = _lin(P, "enc_adapter.3", h).chunk(2, dim=1)
    return torch.clamp(mu, -100, 100), torch.clamp(logvar, -10, 10)


def vessel_forward(P: SD, x: Tensor, m: Tensor, t: Tensor, eps: Tensor, train: bool):
    """CausalViTVAE.forward (models.py:257-307) -> 6-tuple."""
    mu, logvar = vessel_encode(P, x, m, t, train)
    z = reparameterize(mu, logvar, eps)
    m_mu, m_logvar = vessel_morph_head(P, t)
    recon = vessel_decode(P, m, z, (x.shape[2] // 32, x.shape[3] // 32), train)
    return recon, m_mu, mu, logvar, m_mu, m_logvar


# the CNN variant of the vessel model (vessel_analysis/00_core/models.py:9-166), fixed 768x1280 input
VESSEL_CNN_ENC = (1, 32, 64, 128, 256, 512, 512, 512)
VESSEL_CNN_DEC = (512, 512, 512, 256, 128, 64, 32, 1)


def vessel_cnn_shapes(z_dim=128, m_dim=12, t_dim=19) -> Dict[str, Tuple[int, ...]]:
    """state_dict key -> shape of CausalVesselVAE (models.py:32-145), in the reference's key order."""
    s: Dict[str, Tuple[int, ...]] = {}

    def bn(pre, c):
        s[pre + ".weight"] = (c,); s[pre + ".bias"] = (c,); s[pre + ".running_mean"] = (c,)
        s[pre + ".running_var"] = (c,); s[pre + ".num_batches_tracked"] = ()

    def lin(pre, i, o):
        s[pre + ".weight"] = (o, i); s[pre + ".bias"] = (o,)
    for i, (a, b) in enumerate(zip(VESSEL_CNN_ENC[:-1], VESSEL_CNN_ENC[1:])):
        s[f"enc_conv.{3 * i}.weight"] = (b, a, 4, 4); s[f"enc_conv.{3 * i}.bias"] = (b,)
        bn(f"enc_conv.{3 * i + 1}", b)
    flat = 512 * 6 * 10
    lin("enc_fc.0", flat + m_dim + t_dim, 1024); bn("enc_fc.1", 1024); lin("enc_fc.3", 1024, 2 * z_dim)
    lin("morph_predictor_shared.0", t_dim, 64); lin("morph_predictor_shared.2", 64, 64)
    lin("morph_predictor_mu", 64, m_dim); lin("morph_predictor_logvar", 64, m_dim)
    lin("dec_fc.0", m_dim + z_dim, 1024); bn("dec_fc.1", 1024); lin("dec_fc.3", 1024, flat)
    for i, (a, b) in enumerate(zip(VESSEL_CNN_DEC[:-1], VESSEL_CNN_DEC[1:])):
        s[f"dec_conv.{4 * i + 1}.weight"] = (b, a, 3, 3); s[f"dec_conv.{4 * i + 1}.bias"] = (b,)
        if b != 1:
            bn(f"dec_conv.{4 * i + 2}", b)
    return s


def vessel_cnn_decode(P: SD, m: Tensor, z: Tensor, train: bool) -> Tensor:
    """dec_fc -> view(-1, 512, 6, 10) -> 7 x [nearest x2, Conv3x3, BN, ReLU] (last: Conv3x3, Sigmoid)
    (models.py:63-69,123-145,161-164) — m first."""
    h = _act(_bn(P, "dec_fc.1", _lin(P, "dec_fc.0", torch.cat([m, z], dim=1)), train), 0.2, "dec_fc.1")
    h = _act(_lin(P, "dec_fc.3", h), 0.0, "dec_fc.3").view(-1, 512, 6, 10)
    n = len(VESSEL_CNN_DEC) - 1
    for i in range(n):
        h = _conv(P, f"dec_conv.{4 * i + 1}", F.interpolate(h, scale_factor=2, mode="nearest"), 1, 1)
        h = _act(_bn(P, f"dec_conv.{4 * i + 2}", h, train), 0.0, f"dec_conv.{4 * i + 2}") if i + 1 < n else torch.sigmoid(h)
    return h


def vessel_cnn_forward(P: SD, x: Tensor, m: Tensor, t: Tensor, eps: Tensor, train: bool):
    """CausalVesselVAE.forward (models.py:153-166) -> 6-tuple."""
    h = x
    for i in range(len(VESSEL_CNN_ENC) - 1):
        h = _act(_bn(P, f"enc_conv.{3 * i + 1}", _conv(P, f"enc_conv.{3 * i}", h, 2, 1), train), 0.2, f"enc_conv.{3 * i + 1}")
    h = _lin(P, "enc_fc.0", torch.cat([h.flatten(1), m, t], dim=1))
    mu, logvar = _lin(P, "enc_fc.3", _act(_bn(P, "enc_fc.1", h, train), 0.2, "enc_fc.1")).chunk(2, dim=1)
    mu, logvar = torch.clamp(mu, -100, 100), torch.clamp(logvar, -10, 10)
    z = reparameterize(mu, logvar, eps)
    m_mu, m_logvar = vessel_morph_head(P, t)
    return vessel_cnn_decode(P, m, z, train), m_mu, mu, logvar, m_mu, m_logvar


def vessel_cnn_loss_and_grads(P: SD, x, m, t, eps, beta=0.5):
    """One training-mode forward + the vessel loss (train.py:18-60,82) + gradients of every parameter."""
    params = trainable(P)
    for v in params.values():
        v.requires_grad_(True)
    outs = vessel_cnn_forward(P, x, m, t, eps, True)
    recon, kld, morph, sp = vessel_loss(outs[0], x, outs[1], m, outs[2], outs[3], outs[4], outs[5])
    loss = vessel_total(recon, kld, morph, sp, beta)
    grads = dict(zip(params, torch.autograd.grad(loss, list(params.values()))))
    for v in params.values():
        v.requires_grad_(False)
    return outs, {"loss": loss.detach(), "recon": recon.detach(), "kld": kld.detach(), "morph": morph.detach(),
                  "sparsity": sp.detach()}, grads


def vessel_loss(recon_x, x, m_hat, m, mu, logvar, m_mu, m_logvar):
    """loss_function (vessel_analysis/01_train/train.py:18-60): weighted MSE with a
    batch-global pos_weight (no grad), background sparsity L1, KL, Gaussian NLL."""
    mse = (recon_x - x) ** 2
    with torch.no_grad():
        pos_fraction = x.sum() / (x.numel() + 1e-6)
        pos_weight = torch.clamp((1.0 - pos_fraction) / (pos_fraction + 1e-6), 1.0, 50.0)
    recon = torch.sum(mse * (1.0 + (pos_weight - 1.0) * x))
    sparsity = torch.sum(_abs(recon_x, "recon_x") * (x < 0.1).to(recon_x.dtype))
    kld = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())
    morph = 0.5 * torch.sum(m_logvar + (m - m_mu) ** 2 / torch.exp(m_logvar))
    return recon, kld, morph, sparsity


def vessel_total(recon, kld, morph, sparsity, beta=0.5, lambda_morph=1.0):
    """train.py:82 (lambda_morph=1) / train_kfold.py:71 (lambda_morph=CONFIG['LAMBDA_MORPH'])."""
    return recon + beta * kld + lambda_morph * morph + 0.3 * sparsity


# ----------------------------------------------------------------------------------------
# latent_translator ViTVAE (latent_translator/models.py:40-126, engine.py:19-30)
# ----------------------------------------------------------------------------------------
def lt_shapes(H: int, W: int, latent=512) -> Dict[str, Tuple[int, ...]]:
    v = vessel_shapes(H, W)
    s = {k[len("backbone."):]: shp for k, shp in v.items()
         if k.startswith("backbone.") and ".decoder." not in k}
    s["fc_mu.weight"] = (latent, 256); s["fc_mu.bias"] = (latent,)
    s["fc_var.weight"] = (latent, 256); s["fc_var.bias"] = (latent,)
    s["decoder_input.weight"] = (256 * (H // 32) * (W // 32), latent)
    dch = [256, 128, 64, 32, 16, 16]
    i = 0
    for st in range(5):
        p = f"decoder.{i}"
        s[p + ".weight"] = (dch[st], dch[st + 1], 3, 3); s[p + ".bias"] = (dch[st + 1],)
        for leaf, shp in (("weight", (dch[st + 1],)), ("bias", (dch[st + 1],)),
                          ("running_mean", (dch[st + 1],)), ("running_var", (dch[st + 1],)),
                          ("num_batches_tracked", ())):
            s[f"decoder.{i + 1}.{leaf}"] = shp
        i += 3
        if st < 4:
            c = dch[st + 1]
            for j, bnj in ((0, 1), (3, 4)):
                s[f"decoder.{i}.conv.{j}.weight"] = (c, c, 3, 3); s[f"decoder.{i}.conv.{j}.bias"] = (c,)
                for leaf, shp in (("weight", (c,)), ("bias", (c,)), ("running_mean", (c,)),
                                  ("running_var", (c,)), ("num_batches_tracked", ())):
                    s[f"decoder.{i}.conv.{bnj}.{leaf}"] = shp
            i += 1
    s[f"decoder.{i}.weight"] = (1, 16, 3, 3); s[f"decoder.{i}.bias"] = (1,)
    return s


def _prefixed(P: SD, pre: str) -> SD:
    class _V(dict):
        def __init__(self, base, pre):
            self.base, self.pre = base, pre
        def __getitem__(self, k):
            return self.base[k[len(self.pre) + 1:]]
        def __contains__(self, k):
            return k[len(self.pre) + 1:] in self.base
        def __setitem__(self, k, v):
            self.base[k[len(self.pre) + 1:]] = v
    return _V(P, pre)


def lt_encode(P: SD, x: Tensor, train: bool):
    """ViTVAE.encode (latent_translator/models.py:95-111)."""
    Q = _prefixed(P, "b")
    cls = vit_encode_cls(Q, "b", x, train)
    return _lin(P, "fc_mu", cls), _lin(P, "fc_var", cls)


def lt_forward(P: SD, x: Tensor, eps: Tensor, train: bool):
    """ViTVAE.forward (latent_translator/models.py:121-126) -> (recons, input, mu, log_var)."""
    mu, logvar = lt_encode(P, x, train)
    z = reparameterize(mu, logvar, eps)
    rec = vit_decode(_prefixed(P, "b"), "b", z, (x.shape[2] // 32, x.shape[3] // 32), train, res_after=4)
    return rec, x, mu, logvar


def lt_loss(recons, x, mu, logvar, beta=1.0):
    """engine.py:25-27 — mean-reduced MSE + beta * mean-reduced KL."""
    rl = torch.mean((recons - x) ** 2)
    kl = -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())
    return rl + beta * kl, rl, kl


# ----------------------------------------------------------------------------------------
# CausalBioVAE (causal_cascade/models.py:5-89, train.py:5-17)
# ----------------------------------------------------------------------------------------
def cascade_shapes(m_dim=8, t_dim=19, latent=64, img_channels=1):
    s = {}
    ch = [img_channels, 32, 64, 128, 256]
    for i in range(4):
        s[f"enc_conv.{2 * i}.weight"] = (ch[i + 1], ch[i], 4, 4); s[f"enc_conv.{2 * i}.bias"] = (ch[i + 1],)
    for p, o, i in (("enc_fc.0", 512, 4096 + m_dim + t_dim), ("enc_fc.2", 256, 512),
                    ("fc_mu", latent, 256), ("fc_logvar", latent, 256),
                    ("mechanism_net.0", 64, t_dim), ("mechanism_net.3", 64, 64),
                    ("mechanism_net.5", m_dim, 64), ("dec_input", 4096, latent + m_dim)):
        s[p + ".weight"] = (o, i); s[p + ".bias"] = (o,)
    for leaf, shp in (("weight", (64,)), ("bias", (64,)), ("running_mean", (64,)),
                      ("running_var", (64,)), ("num_batches_tracked", ())):
        s["mechanism_net.1." + leaf] = shp
    dch = [256, 128, 64, 32, img_channels]
    for i in range(4):
        s[f"dec_conv.{2 * i}.weight"] = (dch[i], dch[i + 1], 4, 4); s[f"dec_conv.{2 * i}.bias"] = (dch[i + 1],)
    return s


def cascade_forward(P: SD, x: Tensor, m: Tensor, t_idx: Tensor, eps: Tensor, train: bool):
    """CausalBioVAE.forward (causal_cascade/models.py:70-89).  AdaptiveAvgPool2d((4,4)) and the
    final bilinear resize are exact identities when the input is 64x64; other
    sizes go through F.adaptive_avg_pool2d / F.interpolate like the reference."""
    t_dim = P["mechanism_net.0.weight"].shape[1]
    t1 = F.one_hot(t_idx, num_classes=t_dim).to(x.dtype)
    h = x
    for i in range(4):
        h = _act(_conv(P, f"enc_conv.{2 * i}", h, 2, 1), 0.0, f"enc_conv.{2 * i}")
    h = F.adaptive_avg_pool2d(h, (4, 4)).flatten(1)
    h = _act(_lin(P, "enc_fc.0", torch.cat([h, m, t1], dim=1)), 0.0, "enc_fc.0")
    h = _act(_lin(P, "enc_fc.2", h), 0.0, "enc_fc.2")
    mu, logvar = _lin(P, "fc_mu", h), _lin(P, "fc_logvar", h)
    z = reparameterize(mu, logvar, eps)
    g = _act(_bn(P, "mechanism_net.1", _lin(P, "mechanism_net.0", t1), train), 0.0, "mechanism_net.1")
    g = _act(_lin(P, "mechanism_net.3", g), 0.0, "mechanism_net.3")
    m_hat = _lin(P, "mechanism_net.5", g)
    d = _lin(P, "dec_input", torch.cat([z, m_hat], dim=1)).view(-1, 256, 4, 4)
    for i in range(4):
        d = _convT(P, f"dec_conv.{2 * i}", d, 2, 1, 0)
        if i < 3:
            d = _act(d, 0.0, f"dec_conv.{2 * i}")
    if d.shape[2:] != x.shape[2:]:
        d = F.interpolate(d, size=x.shape[2:], mode="bilinear", align_corners=False)
    return d, m_hat, mu, logvar


def cascade_loss(recon_x, x, m_hat, m, mu, logvar, gamma=2000.0):
    """causal_cascade/train.py:5-17."""
    rl = torch.sum((recon_x - x) ** 2)
    ml = torch.sum((m_hat - m) ** 2)
    kld = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())
    return rl + gamma * ml + kld, rl, ml


# ----------------------------------------------------------------------------------------
# CausalMorphVAE12 (mnist_test/01_baseline_causal_vae/models.py:6-72; 06 variant :34-85)
# ----------------------------------------------------------------------------------------
def mnist_shapes(m_dim=12, t_dim=10, z_dim=10, variant="01"):
    s = {"enc_conv.0.weight": (32, 1, 4, 4), "enc_conv.0.bias": (32,),
         "enc_conv.2.weight": (64, 32, 4, 4), "enc_conv.2.bias": (64,),
         "enc_fc.0.weight": (512, 3136 + m_dim + t_dim), "enc_fc.0.bias": (512,),
         "enc_fc.2.weight": (2 * z_dim, 512), "enc_fc.2.bias": (2 * z_dim,),
         "dec_fc.0.weight": (3136, m_dim + z_dim), "dec_fc.0.bias": (3136,),
         "dec_conv.0.weight": (64, 32, 4, 4), "dec_conv.0.bias": (32,),
         "dec_conv.2.weight": (32, 1, 4, 4), "dec_conv.2.bias": (1,)}
    if variant == "01":
        s.update({"morph_predictor.0.weight": (128, t_dim), "morph_predictor.0.bias": (128,),
                  "morph_predictor.2.weight": (m_dim, 128), "morph_predictor.2.bias": (m_dim,)})
    else:
        s.update({"morph_predictor_shared.0.weight": (128, t_dim), "morph_predictor_shared.0.bias": (128,),
                  "morph_predictor_mu.weight": (m_dim, 128), "morph_predictor_mu.bias": (m_dim,),
                  "morph_predictor_logvar.weight": (m_dim, 128), "morph_predictor_logvar.bias": (m_dim,)})
    return s


def disc_shapes(t_dim=10, z_dim=10):
    return {"net.0.weight": (64, z_dim), "net.0.bias": (64,), "net.2.weight": (64, 64),
            "net.2.bias": (64,), "net.4.weight": (t_dim, 64), "net.4.bias": (t_dim,)}


def mnist_decode(P: SD, m: Tensor, z: Tensor) -> Tensor:
    """dec_fc -> view(64,7,7) -> ConvT4x4 s2 p1 + ReLU -> ConvT + Sigmoid (models.py:66-70)."""
    h = F.relu(_lin(P, "dec_fc.0", torch.cat([m, z], dim=1))).view(-1, 64, 7, 7)
    h = F.relu(_convT(P, "dec_conv.0", h, 2, 1, 0))
    return torch.sigmoid(_convT(P, "dec_conv.2", h, 2, 1, 0))


def mnist_forward(P: SD, x: Tensor, m: Tensor, t: Tensor, eps: Tensor, variant="01"):
    """CausalMorphVAE12.forward: 01 -> 4-tuple decoding from m_hat (models.py:55-72);
    06 -> 6-tuple decoding from the real m (06_model_experiment/models.py:62-85)."""
    h = F.relu(_conv(P, "enc_conv.0", x, 2, 1))
    h = F.relu(_conv(P, "enc_conv.2", h, 2, 1)).flatten(1)
    h = F.relu(_lin(P, "enc_fc.0", torch.cat([h, m, t], dim=1)))
    mu, logvar = _lin(P, "enc_fc.2", h).chunk(2, dim=1)
    z = reparameterize(mu, logvar, eps)
    if variant == "01":
        m_hat = _lin(P, "morph_predictor.2", F.relu(_lin(P, "morph_predictor.0", t)))
        return mnist_decode(P, m_hat, z), m_hat, mu, logvar
    g = F.relu(_lin(P, "morph_predictor_shared.0", t))
    m_mu, m_logvar = _lin(P, "morph_predictor_mu", g), _lin(P, "morph_predictor_logvar", g)
    return mnist_decode(P, m, z), m_mu, mu, logvar, m_mu, m_logvar


def disc_forward(P: SD, z: Tensor) -> Tensor:
    """LatentDiscriminator (mnist_test/01_baseline_causal_vae/models.py:93-111)."""
    h = F.leaky_relu(_lin(P, "net.0", z), 0.2)
    h = F.leaky_relu(_lin(P, "net.2", h), 0.2)
    return _lin(P, "net.4", h)


def bce_sum(p: Tensor, y: Tensor) -> Tensor:
    """F.binary_cross_entropy(reduction='sum'): each log clamped at -100 (torch semantics)."""
    return -(y * torch.clamp(torch.log(p), min=-100) + (1 - y) * torch.clamp(torch.log(1 - p), min=-100)).sum()


def mnist_vae_loss(P: SD, D: SD, x, m, t, eps, eps_adv, beta=1.0, lambda_adv=10.0, variant="01"):
    """VAE half of the adversarial step (mnist_test/01_baseline_causal_vae/train.py:65-87;
    06 variant train.py:67-94): BCE_sum + beta*KL + morph + 100*lambda_adv*KL(U || softmax D(z'))."""
    out = mnist_forward(P, x, m, t, eps, variant)
    recon, m_hat, mu, logvar = out[:4]
    l_rec = bce_sum(recon.reshape(-1, 784), x.reshape(-1, 784))
    l_kld = beta * (-0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp()))
    if variant == "01":
        l_m = 100.0 * torch.sum((m_hat - m) ** 2)
    else:
        m_mu, m_logvar = out[4], out[5]
        l_m = 0.5 * torch.sum(m_logvar + (m - m_mu) ** 2 / m_logvar.exp())
    logits = disc_forward(D, reparameterize(mu, logvar, eps_adv))
    T = logits.shape[1]
    logp = F.log_softmax(logits, dim=1)
    u = 1.0 / T
    l_adv = (u * (math.log(u) - logp)).sum() / logits.shape[0] * lambda_adv * 100
    return l_rec + l_kld + l_m + l_adv, l_rec, l_kld, l_m, l_adv


def mnist_disc_loss(P: SD, D: SD, x, m, t, eps, variant="01"):
    """Discriminator half (train.py:41-60): z from a no-grad VAE pass, CE(D(z), argmax t)."""
    with torch.no_grad():
        out = mnist_forward(P, x, m, t, eps, variant)
        z = reparameterize(out[2], out[3], eps)
    logits = disc_forward(D, z)
    return F.cross_entropy(logits, t.argmax(1))


# ----------------------------------------------------------------------------------------
# optimizer step (vessel_analysis/01_train/train.py:84-86,152)
# ----------------------------------------------------------------------------------------
def clip_grad_norm(grads: Dict[str, Tensor], max_norm: float):
    """torch.nn.utils.clip_grad_norm_: total L2 norm; coef = min(1, max_norm/(norm+1e-6))."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).to(next(iter(grads.values())).dtype)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return {k: g * coef for k, g in grads.items()}, total


def adam_step(params: SD, grads: Dict[str, Tensor], state: Dict[str, Dict[str, Tensor]], step: int,
              lr: float, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (no weight decay, no amsgrad) — in place on `params`."""
    for k, g in grads.items():
        st = state.setdefault(k, {"m": torch.zeros_like(g), "v": torch.zeros_like(g)})
        st["m"].mul_(b1).add_(g, alpha=1 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        denom = st["v"].sqrt() / math.sqrt(bc2) + eps
        params[k].data.addcdiv_(st["m"], denom, value=-lr / bc1)


def trainable(P: SD) -> SD:
    return {k: v for k, v in P.items() if v.is_floating_point()
            and not k.endswith(("running_mean", "running_var"))}


def vessel_train_step(P: SD, state, step: int, x, m, t, eps, lr=1e-4, beta=0.5, max_norm=5.0,
                      lambda_morph=1.0):
    """One reference training step (train.py:77-86): fwd, loss, bwd, clip 5.0, Adam(lr).
    Returns (loss scalars dict, grads before clipping, total grad norm)."""
    W = trainable(P)
    for v in W.values():
        v.requires_grad_(True)
        v.grad = None
    out = vessel_forward(P, x, m, t, eps, train=True)
    recon, kld, morph, sp = vessel_loss(out[0], x, out[1], m, out[2], out[3], out[4], out[5])
    loss = vessel_total(recon, kld, morph, sp, beta, lambda_morph)
    names = [k for k in W if not k.startswith(("backbone.fc_mu", "backbone.fc_var"))]
    gl = torch.autograd.grad(loss, [W[k] for k in names])
    grads = dict(zip(names, gl))
    for v in W.values():
        v.requires_grad_(False)
    clipped, total = clip_grad_norm(grads, max_norm)
    adam_step(P, clipped, state, step, lr)
    return ({"loss": loss.detach(), "recon": recon.detach(), "kld": kld.detach(),
             "morph": morph.detach(), "sparsity": sp.detach()}, grads, total)


def counterfactual_do(m: Tensor, k: int, value=None, delta=None) -> Tensor:
    """do(M_k := v) or do(M_k := M_k + delta) on a cloned m
    (generate_counterfactual.py:86-88; analyze_vessel.py:101-104)."""
    mp = m.clone()
    if value is not None:
        mp[:, k] = value
    if delta is not None:
        mp[:, k] = mp[:, k] + delta
    return mp
