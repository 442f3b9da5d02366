"""Per-op parity: native sm_100a kernels (through the C ABI / drop-in modules) vs the oracle's
building blocks evaluated in fp64 on the CPU.  Tolerances: 1e-5 relative on forward values,
1e-4 on gradients (relative to each tensor's max |value|), as stated by BASELINE north_star."""
import math

import pytest
import torch

from oracle import cvae_oracle as O

pytestmark = pytest.mark.gpu

FWD_TOL, GRAD_TOL = 1e-5, 1e-4


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def assert_close(a, b, tol, what):
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    e = rel(a, b)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol}"


def gen(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def run_pair(mod_gpu, sd, ref_fn, x, train=True, tol_f=FWD_TOL, tol_g=GRAD_TOL, need_dx=True):
    """Load `sd` into the native module, run fwd/bwd on GPU and `ref_fn(P, x)` on the CPU in fp64
    (ground truth) and fp32 (the reference's own precision).  A gradient passes when its error
    against fp64 is within tol_g of the tensor's max |g|, or within 4x the fp32 reference's own
    error (zero-true-gradient tensors such as a conv bias feeding BatchNorm are pure rounding noise
    in the reference)."""
    mod_gpu.load_state_dict(sd)
    mod_gpu = mod_gpu.cuda()
    mod_gpu.train(train)
    xg = x.cuda().requires_grad_(need_dx)
    y = mod_gpu(xg)
    gy = gen(*y.shape, seed=99)
    y.backward(gy.cuda())

    def ref(dt):
        P = {k: (v.to(dt).clone() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        W = O.trainable(P)
        for v in W.values():
            v.requires_grad_(True)
        xr = x.to(dt).requires_grad_(need_dx)
        yr = ref_fn(P, xr)
        yr.backward(gy.to(dt))
        return P, W, xr, yr
    P, W, xr, yr = ref(torch.float64)
    _, W32, xr32, _ = ref(torch.float32)
    bad = []
    if rel(y, yr) > tol_f:
        bad.append(("forward", rel(y, yr), 0.0))

    def check(name, got, want, want32):
        want = want.detach().double().cpu()
        err = (got.detach().double().cpu() - want).abs().max().item()
        noise = (want32.detach().double() - want).abs().max().item()
        if err > max(tol_g * want.abs().max().item(), 4 * noise):
            bad.append((name, err / max(want.abs().max().item(), 1e-30), noise / max(want.abs().max().item(), 1e-30)))
    if need_dx:
        check("grad input", xg.grad, xr.grad, xr32.grad)
    for k, p in mod_gpu.named_parameters():
        assert p.grad is not None, f"no grad for {k}"
        check(f"grad {k}", p.grad, W[k].grad, W32[k].grad)
    if train:
        for k, v in mod_gpu.state_dict().items():
            if k.endswith(("running_mean", "running_var")) and rel(v, P[k]) > 1e-5:
                bad.append((k, rel(v, P[k]), 0.0))
            if k.endswith("num_batches_tracked"):
                assert int(v) == int(P[k]), k
    assert not bad, "mismatches (name, rel err, fp32-reference rel noise): " + "; ".join(
        f"{n}: {e:.2e} (noise {z:.1e})" for n, e, z in bad)
    return y, yr


CONV_CASES = [
    # Cin, Cout, k, s, p, H, W, B
    (1, 32, 3, 2, 1, 32, 32, 3),      # stem.0 (thread-per-pixel kernel, Cin = 1)
    (32, 64, 3, 2, 1, 16, 16, 3),
    (64, 128, 3, 2, 1, 10, 14, 2),    # ragged tile sizes
    (16, 1, 3, 1, 1, 20, 24, 2),      # image head (Cout = 1)
    (16, 16, 3, 1, 1, 12, 12, 2),
    (32, 64, 4, 2, 1, 14, 14, 3),     # mnist / cascade 4x4 s2
    (1, 32, 4, 2, 1, 28, 28, 2),
    (128, 256, 3, 2, 1, 8, 8, 5),
    (16, 16, 3, 2, 1, 256, 256, 4),   # image-sized 16 -> 16 stride 2: fp32 tile kernel (conv_few.cu)
    (16, 16, 3, 2, 1, 200, 144, 10),  # the same with ragged tiles
    (1, 32, 4, 2, 1, 64, 64, 70),     # causal_cascade enc_conv.0 at a training-sized batch: 4x4 tile kernels (fwd, wgrad)
    (1, 32, 4, 2, 1, 50, 44, 130),    # the same with ragged tiles
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d(case):
    from causal_vae_b200 import nn
    Cin, Cout, k, s, p, H, W, B = case
    sd = O.fill_state_dict({"weight": (Cout, Cin, k, k), "bias": (Cout,)}, seed=1)
    x = gen(B, Cin, H, W, seed=2)
    run_pair(nn.Conv2d(Cin, Cout, k, s, p), sd, lambda P, xx: O._conv({"c.weight": P["weight"], "c.bias": P["bias"]}, "c", xx, s, p), x)


CONVT_CASES = [
    # Cin, Cout, k, s, p, op, H, W, B
    (256, 128, 3, 2, 1, 1, 4, 4, 3),
    (32, 16, 3, 2, 1, 1, 16, 12, 2),
    (16, 16, 3, 2, 1, 1, 16, 16, 2),
    (16, 16, 3, 2, 1, 1, 128, 128, 4),   # decoder.12 at image size: fp32 tile kernels (conv_few.cu), fwd + input gradient
    (16, 16, 3, 2, 1, 1, 100, 72, 10),   # the same with ragged tiles
    (64, 32, 4, 2, 1, 0, 7, 7, 3),    # mnist dec_conv.0
    (32, 1, 4, 2, 1, 0, 14, 14, 3),   # mnist dec_conv.2 (Cout = 1)
    (32, 1, 4, 2, 1, 0, 32, 32, 70),  # causal_cascade dec_conv.6 at a training-sized batch: 4x4 tile kernels (input gradient, wgrad)
    (32, 1, 4, 2, 1, 0, 25, 22, 130), # the same with ragged tiles
    (16, 1, 4, 2, 1, 0, 12, 12, 2),   # four-phase ConvT(C -> 1) forward kernel with 4 / 16 threads per position
    (64, 1, 4, 2, 1, 0, 9, 7, 2),
]


@pytest.mark.parametrize("case", CONVT_CASES)
def test_conv_transpose2d(case):
    from causal_vae_b200 import nn
    Cin, Cout, k, s, p, op, H, W, B = case
    sd = O.fill_state_dict({"weight": (Cin, Cout, k, k), "bias": (Cout,)}, seed=3)
    x = gen(B, Cin, H, W, seed=4)
    run_pair(nn.ConvTranspose2d(Cin, Cout, k, s, p, op), sd,
             lambda P, xx: O._convT({"c.weight": P["weight"], "c.bias": P["bias"]}, "c", xx, s, p, op), x)


@pytest.mark.parametrize("shape", [(64, 287, 512), (7, 19, 64), (130, 256, 768), (5, 512, 1024), (64, 64, 12),
                                   # causal_cascade trains at batch 256: row blocks of the few-row kernels (linear_small.cu)
                                   (256, 4124, 512), (256, 64, 12), (200, 76, 4096), (300, 512, 256), (129, 20, 64)])
def test_linear(shape):
    from causal_vae_b200 import nn
    B, K, N = shape
    sd = O.fill_state_dict({"weight": (N, K), "bias": (N,)}, seed=5)
    x = gen(B, K, seed=6)
    run_pair(nn.Linear(K, N), sd, lambda P, xx: O._lin({"l.weight": P["weight"], "l.bias": P["bias"]}, "l", xx), x)


def test_stem_like_chain_train():
    """Conv-BN-LReLU x3 in training mode: fused statistics epilogue + deferred normalise."""
    from causal_vae_b200 import nn
    seq = nn.Sequential(nn.Conv2d(1, 32, 3, 2, 1), nn.BatchNorm2d(32), nn.LeakyReLU(),
                        nn.Conv2d(32, 64, 3, 2, 1), nn.BatchNorm2d(64), nn.LeakyReLU(),
                        nn.Conv2d(64, 128, 3, 2, 1), nn.BatchNorm2d(128), nn.LeakyReLU())
    sd = O.fill_state_dict({k: tuple(v.shape) for k, v in seq.state_dict().items()}, seed=7)
    x = (gen(4, 1, 32, 32, seed=8) > 0.8).float()

    def ref(P, xx):
        h = xx
        for i in range(3):
            h = torch.nn.functional.leaky_relu(O._bn(P, f"{3 * i + 1}", O._conv(P, f"{3 * i}", h, 2, 1), True), 0.01)
        return h
    run_pair(seq, sd, ref, x, need_dx=False)


def test_decoder_like_chain_train():
    """ConvT-BN-LReLU + ResBlock + ConvT-BN-LReLU + Conv head."""
    from causal_vae_b200 import nn
    seq = nn.Sequential(nn.ConvTranspose2d(64, 32, 3, 2, 1, 1), nn.BatchNorm2d(32), nn.LeakyReLU(), nn.ResBlock(32),
                        nn.ConvTranspose2d(32, 16, 3, 2, 1, 1), nn.BatchNorm2d(16), nn.LeakyReLU(),
                        nn.Conv2d(16, 1, 3, padding=1))
    sd = O.fill_state_dict({k: tuple(v.shape) for k, v in seq.state_dict().items()}, seed=9)
    x = gen(3, 64, 6, 6, seed=10)

    def ref(P, xx):
        lr = torch.nn.functional.leaky_relu
        h = lr(O._bn(P, "1", O._convT(P, "0", xx, 2, 1, 1), True), 0.01)
        h = O._resblock(P, "3", h, True)
        h = lr(O._bn(P, "5", O._convT(P, "4", h, 2, 1, 1), True), 0.01)
        return O._conv(P, "7", h, 1, 1)
    run_pair(seq, sd, ref, x)


def test_image_sized_16_channel_tail_train():
    """ConvT(32->16)-BN-LReLU -> ConvT(16->16)-BN-LReLU -> Conv(16->1) at a size where the 16 -> 16 layer runs
    on the few-channel tile kernels: BatchNorm + LeakyReLU applied while staging, statistics epilogue
    (forward), activation-derivative + BN-backward sums epilogue (input gradient)."""
    from causal_vae_b200 import nn
    seq = nn.Sequential(nn.ConvTranspose2d(32, 16, 3, 2, 1, 1), nn.BatchNorm2d(16), nn.LeakyReLU(),
                        nn.ConvTranspose2d(16, 16, 3, 2, 1, 1), nn.BatchNorm2d(16), nn.LeakyReLU(),
                        nn.Conv2d(16, 1, 3, padding=1))
    sd = O.fill_state_dict({k: tuple(v.shape) for k, v in seq.state_dict().items()}, seed=31)
    # 4.2 M activations per BatchNorm: with generic affine parameters one or two pre-activations land
    # within fp32 rounding of the LeakyReLU kink and flip a derivative, which moves
    # sum-type gradients by ~1e-3 whatever kernel computes them.  Keep every pre-activation at least 2
    # away from the kink - gamma in [0.5, 1], beta = +-8 alternating, so both branches are exercised:
    for bn in ("1", "4"):
        sd[f"{bn}.weight"] = torch.linspace(0.5, 1.0, 16)
        sd[f"{bn}.bias"] = torch.tensor([8.0, -8.0] * 8)
    x = gen(16, 32, 32, 32, seed=32)

    def ref(P, xx):
        lr = torch.nn.functional.leaky_relu
        h = lr(O._bn(P, "1", O._convT(P, "0", xx, 2, 1, 1), True), 0.01)
        h = lr(O._bn(P, "4", O._convT(P, "3", h, 2, 1, 1), True), 0.01)
        return O._conv(P, "6", h, 1, 1)
    run_pair(seq, sd, ref, x)


def test_chain_eval_mode_forward():
    from causal_vae_b200 import nn
    seq = nn.Sequential(nn.ConvTranspose2d(32, 16, 3, 2, 1, 1), nn.BatchNorm2d(16), nn.LeakyReLU(), nn.ResBlock(16),
                        nn.Conv2d(16, 1, 3, padding=1))
    sd = O.fill_state_dict({k: tuple(v.shape) for k, v in seq.state_dict().items()}, seed=11)
    seq.load_state_dict(sd)
    seq = seq.cuda().eval()
    x = gen(2, 32, 5, 7, seed=12)
    with torch.no_grad():
        y = seq(x.cuda())
    P = {k: v.double() if v.is_floating_point() else v for k, v in sd.items()}
    lr = torch.nn.functional.leaky_relu
    h = lr(O._bn(P, "1", O._convT(P, "0", x.double(), 2, 1, 1), False), 0.01)
    h = O._resblock(P, "3", h, False)
    assert_close(y, O._conv(P, "4", h, 1, 1), FWD_TOL, "eval chain")


def test_adapter_chain_train():
    """Linear -> BatchNorm1d -> LeakyReLU(0.2) -> Linear on an unaligned (287-wide) input."""
    from causal_vae_b200 import nn
    seq = nn.Sequential(nn.Linear(287, 512), nn.BatchNorm1d(512), nn.LeakyReLU(0.2), nn.Linear(512, 256))
    sd = O.fill_state_dict({k: tuple(v.shape) for k, v in seq.state_dict().items()}, seed=13)
    x = gen(16, 287, seed=14)

    def ref(P, xx):
        h = torch.nn.functional.leaky_relu(O._bn(P, "1", O._lin(P, "0", xx), True), 0.2)
        return O._lin(P, "3", h)
    run_pair(seq, sd, ref, x)


def test_mlp_relu_sigmoid_chains():
    from causal_vae_b200 import nn
    seq = nn.Sequential(nn.ConvTranspose2d(64, 32, 4, 2, 1), nn.ReLU(), nn.ConvTranspose2d(32, 1, 4, 2, 1), nn.Sigmoid())
    sd = O.fill_state_dict({k: tuple(v.shape) for k, v in seq.state_dict().items()}, seed=15)
    x = gen(3, 64, 7, 7, seed=16)

    def ref(P, xx):
        h = torch.relu(O._convT(P, "0", xx, 2, 1, 0))
        return torch.sigmoid(O._convT(P, "2", h, 2, 1, 0))
    run_pair(seq, sd, ref, x)


def test_layernorm_and_strided_rows():
    from causal_vae_b200 import nn
    ln = nn.LayerNorm(256)
    sd = O.fill_state_dict({"weight": (256,), "bias": (256,)}, seed=17)
    x = gen(6, 9, 256, seed=18)
    run_pair(ln, sd, lambda P, xx: O._ln({"n.weight": P["weight"], "n.bias": P["bias"]}, "n", xx), x)
    # CLS-row slice (row stride S*D)
    ln.load_state_dict(sd)
    ln = ln.cuda()
    tok = x.cuda()
    y = ln(tok[:, 0])
    yr = O._ln({"n.weight": sd["weight"].double(), "n.bias": sd["bias"].double()}, "n", x.double()[:, 0])
    assert_close(y, yr, FWD_TOL, "cls layernorm")


@pytest.mark.parametrize("S", [5, 17, 65])
def test_vit_block(S):
    from causal_vae_b200.vessel.vit_backbone import ViTBlock
    blk = ViTBlock(256, 8, 512)
    for mod in blk.modules():
        if hasattr(mod, "p") and isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    blk.attn.dropout = 0.0
    sd = O.fill_state_dict({k: tuple(v.shape) for k, v in blk.state_dict().items()}, seed=19)
    x = gen(3, S, 256, seed=20)
    run_pair(blk, sd, lambda P, xx: O.vit_block({"b." + k: v for k, v in P.items()}, "b", xx), x, tol_g=2e-4)


def test_tokens_latent_losses():
    from causal_vae_b200 import functional as F
    B, h, w, D, Z = 3, 2, 3, 256, 128
    feat = gen(B, h, w, D, seed=21)
    cls, pos = gen(1, 1, D, seed=22), gen(1, h * w + 1, D, seed=23)
    fg, cg, pg = (t.cuda().requires_grad_(True) for t in (feat, cls, pos))
    tok = F.tokens(fg, cg, pg)
    fr, cr, pr = (t.double().requires_grad_(True) for t in (feat, cls, pos))
    tr = O.vit_tokens({"b.cls_token": cr, "b.pos_embedding": pr}, "b", fr.permute(0, 3, 1, 2))
    assert_close(tok, tr, FWD_TOL, "tokens")
    g = gen(*tr.shape, seed=24)
    tok.backward(g.cuda()); tr.backward(g.double())
    assert_close(fg.grad, fr.grad, GRAD_TOL, "dfeat"); assert_close(cg.grad, cr.grad, GRAD_TOL, "dcls")
    assert_close(pg.grad, pr.grad, GRAD_TOL, "dpos")

    hh = gen(B, 2 * Z, seed=25, scale=4.0)
    hh[0, 3] = 150.0; hh[1, Z + 5] = 12.0; hh[2, Z + 7] = -11.0     # hit every clamp
    eps = gen(B, Z, seed=26)
    hg = hh.cuda().requires_grad_(True)
    mu, lv, z = F.latent(hg, eps.cuda(), 100.0, 10.0)
    hr = hh.double().requires_grad_(True)
    mur, lvr = hr.chunk(2, dim=1)
    mur, lvr = torch.clamp(mur, -100, 100), torch.clamp(lvr, -10, 10)
    zr = O.reparameterize(mur, lvr, eps.double())
    for a, b, n in ((mu, mur, "mu"), (lv, lvr, "logvar"), (z, zr, "z")):
        assert_close(a, b, FWD_TOL, n)
    kl = F.kld_loss(mu, lv)
    klr = -0.5 * torch.sum(1 + lvr - mur.pow(2) - lvr.exp())
    assert_close(kl, klr, FWD_TOL, "kld")
    gz = gen(B, Z, seed=27)
    (kl * 0.5 + (z * gz.cuda()).sum()).backward()
    (klr * 0.5 + (zr * gz.double()).sum()).backward()
    assert_close(hg.grad, hr.grad, GRAD_TOL, "dh latent")

    # vessel reconstruction / sparsity / gaussian nll
    x = (gen(2, 1, 32, 48, seed=28) > 0.8).float()
    r = gen(2, 1, 32, 48, seed=29)
    r[0, 0, 0, :5] = 0.0
    m, mm, ml = gen(2, 12, seed=30), gen(2, 12, seed=31), gen(2, 12, seed=32)
    rg, mmg, mlg = (t.cuda().requires_grad_(True) for t in (r, mm, ml))
    rec, sp = F.vessel_recon_loss(rg, x.cuda())
    nll = F.gauss_nll_loss(m.cuda(), mmg, mlg)
    rr, mmr, mlr = (t.double().requires_grad_(True) for t in (r, mm, ml))
    recr, _, nllr, spr = O.vessel_loss(rr, x.double(), None, m.double(), torch.zeros(1), torch.zeros(1), mmr, mlr)
    assert_close(rec, recr, FWD_TOL, "recon"); assert_close(sp, spr, FWD_TOL, "sparsity")
    assert_close(nll, nllr, FWD_TOL, "nll")
    (rec + 0.3 * sp + nll).backward(); (recr + 0.3 * spr + nllr).backward()
    assert_close(rg.grad, rr.grad, GRAD_TOL, "d recon"); assert_close(mmg.grad, mmr.grad, GRAD_TOL, "d m_mu")
    assert_close(mlg.grad, mlr.grad, GRAD_TOL, "d m_logvar")

    p = torch.sigmoid(gen(4, 784, seed=33)); y = torch.rand(4, 784, generator=torch.Generator().manual_seed(34))
    p[0, 0] = 0.0; p[0, 1] = 1.0
    pg2 = p.cuda().requires_grad_(True)
    b = F.bce_sum(pg2, y.cuda())
    pr2 = p.double().requires_grad_(True)
    br = O.bce_sum(pr2, y.double())
    assert_close(b, br, FWD_TOL, "bce")
    ms = F.mse_sum(pg2, y.cuda(), 2000.0)
    assert_close(ms, 2000.0 * ((p.double() - y.double()) ** 2).sum(), FWD_TOL, "mse")


def test_fused_clip_adam_matches_oracle():
    from causal_vae_b200.optim import FlatParams, FusedClipAdam
    mod = torch.nn.ModuleDict({"a": torch.nn.Linear(37, 19), "b": torch.nn.Linear(19, 5)}).cuda()
    flat = FlatParams(mod)
    opt = FusedClipAdam(flat, lr=1e-3, max_norm=5.0)
    P = {k: v.detach().cpu().clone() for k, v in mod.named_parameters()}
    state = {}
    for step in range(1, 4):
        grads = {k: gen(*v.shape, seed=40 + step, scale=3.0) for k, v in P.items()}
        opt.zero_grad()
        for k, p in mod.named_parameters():
            p.grad.copy_(grads[k].cuda())
        opt.step()
        clipped, total = O.clip_grad_norm({k: g.clone() for k, g in grads.items()}, 5.0)
        O.adam_step(P, clipped, state, step, 1e-3)
        assert abs(opt.grad_norm().item() - total.item()) <= 1e-5 * total.item()
        for k, p in mod.named_parameters():
            assert_close(p, P[k], 1e-5, f"step{step} {k}")


def test_dropout_statistics():
    from causal_vae_b200 import functional as F
    x = torch.ones(1 << 20, device="cuda", requires_grad=True)
    F.manual_seed(1234)
    y = F.dropout(x, 0.1, True)
    keep = (y != 0).float().mean().item()
    assert abs(keep - 0.9) < 3e-3
    assert abs(y.max().item() - 1.0 / 0.9) < 1e-6
    y.sum().backward()
    assert torch.equal((x.grad != 0), (y != 0)), "backward must regenerate the same mask"


def test_head_backward_fused_vs_separate_kernels_and_fp64():
    """cvae_head_bwd (weight + input gradient of the Conv 16->1 image head in one pass) against fp64 torch and
    against the separate weight-gradient / input-gradient kernels it replaces."""
    from causal_vae_b200 import _lib as L
    from causal_vae_b200 import ops
    N, H, W, C = 5, 120, 112, 16                       # ragged 8 x 32 patches
    assert ops.head_bwd_eligible(N, H, W, C)
    y = gen(N, H, W, C, seed=60)
    g = gen(N, H, W, 1, seed=61)
    w = gen(1, C, 3, 3, seed=62) * 0.2
    scale, shift, center = gen(C, seed=63).abs() + 0.5, gen(C, seed=64), gen(C, seed=65)
    xf = ops.XF(scale.cuda(), shift.cuda(), 0.01, center.cuda())
    stats = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    dw = torch.empty(1, C, 3, 3, device="cuda")
    dz = ops.head_bwd(g.cuda(), y.cuda(), xf, w.cuda(), dw, stats)
    # fp64 reference
    yr = y.double().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.double().requires_grad_(True)
    z = (yr - center.double().view(1, C, 1, 1)) * scale.double().view(1, C, 1, 1) + shift.double().view(1, C, 1, 1)
    a = torch.where(z > 0, z, z * 0.01)
    out = torch.nn.functional.conv2d(a, wr, None, 1, 1)
    out.backward(g.double().permute(0, 3, 1, 2))
    # dz is the gradient w.r.t. the pre-activation z (the BatchNorm-backward factor `scale` is applied later)
    want_dz = (yr.grad / scale.double().view(1, C, 1, 1)).permute(0, 2, 3, 1)
    assert_close(dz, want_dz, GRAD_TOL, "head dz")
    assert_close(dw, wr.grad, GRAD_TOL, "head dW")
    refc = (y.double() - center.double())
    want_st = torch.cat([want_dz.sum((0, 1, 2)), (want_dz * refc).sum((0, 1, 2))])
    assert rel(stats, want_st) <= 1e-5, rel(stats, want_st)


def test_batched_weight_packing_equals_per_use_packing():
    """ops.PackPlan: the one-launch re-pack reproduces every per-use layout bit for bit after the weights change."""
    from causal_vae_b200 import ops
    ws = [gen(64, 32, 9, seed=70).cuda(), gen(48, 16, 1, seed=71).cuda(), gen(16, 16, 9, seed=72).cuda(),
          gen(768, 256, 1, seed=73).cuda(), gen(128, 140, 1, seed=74).cuda(), gen(256, 64, 1, seed=75).cuda(),
          gen(64, 128, 9, seed=76).cuda(), gen(192, 96, 1, seed=77).cuda()]
    # (A, A_pad, B, taps, src_bat, src_ld, tc)
    args = [(32, 32, 64, 9, True, 32, True), (48, 48, 16, 1, False, 16, True), (16, 16, 16, 9, False, 16, False),
            (256, 256, 768, 1, True, 256, True),     # ViT Linear [out][in]: 128-bit source reads
            (140, 144, 128, 1, True, 140, True),     # ragged K (adapter): scalar tail, rows of 140 floats are not 16-byte aligned
            (64, 64, 256, 1, True, 64, False),       # fp32 transposition tiles (decoder_input layout)
            (64, 64, 128, 9, False, 128, True),      # transposed-conv weight [Cin][Cout][tap] on the tensor-core image
            (90, 96, 192, 1, True, 96, False)]       # fp32 transposition with zero-padded rows (A < A_pad)
    plan = ops.PackPlan()
    ops.set_pack_plan(plan)
    try:
        first = [ops.pack_weight(w, *a[:6], tc=a[6]) for w, a in zip(ws, args)]
        plan.finalize()
        for w in ws:
            w.mul_(1.7).add_(0.3)                       # "optimizer step"
        plan.run()
        cached = [ops.pack_weight(w, *a[:6], tc=a[6]) for w, a in zip(ws, args)]
    finally:
        ops.set_pack_plan(None)
    fresh = [ops.pack_weight(w, *a[:6], tc=a[6]) for w, a in zip(ws, args)]
    for c, f, o in zip(cached, fresh, first):
        assert c.data_ptr() == o.data_ptr()             # the plan's own buffer, re-packed in place
        assert torch.equal(c, f)


def test_fused_dropout_forms_draw_the_same_mask():
    """dropout(gelu(x)) and res + dropout(x) as single kernels: same values and gradients as the unfused
    sequence at the same point of the random stream (fusing must change launches, not masks)."""
    from causal_vae_b200 import _lib as L
    from causal_vae_b200 import functional as F
    x = gen(37, 65, 512, seed=50).cuda()
    res = gen(37, 65, 512, seed=51).cuda()
    gy = gen(37, 65, 512, seed=52).cuda()
    cases = [(lambda a: F.act_dropout(a, L.ACT_GELU, 0.1, True), lambda a: F.dropout(F.activation(a, L.ACT_GELU), 0.1, True)),
             (lambda a: F.dropout_add(a, res, 0.1, True), lambda a: F.add(res, F.dropout(a, 0.1, True)))]
    for fused, plain in cases:
        F.manual_seed(99)
        a1 = x.clone().requires_grad_(True)
        y1 = fused(a1)
        y1.backward(gy)
        F.manual_seed(99)
        a2 = x.clone().requires_grad_(True)
        y2 = plain(a2)
        y2.backward(gy)
        assert rel(y1, y2) <= 1e-6, rel(y1, y2)
        assert rel(a1.grad, a2.grad) <= 1e-6, rel(a1.grad, a2.grad)
    # odd length (vector tail) and p = 0 / eval fall back to the plain kernels
    v = gen(1027, seed=53).cuda()
    F.manual_seed(5)
    t1 = F.act_dropout(v, L.ACT_GELU, 0.25, True)
    F.manual_seed(5)
    t2 = F.dropout(F.activation(v, L.ACT_GELU), 0.25, True)
    assert rel(t1, t2) <= 1e-6
    assert rel(F.act_dropout(v, L.ACT_GELU, 0.25, False), F.activation(v, L.ACT_GELU)) == 0.0


@pytest.mark.parametrize("S,p", [(65, 0.0), (65, 0.3), (17, 0.1), (128, 0.2), (7, 0.0),
                                 (129, 0.0), (200, 0.25), (961, 0.0), (961, 0.1), (1025, 0.1)])
def test_attention_core_with_dropout(S, p):
    """softmax(QK^T/sqrt(d)) (dropout) V and its gradient against fp64 torch, using the very mask the
    kernel drew: forward saves it in the sign bits of the probabilities, backward reads it from there.
    S > 128 takes the strip kernels of csrc/attention_long.cu (961 = the reference's default 768x1280 image)."""
    from causal_vae_b200 import functional as F
    from causal_vae_b200 import ops
    B, H, d = (3, 8, 32) if S <= 200 else (2, 8, 32)
    D = H * d
    qkv = gen(B, S, 3 * D, seed=40, scale=0.7)
    g = gen(B, S, D, seed=41)
    F.manual_seed(77)
    seed, off, cnt = F.next_rng() if p > 0 else (0, 0, None)
    qg = qkv.cuda()
    out, probs = ops.attention_fwd(qg, B, S, H, d, p, seed, off, cnt)
    dq = ops.attention_bwd(qg, probs, g.cuda(), B, S, H, d, p, seed, off, cnt)
    mask = (~torch.signbit(probs)).double().cpu()
    if p > 0:
        assert abs(mask.mean().item() - (1 - p)) < 0.02
    else:
        assert mask.min().item() == 1.0
    qr = qkv.double().requires_grad_(True)
    q, k, v = (t.view(B, S, H, d).transpose(1, 2) for t in qr.chunk(3, dim=-1))
    P = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), dim=-1)
    assert_close(probs.abs(), P, FWD_TOL, "probabilities")
    o = ((P * mask / (1 - p)) @ v).transpose(1, 2).reshape(B, S, D)
    assert_close(out, o, FWD_TOL, "attention out")
    o.backward(g.double())
    assert_close(dq, qr.grad, GRAD_TOL, "dqkv")


def test_counterfactual_helpers():
    from causal_vae_b200.counterfactual import do_expand, rowdiff_l2
    S, K, Z = 5, 12, 128
    m, z = gen(S, K, seed=50), gen(S, Z, seed=51)
    out = do_expand(m.cuda(), z.cuda(), delta=5.0).cpu()
    for s in range(S):
        for k in range(K):
            ref = torch.cat([O.counterfactual_do(m[s:s + 1], k, delta=5.0), z[s:s + 1]], dim=1)[0]
            assert torch.equal(out[s * K + k], ref)          # index / scatter work is bit-exact
    a, b = gen(S * K, 1, 16, 16, seed=52), gen(S, 1, 16, 16, seed=53)
    d = rowdiff_l2(a.cuda(), b.cuda(), K).cpu()
    ref = (a.view(S, K, -1) - b.view(S, 1, -1)).double().norm(dim=2).view(-1)
    assert_close(d, ref, 1e-5, "rowdiff")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 32, 6, 10), (1, 3, 5, 7), (3, 512, 12, 20)])
def test_upsample_nearest2x_exact(shape):
    """nn.Upsample(scale_factor=2, mode='nearest') of the CNN vessel decoder: forward is a copy, backward a 2x2 sum —
    both bit-exact against torch on the same device-independent inputs."""
    from causal_vae_b200 import nn
    g = torch.Generator().manual_seed(1)
    x = torch.randn(*shape, generator=g)
    dy = torch.randn(shape[0], shape[1], 2 * shape[2], 2 * shape[3], generator=g)
    xr = x.clone().requires_grad_(True)
    yr = torch.nn.functional.interpolate(xr, scale_factor=2, mode="nearest")
    yr.backward(dy)
    xg = x.cuda().requires_grad_(True)
    y = nn.Upsample(scale_factor=2, mode="nearest")(xg)
    y.backward(dy.cuda())
    assert torch.equal(y.detach().cpu(), yr.detach())
    # 2x2 sums: (a + b) + (c + d) here, torch's order may differ by an ulp
    assert torch.allclose(xg.grad.cpu(), xr.grad, rtol=0, atol=4e-7 * float(dy.abs().max()))
