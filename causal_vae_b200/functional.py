"""autograd nodes over the native kernels (everything that is not a conv/linear chain)."""
import torch

from . import _lib as L
from . import ops

# ---- random-number state for dropout ------------------------------------------------------------
# Every dropout site of a forward pass takes a fresh `offset`; (seed, offset) are host integers baked
# into the launch, and `counter` is an optional device int64 (e.g. the optimizer step) mixed into the
# seed on the device so that a replayed CUDA graph still draws new masks every step.
#
# The state is an object, not a module global: every trainer owns one (seed derived from the user's
# torch.manual_seed / F.manual_seed and from the data-parallel RANK, so shards draw independent masks as the
# reference's single-device run over the concatenated batch would; counter = that trainer's own step counter)
# and installs it around its forward pass with `use_rng`.  Code outside a trainer uses the default state.
_MASK63 = 2 ** 63 - 1


class RngState:
    def __init__(self, seed=None, counter=None, rank=0):
        base = default_seed() if seed is None else int(seed)
        # splitmix-style mixing: distinct (seed, rank) pairs give unrelated Philox keys
        self.seed = ((base + 1) * 0x9E3779B97F4A7C15 + int(rank) * 0xD1B54A32D192ED03) & _MASK63
        self.offset = 0
        self.counter = counter


_explicit_seed = [None]
_default = [None]
_current = [None]


def default_seed():
    """F.manual_seed(s) if it was called, else the seed of torch's default generator (torch.manual_seed)."""
    return torch.initial_seed() & _MASK63 if _explicit_seed[0] is None else _explicit_seed[0]


def manual_seed(seed):
    _explicit_seed[0] = int(seed) & _MASK63
    _default[0] = None


def _state():
    if _current[0] is not None:
        return _current[0]
    if _default[0] is None:
        _default[0] = RngState()
    return _default[0]


class use_rng:
    """with use_rng(state): dropout sites inside draw from `state` (trainers wrap their forward pass)."""
    def __init__(self, state):
        self.state = state
    def __enter__(self):
        self.prev = _current[0]
        _current[0] = self.state
        return self.state
    def __exit__(self, *a):
        _current[0] = self.prev


def set_rng_counter(t):
    """device step counter of the DEFAULT state (trainers pass theirs through RngState(counter=...))."""
    _state().counter = t


def next_rng():
    st = _state()
    st.offset += 1
    return st.seed, st.offset, st.counter


# ---- layout -------------------------------------------------------------------------------------
class _TransposeBC(torch.autograd.Function):
    """[B, R, C] -> [B, C, R] contiguous."""

    @staticmethod
    def forward(ctx, x):
        B, R, Cc = x.shape
        return ops.transpose_bc(x.contiguous(), B, R, Cc)

    @staticmethod
    def backward(ctx, g):
        B, Cc, R = g.shape
        return ops.transpose_bc(g.contiguous(), B, Cc, R)


def to_nhwc(x):
    """logical NCHW (any strides) -> plain contiguous [N,H,W,C] tensor."""
    N, Cc, H, W = x.shape
    if Cc == 1:
        return x.contiguous().view(N, H, W, 1)
    if H * W == 1:
        return x.contiguous().view(N, 1, 1, Cc)
    v = x.permute(0, 2, 3, 1)
    if v.is_contiguous():
        return v
    return _TransposeBC.apply(x.contiguous().view(N, Cc, H * W)).view(N, H, W, Cc)


def from_nhwc(y):
    """[N,H,W,C] -> logical NCHW (a channels_last view; C == 1 gives a contiguous tensor)."""
    N, H, W, Cc = y.shape
    if Cc == 1 or H * W == 1:
        return y.reshape(N, Cc, H, W)
    return y.permute(0, 3, 1, 2)


def nchw_flatten(x):
    """x.flatten(1) in NCHW element order for a logical-NCHW tensor (nn.Flatten semantics)."""
    N, Cc, H, W = x.shape
    if x.is_contiguous():
        return x.view(N, -1)
    v = x.permute(0, 2, 3, 1)
    if v.is_contiguous():
        return _TransposeBC.apply(v.reshape(N, H * W, Cc)).view(N, Cc * H * W)
    return x.contiguous().view(N, -1)


class _CatPad(torch.autograd.Function):
    """torch.cat(parts, dim=1) into a zero-padded [B, pad4(sum)] buffer (row width a multiple of 4
    so the GEMM kernels can use 128-bit loads; e.g. 256+12+19 = 287 -> 288)."""

    @staticmethod
    def forward(ctx, *parts):
        B = parts[0].shape[0]
        widths = [p.shape[1] for p in parts]
        total = sum(widths)
        ld = (total + 3) // 4 * 4
        out = ops.zeros(B, ld, like=parts[0]) if ld != total else ops.empty(B, ld, like=parts[0])
        c0 = 0
        for p, w in zip(parts, widths):
            pc = p.contiguous()
            ops.copy_cols(pc, w, 0, out, ld, c0, B, w)
            c0 += w
        ctx.widths, ctx.ld = widths, ld
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        B = g.shape[0]
        outs, c0 = [], 0
        for i, w in enumerate(ctx.widths):
            if ctx.needs_input_grad[i]:
                d = ops.empty(B, w, like=g)
                ops.copy_cols(g, ctx.ld, c0, d, w, 0, B, w)
                outs.append(d)
            else:
                outs.append(None)
            c0 += w
        return tuple(outs)


def cat_pad(parts):
    return _CatPad.apply(*parts)


# ---- elementwise ---------------------------------------------------------------------------------
class _Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act, slope):
        xc = x.contiguous()
        y = ops.act_fwd(xc, act, slope)
        ctx.act, ctx.slope = act, slope
        ctx.save_for_backward(y if act == L.ACT_SIGMOID else xc)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, g):
        (ref,) = ctx.saved_tensors
        return ops.act_bwd(g.contiguous(), ref, ctx.act, ctx.slope).view(g.shape), None, None


def _elementwise_view(x):
    """A dense tensor in its own memory order (channels_last views stay as they are)."""
    if x.dim() == 4 and not x.is_contiguous() and x.permute(0, 2, 3, 1).is_contiguous():
        return x.permute(0, 2, 3, 1), True
    return x.contiguous(), False


def activation(x, act, slope=0.0):
    v, perm = _elementwise_view(x)
    y = _Act.apply(v, act, float(slope))
    return y.permute(0, 3, 1, 2) if perm else y


class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p):
        ctx.p = p
        ctx.rng = next_rng()
        seed, off, cnt = ctx.rng
        return ops.dropout(x.contiguous(), p, seed, off, cnt).view(x.shape)

    @staticmethod
    def backward(ctx, g):
        seed, off, cnt = ctx.rng
        return ops.dropout(g.contiguous(), ctx.p, seed, off, cnt).view(g.shape), None


def dropout(x, p, training):
    if not training or p == 0.0:
        return x
    v, perm = _elementwise_view(x)
    y = _Dropout.apply(v, float(p))
    return y.permute(0, 3, 1, 2) if perm else y


class _ActDropout(torch.autograd.Function):
    """dropout(act(x)) as one kernel each way (same mask as _Dropout would draw at this point of the stream)."""

    @staticmethod
    def forward(ctx, x, act, slope, p):
        xc = x.contiguous()
        ctx.cfg = (act, slope, p, next_rng())
        seed, off, cnt = ctx.cfg[3]
        y = torch.empty_like(xc)
        L.check(L.lib.cvae_dropout_fused(L.ptr(xc), None, L.ptr(y), xc.numel(), 0, act, slope, p, seed, off, L.ptr(cnt),
                                         L.stream()), "act_dropout_fwd")
        ctx.save_for_backward(xc)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, g):
        (xc,) = ctx.saved_tensors
        act, slope, p, (seed, off, cnt) = ctx.cfg
        gc = g.contiguous()
        dx = torch.empty_like(xc)
        L.check(L.lib.cvae_dropout_fused(L.ptr(xc), L.ptr(gc), L.ptr(dx), xc.numel(), 1, act, slope, p, seed, off,
                                         L.ptr(cnt), L.stream()), "act_dropout_bwd")
        return dx.view(g.shape), None, None, None


def act_dropout(x, act, p, training, slope=0.0):
    """nn.Dropout(p)(activation(x))"""
    if not training or p == 0.0:
        return activation(x, act, slope)
    v, perm = _elementwise_view(x)
    y = _ActDropout.apply(v, act, float(slope), float(p))
    return y.permute(0, 3, 1, 2) if perm else y


class _DropoutAdd(torch.autograd.Function):
    """res + dropout(x)"""

    @staticmethod
    def forward(ctx, x, res, p):
        xc, rc = x.contiguous(), res.contiguous()
        ctx.p, ctx.rng = p, next_rng()
        seed, off, cnt = ctx.rng
        y = torch.empty_like(xc)
        L.check(L.lib.cvae_dropout_fused(L.ptr(xc), L.ptr(rc), L.ptr(y), xc.numel(), 2, 0, 0.0, p, seed, off, L.ptr(cnt),
                                         L.stream()), "dropout_add")
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, g):
        seed, off, cnt = ctx.rng
        gc = g.contiguous()
        return ops.dropout(gc, ctx.p, seed, off, cnt).view(g.shape), g, None


def dropout_add(x, res, p, training):
    """res + nn.Dropout(p)(x)"""
    if not training or p == 0.0:
        return add(res, x)
    if x.shape != res.shape or x.dim() == 4:
        return add(res, dropout(x, p, training))
    return _DropoutAdd.apply(x, res, float(p))


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return ops.add(a.contiguous(), b.contiguous()).view(a.shape)

    @staticmethod
    def backward(ctx, g):
        return g, g


def add(a, b):
    return _Add.apply(a, b)


class _Clamp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lo, hi):
        xc = x.contiguous()
        y = torch.empty_like(xc)
        L.check(L.lib.cvae_clamp_fwd(L.ptr(xc), L.ptr(y), xc.numel(), lo, hi, L.stream()), "clamp_fwd")
        ctx.save_for_backward(xc)
        ctx.lo, ctx.hi = lo, hi
        return y

    @staticmethod
    def backward(ctx, g):
        (xc,) = ctx.saved_tensors
        dx = torch.empty_like(xc)
        L.check(L.lib.cvae_clamp_bwd(L.ptr(g.contiguous()), L.ptr(xc), L.ptr(dx), xc.numel(), ctx.lo, ctx.hi,
                                     L.stream()), "clamp_bwd")
        return dx, None, None


def clamp(x, lo, hi):
    return _Clamp.apply(x, float(lo), float(hi))


class _Upsample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):                                   # x: contiguous [N, H, W, C]
        N, H, W, Cc = x.shape
        y = torch.empty(N, 2 * H, 2 * W, Cc, dtype=x.dtype, device=x.device)
        L.check(L.lib.cvae_upsample2x_fwd(L.ptr(x), L.ptr(y), N, H, W, Cc, L.stream()), "upsample2x_fwd")
        ctx.shape = (N, H, W, Cc)
        return y

    @staticmethod
    def backward(ctx, g):
        N, H, W, Cc = ctx.shape
        dx = torch.empty(N, H, W, Cc, dtype=g.dtype, device=g.device)
        L.check(L.lib.cvae_upsample2x_bwd(L.ptr(g.contiguous()), L.ptr(dx), N, H, W, Cc, L.stream()), "upsample2x_bwd")
        return dx


def upsample_nearest2x(x):
    """nn.Upsample(scale_factor=2, mode='nearest') on a logical NCHW tensor (vessel_analysis/00_core/models.py:123)."""
    return from_nhwc(_Upsample2x.apply(to_nhwc(x).contiguous()))


# ---- LayerNorm ------------------------------------------------------------------------------------
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        D = x.shape[-1]
        if x.dim() == 2 and x.stride(1) == 1 and not x.is_contiguous():
            xs, rows, xr = x.stride(0), x.shape[0], x            # strided rows (e.g. the CLS slice tok[:, 0])
        else:
            xr = x.contiguous()
            xs, rows = D, xr.numel() // D
        y, mean, rstd = ops.layernorm_fwd(xr, gamma, beta, rows, D, xs, eps)
        ctx.save_for_backward(xr, gamma, mean, rstd)
        ctx.geom = (rows, D, xs)
        ctx.params = (gamma, beta)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, g):
        xr, gamma, mean, rstd = ctx.saved_tensors
        rows, D, xs = ctx.geom
        g = g.contiguous()
        dx = ops.empty(rows, D, like=g)
        from .chain import direct_ok
        pg, pb = ctx.params
        if direct_ok(pg) and direct_ok(pb):          # the kernel accumulates into the (zeroed) .grad views
            ops.layernorm_bwd(g, xr, gamma, mean, rstd, rows, D, xs, dx, D, False, pg.grad, pb.grad)
            return dx.view(g.shape), None, None, None
        dgamma, dbeta = ops.zeros(D, like=g), ops.zeros(D, like=g)
        ops.layernorm_bwd(g, xr, gamma, mean, rstd, rows, D, xs, dx, D, False, dgamma, dbeta)
        return dx.view(g.shape), dgamma, dbeta, None


def layer_norm(x, gamma, beta, eps):
    return _LayerNorm.apply(x, gamma, beta, float(eps))


class _AddLayerNorm(torch.autograd.Function):
    """(s, y) = (a + b, LayerNorm(a + b)): the residual add in front of a LayerNorm as one kernel; backward adds the
    gradient reaching `s` through its other consumer inside the LayerNorm-backward kernel (no accumulate launch)."""

    @staticmethod
    def forward(ctx, a, b, gamma, beta, eps):
        D = a.shape[-1]
        ac, bc = a.contiguous(), b.contiguous()
        rows = ac.numel() // D
        s, y = torch.empty_like(ac), torch.empty_like(ac)
        mean, rstd = ops.empty(rows, like=ac), ops.empty(rows, like=ac)
        L.check(L.lib.cvae_add_layernorm_fwd(L.ptr(ac), L.ptr(bc), L.ptr(gamma), L.ptr(beta), L.ptr(s), L.ptr(y),
                                             L.ptr(mean), L.ptr(rstd), rows, D, float(eps), L.stream()), "add_layernorm_fwd")
        ctx.save_for_backward(s, gamma, mean, rstd)
        ctx.params = (gamma, beta)
        ctx.geom = (rows, D)
        return s, y

    @staticmethod
    def backward(ctx, gs, gy):
        from .chain import direct_ok
        s, gamma, mean, rstd = ctx.saved_tensors
        rows, D = ctx.geom
        pg, pb = ctx.params
        direct = direct_ok(pg) and direct_ok(pb)
        dgamma, dbeta = (pg.grad, pb.grad) if direct else (ops.zeros(D, like=s), ops.zeros(D, like=s))
        if gy is None:
            return gs, gs, None, None, None
        gy = gy.contiguous()
        dx = torch.empty_like(s)
        if gs is not None and D % 32 == 0 and D <= 256:
            L.check(L.lib.cvae_layernorm_bwd_add(L.ptr(gy), L.ptr(s), L.ptr(gamma), L.ptr(mean), L.ptr(rstd),
                                                 L.ptr(gs.contiguous()), L.ptr(dx), L.ptr(dgamma), L.ptr(dbeta), rows, D,
                                                 L.stream()), "layernorm_bwd_add")
        else:
            ops.layernorm_bwd(gy, s, gamma, mean, rstd, rows, D, D, dx, D, False, dgamma, dbeta)
            if gs is not None:
                dx = ops.add(dx, gs.contiguous())
        dx = dx.view(s.shape)
        return dx, dx, (None if direct else dgamma), (None if direct else dbeta), None


def add_layer_norm(a, b, gamma, beta, eps):
    """(a + b, layer_norm(a + b))"""
    return _AddLayerNorm.apply(a, b, gamma, beta, float(eps))


# ---- attention core --------------------------------------------------------------------------------
class _AttnCore(torch.autograd.Function):
    """qkv [B,S,3D] -> softmax(QK^T/sqrt(d)) V merged over heads [B,S,D] (vit_backbone.py:28-30,43)."""

    @staticmethod
    def forward(ctx, qkv, H, p):
        B, S, D3 = qkv.shape
        d = D3 // 3 // H
        qkv = qkv.contiguous()
        seed, off, cnt = next_rng() if p > 0 else (0, 0, None)
        out, probs = ops.attention_fwd(qkv, B, S, H, d, p, seed, off, cnt)
        ctx.save_for_backward(qkv, probs)
        ctx.cfg = (B, S, H, d, p, seed, off, cnt)
        return out

    @staticmethod
    def backward(ctx, g):
        qkv, probs = ctx.saved_tensors
        B, S, H, d, p, seed, off, cnt = ctx.cfg
        return ops.attention_bwd(qkv, probs, g.contiguous(), B, S, H, d, p, seed, off, cnt), None, None


def attention_core(qkv, heads, p):
    return _AttnCore.apply(qkv, heads, float(p))


# ---- tokens ---------------------------------------------------------------------------------------
class _Tokens(torch.autograd.Function):
    """feat NHWC [B,h,w,D] -> tokens [B, h*w+1, D] = cat(cls, feat) + pos[:, :n+1]
    (vessel_analysis/00_core/models.py:267-271).  NHWC makes `b c h w -> b (h w) c` a no-op."""

    @staticmethod
    def forward(ctx, feat, cls, pos):
        B, h, w, D = feat.shape
        n = h * w
        tok = ops.empty(B, n + 1, D, like=feat)
        posc = pos.contiguous()
        L.check(L.lib.cvae_tokens_fwd(L.ptr(feat.contiguous()), L.ptr(cls), L.ptr(posc), L.ptr(tok), B, n, D,
                                      L.stream()), "tokens_fwd")
        ctx.geom = (B, h, w, D, pos.shape[1])
        ctx.params = (cls, pos)
        return tok

    @staticmethod
    def backward(ctx, g):
        B, h, w, D, npos = ctx.geom
        n = h * w
        g = g.contiguous()
        dfeat = ops.empty(B, h, w, D, like=g) if ctx.needs_input_grad[0] else None
        from .chain import direct_ok
        pc, pp = ctx.params
        direct = direct_ok(pc) and direct_ok(pp) and pc.grad.numel() == D and pp.grad.numel() == npos * D
        if direct:                                   # rows of pos beyond n+1 keep their zero_grad() zeros
            dcls, dpos = pc.grad, pp.grad
        else:
            dcls = ops.empty(1, 1, D, like=g)
            dpos = ops.zeros(1, npos, D, like=g) if npos != n + 1 else ops.empty(1, npos, D, like=g)
        L.check(L.lib.cvae_tokens_bwd(L.ptr(g), L.ptr(dfeat), L.ptr(dcls), L.ptr(dpos), B, n, D, L.stream()),
                "tokens_bwd")
        return (dfeat, None, None) if direct else (dfeat, dcls, dpos)


def tokens(feat_nhwc, cls, pos):
    return _Tokens.apply(feat_nhwc, cls, pos)


# ---- latent ---------------------------------------------------------------------------------------
class _Latent(torch.autograd.Function):
    """h [B,2Z] -> (mu, logvar, z): chunk, clamp, reparameterise with the given eps
    (vessel_analysis/00_core/models.py:282-288)."""

    @staticmethod
    def forward(ctx, h, eps, mu_clamp, lv_clamp):
        B, Z2 = h.shape
        Z = Z2 // 2
        hc, ec = h.contiguous(), eps.contiguous()
        mu, lv, z = ops.empty(B, Z, like=h), ops.empty(B, Z, like=h), ops.empty(B, Z, like=h)
        L.check(L.lib.cvae_latent_fwd(L.ptr(hc), L.ptr(ec), L.ptr(mu), L.ptr(lv), L.ptr(z), None, B, Z, mu_clamp,
                                      lv_clamp, L.stream()), "latent_fwd")
        ctx.save_for_backward(hc, ec)
        ctx.cfg = (B, Z, mu_clamp, lv_clamp)
        return mu, lv, z

    @staticmethod
    def backward(ctx, dmu, dlv, dz):
        hc, ec = ctx.saved_tensors
        B, Z, mc, lc = ctx.cfg
        dh = torch.empty_like(hc)
        c = lambda t: None if t is None else t.contiguous()
        L.check(L.lib.cvae_latent_bwd(L.ptr(hc), L.ptr(ec), L.ptr(c(dz)), L.ptr(c(dmu)), L.ptr(c(dlv)), L.ptr(dh), B,
                                      Z, mc, lc, L.stream()), "latent_bwd")
        return dh, None, None, None


def latent(h, eps, mu_clamp=0.0, lv_clamp=0.0):
    return _Latent.apply(h, eps, float(mu_clamp), float(lv_clamp))


class _Reparam(torch.autograd.Function):
    """z = mu + eps * exp(0.5*logvar) for separate mu / logvar tensors (reparameterize())."""

    @staticmethod
    def forward(ctx, mu, logvar, eps):
        B, Z = mu.shape
        h = ops.empty(B, 2 * Z, like=mu)
        ops.copy_cols(mu.contiguous(), Z, 0, h, 2 * Z, 0, B, Z)
        ops.copy_cols(logvar.contiguous(), Z, 0, h, 2 * Z, Z, B, Z)
        m2, l2, z = ops.empty(B, Z, like=mu), ops.empty(B, Z, like=mu), ops.empty(B, Z, like=mu)
        ec = eps.contiguous()
        L.check(L.lib.cvae_latent_fwd(L.ptr(h), L.ptr(ec), L.ptr(m2), L.ptr(l2), L.ptr(z), None, B, Z, 0.0, 0.0,
                                      L.stream()), "latent_fwd")
        ctx.save_for_backward(h, ec)
        return z

    @staticmethod
    def backward(ctx, dz):
        h, ec = ctx.saved_tensors
        B, Z2 = h.shape
        Z = Z2 // 2
        dh = torch.empty_like(h)
        L.check(L.lib.cvae_latent_bwd(L.ptr(h), L.ptr(ec), L.ptr(dz.contiguous()), None, None, L.ptr(dh), B, Z, 0.0,
                                      0.0, L.stream()), "latent_bwd")
        dmu, dlv = ops.empty(B, Z, like=h), ops.empty(B, Z, like=h)
        ops.copy_cols(dh, Z2, 0, dmu, Z, 0, B, Z)
        ops.copy_cols(dh, Z2, Z, dlv, Z, 0, B, Z)
        return dmu, dlv, None


def reparameterize(mu, logvar, eps=None):
    if eps is None:
        eps = torch.randn_like(mu)
    return _Reparam.apply(mu, logvar, eps)


# ---- losses ---------------------------------------------------------------------------------------
class _VesselRecon(torch.autograd.Function):
    """(recon_loss, sparsity_loss) of vessel_analysis/01_train/train.py:27-46."""

    @staticmethod
    def forward(ctx, recon, x):
        rc, xc = recon.contiguous(), x.contiguous()
        n = rc.numel()
        sums = ops.zeros(3, dtype=torch.float64, like=rc)
        L.check(L.lib.cvae_vessel_xsum(L.ptr(xc), n, L.ptr(sums), L.stream()), "vessel_xsum")
        L.check(L.lib.cvae_vessel_recon_fwd(L.ptr(rc), L.ptr(xc), n, L.ptr(sums), L.stream()), "vessel_recon_fwd")
        ctx.save_for_backward(rc, xc, sums)
        return ops.finish_scalar(sums[1:2]), ops.finish_scalar(sums[2:3])

    @staticmethod
    def backward(ctx, g_rec, g_sp):
        rc, xc, sums = ctx.saved_tensors
        d = torch.empty_like(rc)
        L.check(L.lib.cvae_vessel_recon_bwd(L.ptr(rc), L.ptr(xc), rc.numel(), L.ptr(sums), L.ptr(g_rec.contiguous()),
                                            L.ptr(g_sp.contiguous()), L.ptr(d), L.stream()), "vessel_recon_bwd")
        return d, None


def vessel_recon_loss(recon, x):
    return _VesselRecon.apply(recon, x)


class _Kld(torch.autograd.Function):
    """-0.5 * sum(1 + logvar - mu^2 - exp(logvar))  (train.py:49), scaled by `mul`."""

    @staticmethod
    def forward(ctx, mu, logvar, mul):
        mc, lc = mu.contiguous(), logvar.contiguous()
        acc = ops.zeros(1, dtype=torch.float64, like=mc)
        L.check(L.lib.cvae_kld_fwd(L.ptr(mc), L.ptr(lc), mc.numel(), L.ptr(acc), L.stream()), "kld_fwd")
        ctx.save_for_backward(mc, lc)
        ctx.mul = mul
        return ops.finish_scalar(acc, mul)

    @staticmethod
    def backward(ctx, g):
        mc, lc = ctx.saved_tensors
        dmu, dlv = torch.empty_like(mc), torch.empty_like(lc)
        L.check(L.lib.cvae_kld_bwd(L.ptr(mc), L.ptr(lc), L.ptr(g.contiguous()), ctx.mul, L.ptr(dmu), L.ptr(dlv),
                                   mc.numel(), 0, L.stream()), "kld_bwd")
        return dmu, dlv, None


def kld_loss(mu, logvar, mul=1.0):
    return _Kld.apply(mu, logvar, float(mul))


class _GaussNLL(torch.autograd.Function):
    """0.5 * sum(logvar + (m - mu)^2 / exp(logvar))  (train.py:55-58)."""

    @staticmethod
    def forward(ctx, m, m_mu, m_logvar):
        mc, uc, lc = m.contiguous(), m_mu.contiguous(), m_logvar.contiguous()
        acc = ops.zeros(1, dtype=torch.float64, like=mc)
        L.check(L.lib.cvae_gauss_nll_fwd(L.ptr(mc), L.ptr(uc), L.ptr(lc), None, L.ptr(acc), mc.numel(), 0.0,
                                         L.stream()), "gauss_nll_fwd")
        ctx.save_for_backward(mc, uc, lc)
        return ops.finish_scalar(acc)

    @staticmethod
    def backward(ctx, g):
        mc, uc, lc = ctx.saved_tensors
        dmu, dlv = torch.empty_like(uc), torch.empty_like(lc)
        L.check(L.lib.cvae_gauss_nll_bwd(L.ptr(mc), L.ptr(uc), L.ptr(lc), L.ptr(g.contiguous()), 1.0, L.ptr(dmu),
                                         L.ptr(dlv), mc.numel(), 0.0, L.stream()), "gauss_nll_bwd")
        return None, dmu, dlv


def gauss_nll_loss(m, m_mu, m_logvar):
    return _GaussNLL.apply(m, m_mu, m_logvar)


class _PairLoss(torch.autograd.Function):
    """sum-reduced squared error (kind 0) or BCE with log clamp -100 (kind 1), times `mul`;
    gradient w.r.t. the first argument only."""

    @staticmethod
    def forward(ctx, a, b, kind, mul):
        ac, bc = a.contiguous(), b.contiguous()
        acc = ops.zeros(1, dtype=torch.float64, like=ac)
        fn = L.lib.cvae_mse_fwd if kind == 0 else L.lib.cvae_bce_fwd
        L.check(fn(L.ptr(ac), L.ptr(bc), ac.numel(), L.ptr(acc), L.stream()), "pair_loss_fwd")
        ctx.save_for_backward(ac, bc)
        ctx.kind, ctx.mul = kind, mul
        return ops.finish_scalar(acc, mul)

    @staticmethod
    def backward(ctx, g):
        ac, bc = ctx.saved_tensors
        da = torch.empty_like(ac)
        fn = L.lib.cvae_mse_bwd if ctx.kind == 0 else L.lib.cvae_bce_bwd
        L.check(fn(L.ptr(ac), L.ptr(bc), ac.numel(), L.ptr(g.contiguous()), ctx.mul, L.ptr(da), L.stream()),
                "pair_loss_bwd")
        return da, None, None, None


def mse_sum(a, b, mul=1.0):
    return _PairLoss.apply(a, b, 0, float(mul))


def bce_sum(p, y, mul=1.0):
    return _PairLoss.apply(p, y, 1, float(mul))


def mse_mean(a, b):
    """F.mse_loss(a, b, reduction='mean') (latent_translator/engine.py:25)."""
    return _PairLoss.apply(a, b, 0, 1.0 / a.numel())


def kld_mean(mu, logvar):
    """-0.5 * mean(1 + logvar - mu^2 - exp(logvar)) (latent_translator/engine.py:26)."""
    return _Kld.apply(mu, logvar, 1.0 / mu.numel())


class _WeightedSum(torch.autograd.Function):
    """sum_i w_i * x_i over up to four 0-dim loss tensors, one kernel each way."""

    @staticmethod
    def forward(ctx, w, *xs):
        ctx.w = tuple(float(v) for v in w) + (0.0,) * (4 - len(w))
        ctx.n = len(xs)
        ps = [L.ptr(x.contiguous()) for x in xs] + [None] * (4 - len(xs))
        out = torch.empty((), dtype=torch.float32, device=xs[0].device)
        L.check(L.lib.cvae_scalar_combine(ps[0], ps[1], ps[2], ps[3], *ctx.w, L.ptr(out), L.stream()), "scalar_combine")
        return out

    @staticmethod
    def backward(ctx, g):
        out4 = torch.empty(4, dtype=torch.float32, device=g.device)
        L.check(L.lib.cvae_scalar_scale4(L.ptr(g.contiguous()), *ctx.w, L.ptr(out4), L.stream()), "scalar_scale4")
        return (None, *[out4[i] for i in range(ctx.n)])


def weighted_sum(xs, ws):
    """sum_i ws[i] * xs[i] for 0-dim CUDA tensors (at most four)."""
    assert 1 <= len(xs) <= 4 and len(xs) == len(ws)
    return _WeightedSum.apply(tuple(ws), *xs)


# ---- treatment labels -------------------------------------------------------------------------------
def argmax_rows(t):
    """torch.argmax(t, dim=1) -> int64 [rows] (mnist_test/01_baseline_causal_vae/train.py:38)."""
    tc = t.contiguous()
    rows, T = tc.shape
    out = torch.empty(rows, dtype=torch.int64, device=tc.device)
    L.check(L.lib.cvae_argmax_rows(L.ptr(tc), rows, T, L.ptr(out), L.stream()), "argmax_rows")
    return out


def one_hot(idx, T):
    """F.one_hot(idx, T).float() (causal_cascade/models.py:71)."""
    if idx.dtype != torch.int64:
        raise RuntimeError(f"one_hot expects int64 class indices, got {idx.dtype}")
    ic = idx.contiguous()
    out = ops.empty(ic.numel(), T, like=ic)
    L.check(L.lib.cvae_one_hot(L.ptr(ic), ic.numel(), T, L.ptr(out), L.stream()), "one_hot")
    return out


class _SoftmaxLoss(torch.autograd.Function):
    """kind 0: F.cross_entropy(logits, target) (mean over rows); kind 1:
    F.kl_div(log_softmax(logits), U(1/T), reduction='batchmean'); both times `mul`."""

    @staticmethod
    def forward(ctx, logits, target, kind, mul):
        lc = logits.contiguous()
        rows, T = lc.shape
        acc = ops.zeros(1, dtype=torch.float64, like=lc)
        if kind == 0:
            L.check(L.lib.cvae_softmax_ce_fwd(L.ptr(lc), L.ptr(target), rows, T, L.ptr(acc), L.stream()), "ce_fwd")
        else:
            L.check(L.lib.cvae_uniform_kl_fwd(L.ptr(lc), rows, T, L.ptr(acc), L.stream()), "ukl_fwd")
        ctx.save_for_backward(lc)
        ctx.target, ctx.kind, ctx.scale = target, kind, mul / rows
        return ops.finish_scalar(acc, mul / rows)

    @staticmethod
    def backward(ctx, g):
        (lc,) = ctx.saved_tensors
        rows, T = lc.shape
        d = torch.empty_like(lc)
        if ctx.kind == 0:
            L.check(L.lib.cvae_softmax_ce_bwd(L.ptr(lc), L.ptr(ctx.target), rows, T, L.ptr(g.contiguous()), ctx.scale,
                                              L.ptr(d), L.stream()), "ce_bwd")
        else:
            L.check(L.lib.cvae_uniform_kl_bwd(L.ptr(lc), rows, T, L.ptr(g.contiguous()), ctx.scale, L.ptr(d),
                                              L.stream()), "ukl_bwd")
        return d, None, None, None


def cross_entropy(logits, target):
    return _SoftmaxLoss.apply(logits, target.contiguous(), 0, 1.0)


def uniform_kl_batchmean(logits, mul=1.0):
    return _SoftmaxLoss.apply(logits, None, 1, float(mul))
