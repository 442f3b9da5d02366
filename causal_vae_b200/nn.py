"""Drop-in layer classes: same constructors, parameter names and `state_dict` layout as `torch.nn`,
forward/backward on the native sm_100a kernels.

They subclass the `torch.nn` classes purely as *parameter containers* (so `state_dict()`,
`load_state_dict()`, `.to()`, `.eval()`, hooks, `Sequential` indexing and `torch.optim.Adam` behave
exactly as the reference's callers expect — SURVEY §8(b)); every `forward` is replaced.  There is no
ATen compute fallback: a CPU tensor raises.

`Sequential` fuses runs of [Conv2d | ConvTranspose2d | Linear] (+ BatchNorm) (+ LeakyReLU/ReLU/
Sigmoid) and ResBlocks into one chain (see chain.py); each layer is still callable on its own.
"""
import torch
import torch.nn as tnn

from . import _lib as L
from . import functional as F
from .chain import ResUnit, Unit, run_chain


def _check_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("causal_vae_b200 modules run on CUDA tensors only (no CPU fallback); "
                           "move the model and inputs to a B200 device")
    if x.dtype != torch.float32:
        raise RuntimeError(f"causal_vae_b200 computes in fp32; got {x.dtype}")


def _run_units(units, x):
    """x: logical NCHW 4-D tensor or [B, K] matrix -> same kind of tensor."""
    _check_cuda(x)
    if x.dim() == 2:
        B, K = x.shape
        if K % 4 != 0:
            x = F.cat_pad([x])
        y = run_chain(units, x.contiguous().view(B, 1, 1, x.shape[1]), True)
        return y.view(B, y.shape[-1])
    if x.dim() == 4:
        first = units[0] if isinstance(units[0], Unit) else units[0].u1
        if first.kind == "linear":
            raise RuntimeError("Linear expects a 2-D input")
        return F.from_nhwc(run_chain(units, F.to_nhwc(x), True))
    if x.dim() == 3:           # [B, S, K] token matrices through Linear layers
        B, S, K = x.shape
        y = _run_units(units, x.reshape(B * S, K))
        return y.view(B, S, y.shape[-1])
    raise RuntimeError(f"unsupported input rank {x.dim()}")


class Conv2d(tnn.Conv2d):
    def forward(self, x):
        return _run_units([Unit("conv", self)], x)


class ConvTranspose2d(tnn.ConvTranspose2d):
    def forward(self, x, output_size=None):
        return _run_units([Unit("convT", self)], x)


class Linear(tnn.Linear):
    def forward(self, x):
        return _run_units([Unit("linear", self)], x)


class _ActBase:
    pass


class LeakyReLU(tnn.LeakyReLU, _ActBase):
    def forward(self, x):
        _check_cuda(x)
        return F.activation(x, L.ACT_LRELU, self.negative_slope)


class ReLU(tnn.ReLU, _ActBase):
    def forward(self, x):
        _check_cuda(x)
        return F.activation(x, L.ACT_LRELU, 0.0)


class Sigmoid(tnn.Sigmoid):
    def forward(self, x):
        _check_cuda(x)
        return F.activation(x, L.ACT_SIGMOID)


class GELU(tnn.GELU):
    def forward(self, x):
        _check_cuda(x)
        return F.activation(x, L.ACT_GELU)


class Dropout(tnn.Dropout):
    def forward(self, x):
        return F.dropout(x, self.p, self.training)


class LayerNorm(tnn.LayerNorm):
    def forward(self, x):
        _check_cuda(x)
        return F.layer_norm(x, self.weight, self.bias, self.eps)


class Flatten(tnn.Flatten):
    def forward(self, x):
        if x.dim() == 4 and self.start_dim == 1:
            return F.nchw_flatten(x)
        return super().forward(x)


class Upsample(tnn.Upsample):
    """nearest x2 only (the CNN vessel decoder, vessel_analysis/00_core/models.py:123-145)."""

    def forward(self, x):
        _check_cuda(x)
        sf = self.scale_factor
        sf = sf if isinstance(sf, (int, float)) else (sf[0] if sf and sf[0] == sf[-1] else None)
        if self.mode != "nearest" or sf is None or float(sf) != 2.0 or x.dim() != 4:
            raise RuntimeError("Upsample: only scale_factor=2, mode='nearest' on 4-D input is implemented")
        return F.upsample_nearest2x(x)


class AdaptiveAvgPool2d(tnn.AdaptiveAvgPool2d):
    """Identity when the input already has the target size (64x64 cascade config: 4x4 -> 4x4,
    causal_cascade/models.py:18); other sizes are outside the B200 hot path."""

    def forward(self, x):
        tgt = self.output_size if isinstance(self.output_size, tuple) else (self.output_size,) * 2
        if tuple(x.shape[2:]) == tuple(tgt):
            return x
        raise RuntimeError(f"AdaptiveAvgPool2d {tuple(x.shape[2:])}->{tgt}: only the identity case is implemented")


class _BNStandalone(torch.autograd.Function):
    """BatchNorm applied on its own (not fused behind a conv): column statistics + finalize + affine."""

    @staticmethod
    def forward(ctx, x2d, bn, gamma, beta):
        from . import ops
        rows, C = x2d.shape
        if bn.training or not bn.track_running_stats:
            stats = ops.zeros(2 * C, dtype=torch.float64, like=x2d)
            ops.col_stats(rows, C, x2d, stats)
            scale, shift, mean, rstd = ops.bn_finalize(stats, C, rows, bn, train_buffers=bn.training)
            ctx.train = True
            ctx.save_for_backward(x2d, gamma, mean, rstd)
        else:
            scale, shift = ops.bn_eval_coeffs(bn)
            mean = bn.running_mean
            ctx.train = False
            ctx.save_for_backward(x2d, scale)
        return ops.affine_act(x2d, ops.XF(scale, shift, 1.0, mean))

    @staticmethod
    def backward(ctx, g):
        from . import ops
        g = g.contiguous()
        if not ctx.train:
            x2d, scale = ctx.saved_tensors
            zero = torch.zeros_like(scale)
            return ops.bn_bwd_apply(g, x2d, scale, zero, zero, zero), None, None, None
        x2d, gamma, mean, rstd = ctx.saved_tensors
        rows, C = x2d.shape
        stats = ops.zeros(2 * C, dtype=torch.float64, like=g)
        one = torch.ones_like(mean)
        dz = ops.dact_stats(g, x2d, ops.XF(one, torch.zeros_like(mean), 1.0, mean), stats)
        ca, cb, cc, dg, db, _ = ops.bn_bwd_finalize(stats, C, rows, gamma, mean, rstd, False)
        return ops.bn_bwd_apply(dz, x2d, ca, cb, cc, mean), None, dg, db


class _BNMixin:
    def forward(self, x):
        _check_cuda(x)
        if x.dim() == 4:
            v = F.to_nhwc(x)
            y = _BNStandalone.apply(v.reshape(-1, v.shape[-1]), self, self.weight, self.bias)
            return F.from_nhwc(y.view(v.shape))
        return _BNStandalone.apply(x.contiguous(), self, self.weight, self.bias)


class BatchNorm2d(_BNMixin, tnn.BatchNorm2d):
    pass


class BatchNorm1d(_BNMixin, tnn.BatchNorm1d):
    pass


class ResBlock(tnn.Module):
    """x + conv(x), conv = Conv3x3-BN-LeakyReLU(0.2)-Conv3x3-BN  (vit_backbone.py:7-19)."""

    def __init__(self, channels):
        super().__init__()
        self.conv = Sequential(
            Conv2d(channels, channels, 3, 1, 1),
            BatchNorm2d(channels),
            LeakyReLU(0.2, inplace=True),
            Conv2d(channels, channels, 3, 1, 1),
            BatchNorm2d(channels),
        )

    def as_unit(self):
        c = self.conv
        return ResUnit(Unit("conv", c[0], c[1], float(c[2].negative_slope)), Unit("conv", c[3], c[4], None))

    def forward(self, x):
        return _run_units([self.as_unit()], x)


class Sequential(tnn.Sequential):
    """nn.Sequential whose forward runs maximal fusable runs as one native chain."""

    def _plan(self):
        mods = list(self)
        plan, cur, i = [], [], 0

        def flush():
            nonlocal cur
            if cur:
                plan.append(("chain", cur))
                cur = []
        while i < len(mods):
            m = mods[i]
            kind = "conv" if isinstance(m, Conv2d) else "convT" if isinstance(m, ConvTranspose2d) else \
                "linear" if isinstance(m, Linear) else None
            if kind is not None:
                if cur:
                    prev = cur[-1] if isinstance(cur[-1], Unit) else cur[-1].u2
                    if (prev.kind == "linear") != (kind == "linear"):
                        flush()
                j, bn, act = i + 1, None, None
                if j < len(mods) and isinstance(mods[j], (BatchNorm2d, BatchNorm1d)):
                    bn, j = mods[j], j + 1
                if j < len(mods) and isinstance(mods[j], LeakyReLU):
                    act, j = float(mods[j].negative_slope), j + 1
                elif j < len(mods) and isinstance(mods[j], ReLU):
                    act, j = 0.0, j + 1
                elif j < len(mods) and isinstance(mods[j], Sigmoid):
                    act, j = "sigmoid", j + 1
                cur.append(Unit(kind, m, bn, act))
                if act == "sigmoid":
                    flush()
                i = j
            elif isinstance(m, ResBlock):
                if cur and (cur[-1] if isinstance(cur[-1], Unit) else cur[-1].u2).kind == "linear":
                    flush()
                cur.append(m.as_unit())
                i += 1
            else:
                flush()
                plan.append(("mod", m))
                i += 1
        flush()
        return plan

    def forward(self, x):
        for kind, item in self._plan():
            x = _run_units(item, x) if kind == "chain" else item(x)
        return x


class MultiheadAttention(tnn.Module):
    """Self-attention with nn.MultiheadAttention's parameter layout (packed in_proj, out_proj Linear,
    batch_first=True; vit_backbone.py:28-30).  Returns (output, None) like need_weights=False —
    the reference discards the averaged weights (vit_backbone.py:43)."""

    def __init__(self, embed_dim, num_heads, dropout=0.0, batch_first=True):
        super().__init__()
        assert batch_first, "batch_first=True only (as in the reference)"
        self.embed_dim, self.num_heads, self.dropout, self.batch_first = embed_dim, num_heads, dropout, True
        self.in_proj_weight = tnn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = tnn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = Linear(embed_dim, embed_dim)
        tnn.init.xavier_uniform_(self.in_proj_weight)
        tnn.init.constant_(self.out_proj.bias, 0.0)
        self._in = _InProj(self)

    def forward(self, query, key=None, value=None, need_weights=False, **kw):
        if (key is not None and key is not query) or (value is not None and value is not query):
            if not (torch.equal(key, query) and torch.equal(value, query)):
                raise RuntimeError("only self-attention (query is key is value) is implemented")
        qkv = _run_units([Unit("linear", self._in)], query)
        p = self.dropout if self.training else 0.0
        att = F.attention_core(qkv, self.num_heads, p)
        return self.out_proj(att), None


class _InProj:
    """View of MultiheadAttention's packed projection as a Linear-like object for the chain code."""

    def __init__(self, mha):
        self._m = mha

    @property
    def weight(self):
        return self._m.in_proj_weight

    @property
    def bias(self):
        return self._m.in_proj_bias
