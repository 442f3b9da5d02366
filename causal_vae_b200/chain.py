"""Fused layer chains: Conv2d / ConvTranspose2d / Linear (+ BatchNorm + LeakyReLU/ReLU/Sigmoid) and
ResBlocks executed as one autograd node with a hand-written backward.

Why a chain and not per-layer ops: training-mode BatchNorm needs the whole batch before the
activation can be applied, so each layer writes its raw (pre-BN) output once together with the batch
statistics (conv epilogue), and the normalise + activation is applied by the *consumer* while it
loads its operand.  Backward mirrors this: the consumer's input-gradient kernel multiplies by the
activation derivative and accumulates the BN-backward sums in its epilogue.

A Linear is a 1x1 conv on a [B, 1, 1, K] tensor, so MLP adapters (Linear -> BatchNorm1d ->
LeakyReLU -> Linear; vessel_analysis/00_core/models.py:225-240) run through the same code.
"""
import torch

from . import _lib as L
from . import ops
from .ops import IDENT, XF


class Unit:
    """conv-like layer + optional BN + optional activation."""

    def __init__(self, kind, mod, bn=None, act=None):
        self.kind, self.mod, self.bn, self.act = kind, mod, bn, act   # act: None | float slope | "sigmoid"
        if kind == "linear":
            self.k, self.stride, self.pad, self.opad = 1, 1, 0, 0
        else:
            ks, st, pd = mod.kernel_size, mod.stride, mod.padding
            assert ks[0] == ks[1] and st[0] == st[1] and pd[0] == pd[1], "square kernels / strides only"
            assert mod.dilation == (1, 1) and mod.groups == 1, "dilation / groups unsupported"
            self.k, self.stride, self.pad = ks[0], st[0], pd[0]
            self.opad = mod.output_padding[0] if kind == "convT" else 0

    def params(self):
        p = [self.mod.weight, self.mod.bias]
        if self.bn is not None:
            p += [self.bn.weight, self.bn.bias]
        return p

    def out_geom(self, Hs, Ws):
        w = self.mod.weight
        if self.kind == "conv":
            return ((Hs + 2 * self.pad - self.k) // self.stride + 1, (Ws + 2 * self.pad - self.k) // self.stride + 1,
                    w.shape[0])
        if self.kind == "convT":
            return ((Hs - 1) * self.stride - 2 * self.pad + self.k + self.opad,
                    (Ws - 1) * self.stride - 2 * self.pad + self.k + self.opad, w.shape[1])
        return (1, 1, w.shape[0])


class ResUnit:
    """x + BN(conv(LReLU(BN(conv x))))  (vit_backbone.py:7-19)."""

    def __init__(self, u1, u2):
        self.u1, self.u2 = u1, u2

    def params(self):
        return self.u1.params() + self.u2.params()


class State:
    __slots__ = ("t", "x")

    def __init__(self, t, x=IDENT):
        self.t, self.x = t, x


class _Rec:
    pass


# Test hook.  TRACE[0] = {} makes every activation site of the following forward passes record
# id(module producing the pre-activation) -> (raw tensor, XF): the sign of fmaf(raw - center, scale, shift) is the
# LeakyReLU / ReLU side the kernels take for that unit, in forward and in backward (tests/kinks.py turns it into the
# derivative masks the checker is evaluated with).  None (the default) costs nothing.
TRACE = [None]


# Trainers that own zero_grad() (FlatParams) and use every parameter exactly once per backward set
# DIRECT_GRADS[0] = True: parameter gradients are then written straight into `p.grad` by the kernels
# that produce them and autograd receives None, which removes one accumulate launch (and one
# temporary) per parameter tensor - 180 launches per vessel step.
DIRECT_GRADS = [False]


class direct_grads:
    """with direct_grads(): loss.backward()  - see DIRECT_GRADS."""

    def __enter__(self):
        self.prev = DIRECT_GRADS[0]
        DIRECT_GRADS[0] = True

    def __exit__(self, *exc):
        DIRECT_GRADS[0] = self.prev
        return False


# Weight-gradient kernels on a side stream.  In backward a layer's weight gradient (and bias column sum) and
# its input gradient are independent; the ViT / adapter layers launch grids of 66-132 CTAs on 148 SMs, so
# running the two families concurrently fills the idle SMs (and overlaps the tails of the big layers).
# Only used together with direct_grads (the kernels then write into the flat gradient buffer and autograd
# never touches the result); the trainer joins the side stream before the optimizer.
SIDE = {"stream": None, "keep": []}


class side_wgrad:
    """with direct_grads(), side_wgrad(stream): loss.backward()"""

    def __init__(self, stream, join_first=False):
        self.stream, self.join_first = stream, join_first

    def __enter__(self):
        if self.stream is not None and self.join_first:
            # work the trainer put on the side stream during forward (gradient zeroing, backward-only weight
            # layouts) is complete before the first backward kernel
            torch.cuda.current_stream().wait_stream(self.stream)
        SIDE["stream"], SIDE["keep"] = self.stream, []

    def __exit__(self, *exc):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)     # join before anything reads the gradients
        # operands of the side-stream kernels were kept alive until here: blocks freed now can only be reused by
        # work that is ordered after the join
        SIDE["stream"], SIDE["keep"] = None, []
        return False


def direct_ok(p):
    return DIRECT_GRADS[0] and p is not None and p.requires_grad and p.grad is not None and p.grad.is_contiguous()


def _pack_fwd(u, Cs_phys, tc=False):
    w = u.mod.weight
    taps = u.k * u.k
    if u.kind == "conv":      # [Cout][Cin][tap] -> [tap][Cin][Cout]
        return ops.pack_weight(w, w.shape[1], Cs_phys, w.shape[0], taps, True, w.shape[1], tc=tc)
    if u.kind == "convT":     # [Cin][Cout][tap] -> [tap][Cin][Cout]
        return ops.pack_weight(w, w.shape[0], Cs_phys, w.shape[1], taps, False, w.shape[1], tc=tc)
    return ops.pack_weight(w, w.shape[1], Cs_phys, w.shape[0], 1, True, w.shape[1], tc=tc)  # [out][in] -> [in_pad][out]


def _pack_dgrad(u, grad_cols, tc=False, dy_cols=None):
    """weights for the input-gradient: [tap][C_of_dy][C_of_dx]."""
    w = u.mod.weight
    taps = u.k * u.k
    if u.kind == "linear" and dy_cols is not None and dy_cols != w.shape[0]:
        return ops.pack_weight(w, w.shape[0], dy_cols, grad_cols, 1, False, w.shape[1], tc=tc)  # zero rows for padded dy
    if u.kind == "conv":      # [Cout][Cin][tap] -> [tap][Cout][Cin]
        return ops.pack_weight(w, w.shape[0], w.shape[0], w.shape[1], taps, False, w.shape[1], tc=tc)
    if u.kind == "convT":     # [Cin][Cout][tap] -> [tap][Cout][Cin]
        return ops.pack_weight(w, w.shape[1], w.shape[1], w.shape[0], taps, True, w.shape[1], tc=tc)
    if grad_cols == w.shape[1] and not tc:
        return w                                  # [out][in] already is [C_of_dy][C_of_dx]
    return ops.pack_weight(w, w.shape[0], w.shape[0], grad_cols, 1, False, w.shape[1], tc=tc)  # [out][in] -> [out][cols]


def _min_phase_rows(N, Hd, Wd, stride, scatter):
    """rows (output pixels) of the smallest GEMM the layer decomposes into."""
    return N * (Hd // stride) * (Wd // stride) if scatter else N * Hd * Wd


def _unit_fwd(u, S, train, keep):
    N, Hs, Ws, Cs = S.t.shape
    Hd, Wd, Cd = u.out_geom(Hs, Ws)
    mode = L.MODE_SCATTER if u.kind == "convT" else L.MODE_GATHER
    bn_train = u.bn is not None and (u.bn.training or not u.bn.track_running_stats)
    tc = ops.tc_eligible(Cs, Cd, _min_phase_rows(N, Hd, Wd, u.stride, mode == L.MODE_SCATTER)) and not \
        ops.few_eligible(Cs, Cd, u.k, u.stride, u.pad, mode, N, Hs, Ws, Hd, Wd, L.EPI_STATS if bn_train else L.EPI_PLAIN)
    wt = _pack_fwd(u, Cs, tc)
    stats = ops.zeros(2 * Cd, dtype=torch.float64, like=S.t) if bn_train else None
    y = ops.conv_gather(S.t, wt, u.mod.bias, (Hd, Wd, Cd), u.k, u.stride, u.pad, mode, in_x=S.x,
                        epi=L.EPI_STATS if bn_train else L.EPI_PLAIN, stats=stats, tc=tc)
    rec = _Rec()
    rec.S_in, rec.y, rec.mean, rec.rstd, rec.sig = S, y, None, None, None
    slope = u.act if isinstance(u.act, float) else 1.0
    if u.bn is not None:
        if bn_train:
            scale, shift, rec.mean, rec.rstd = ops.bn_finalize(stats, Cd, N * Hd * Wd, u.bn,
                                                               train_buffers=u.bn.training)
        else:
            scale, shift = ops.bn_eval_coeffs(u.bn)
        out = State(y, XF(scale, shift, slope, rec.mean if bn_train else u.bn.running_mean))
    elif isinstance(u.act, float):
        out = State(y, XF(None, None, slope))
    elif u.act == "sigmoid":
        rec.sig = ops.act_fwd(y, L.ACT_SIGMOID)
        out = State(rec.sig)
    else:
        out = State(y)
    rec.S_out = out
    if TRACE[0] is not None and isinstance(u.act, float):
        TRACE[0][id(u.bn if u.bn is not None else u.mod)] = (y, out.x)
    return out, (rec if keep else None)


def chain_forward(chain, x, train, keep):
    """x: [N,H,W,C] plain tensor.  Returns (materialised output, records)."""
    S = State(x)
    recs = []
    for e in chain:
        if isinstance(e, Unit):
            S, r = _unit_fwd(e, S, train, keep)
            recs.append(r)
        else:
            S_in = S
            S1, r1 = _unit_fwd(e.u1, S_in, train, keep)
            S2, r2 = _unit_fwd(e.u2, S1, train, keep)
            r = ops.affine_act(S2.t, S2.x, S_in.t, S_in.x)
            S = State(r)
            recs.append((r1, r2))
    final = S
    out = S.t if S.x.identity else ops.affine_act(S.t, S.x)
    return out, recs, final


def _unit_bwd(u, rec, dz, stats, prev_entry, prev_stats, need_dx, add, grads, grad_cols=None):
    """dz: gradient w.r.t. the BN output / pre-activation of this unit's raw output y (the
    activation derivative has already been applied); stats = (sum dz, sum dz*y) when u.bn.
    prev_entry: State describing how the input value derives from a raw tensor (for the DACT
    epilogue of the input-gradient kernel).  Appends this unit's parameter grads to `grads`."""
    y = rec.y
    N, Hd, Wd, Cd = y.shape
    S_in = rec.S_in
    _, Hs, Ws, Cs = S_in.t.shape
    w = u.mod.weight
    has_bias = u.mod.bias is not None
    if u.bn is not None:
        if rec.mean is None:
            raise NotImplementedError("backward through eval-mode BatchNorm is not supported")
        outs = tuple(p.grad if direct_ok(p) else None for p in (u.bn.weight, u.bn.bias, u.mod.bias))
        dy, dgamma, dbeta, dbias = ops.bn_bwd(dz, y, stats, N * Hd * Wd, u.bn.weight, rec.mean, rec.rstd, has_bias, outs)
        dgamma, dbeta = (None if o is not None else g for o, g in zip(outs[:2], (dgamma, dbeta)))
        dbias = None if outs[2] is not None else dbias
    else:
        dy = dz
        dgamma = dbeta = None
        bias_direct = has_bias and direct_ok(u.mod.bias)
        dbias = ops.col_sum(dy, Cd) if (has_bias and not bias_direct) else None
    w_direct = direct_ok(w)
    gw = w.grad if w_direct else torch.empty_like(w)
    side = SIDE["stream"] if w_direct else None
    bias_direct = u.bn is None and has_bias and direct_ok(u.mod.bias)
    if bias_direct and side is None:
        ops.col_sum(dy, Cd, out=u.mod.bias.grad)
    if (u.kind == "conv" and u.k == 3 and u.stride == 1 and u.pad == 1 and Cd == 1 and u.bn is None and add is None
            and need_dx and grad_cols is None and prev_entry is S_in and not S_in.x.identity and w.is_contiguous()
            and dy.is_contiguous() and ops.head_bwd_eligible(N, Hs, Ws, Cs)):
        # image head: weight gradient and input gradient share one pass over the (large) raw input
        if bias_direct and side is not None:
            ops.col_sum(dy, Cd, out=u.mod.bias.grad)
        d_in = ops.head_bwd(dy, S_in.t, S_in.x, w, gw, prev_stats)
        grads.append((None if w_direct else gw, dbias, dgamma, dbeta))
        return d_in

    def weight_side():
        if bias_direct and side is not None:
            ops.col_sum(dy, Cd, out=u.mod.bias.grad)
        if u.kind == "convT":
            ops.conv_wgrad(dy, S_in.t, IDENT, S_in.x, u.k, u.stride, u.pad, gw, zeroed=w_direct)
        else:
            ca_real = w.shape[1] if u.kind == "linear" else None
            ops.conv_wgrad(S_in.t, dy, S_in.x, IDENT, u.k, u.stride, u.pad, gw, ca_real=ca_real, zeroed=w_direct)
    if side is not None:
        side.wait_stream(torch.cuda.current_stream())     # dy (and the statistics it depends on) are complete
        with torch.cuda.stream(side):
            weight_side()
        SIDE["keep"].append((dy, rec, gw))                # operands stay allocated until the join
    else:
        weight_side()
    grads.append((None if w_direct else gw, dbias, dgamma, dbeta))
    if not need_dx:
        return None
    cols = Cs if grad_cols is None else grad_cols
    mode = L.MODE_GATHER if u.kind == "convT" else L.MODE_SCATTER
    dy_cols = None
    if u.kind == "linear" and (Cd % 4 != 0 or Cd < 8):
        # narrow heads (e.g. 10 treatment logits, 4 concepts): the gather kernels read 128-bit channel
        # vectors of at least 8 channels, so the gradient rows are zero-padded (matching zero weight rows)
        dy_cols = max(8, (Cd + 3) // 4 * 4)
        dyp = ops.zeros(N, 1, 1, dy_cols, like=dy)
        ops.copy_cols(dy, Cd, 0, dyp, dy_cols, 0, N, Cd)
        dy, Cd = dyp, dy_cols
    epi_b = L.EPI_DACT if ((prev_entry is not None and not prev_entry.x.identity) or add is not None) else L.EPI_PLAIN
    tc = ops.tc_eligible(Cd, cols, _min_phase_rows(N, Hs, Ws, u.stride, mode == L.MODE_SCATTER)) and not \
        ops.few_eligible(Cd, cols, u.k, u.stride, u.pad, mode, N, Hd, Wd, Hs, Ws, epi_b)
    wt = _pack_dgrad(u, cols, tc, dy_cols)
    if prev_entry is not None and not prev_entry.x.identity:
        return ops.conv_gather(dy, wt, None, (Hs, Ws, cols), u.k, u.stride, u.pad, mode, epi=L.EPI_DACT,
                               epi_ref=prev_entry.t, epi_add=add, epi_x=prev_entry.x, stats=prev_stats, tc=tc)
    if add is not None:
        return ops.conv_gather(dy, wt, None, (Hs, Ws, cols), u.k, u.stride, u.pad, mode, epi=L.EPI_DACT,
                               epi_ref=add, epi_add=add, epi_x=IDENT, stats=None, tc=tc)
    return ops.conv_gather(dy, wt, None, (Hs, Ws, cols), u.k, u.stride, u.pad, mode, tc=tc)


def _entry(chain, recs, i):
    """(State whose raw tensor / xform define the output value of element i, its BN unit or None)."""
    e, r = chain[i], recs[i]
    if isinstance(e, Unit):
        return r.S_out, e
    # ResUnit: r = T2(y2) + skip; d r / d T2 = 1 -> slope-1 entry on y2 so the consumer gathers BN2's sums
    r2 = r[1]
    return State(r2.y, XF(r2.S_out.x.scale, r2.S_out.x.shift, 1.0, r2.S_out.x.center)), e.u2


def chain_backward(chain, recs, final, gout, need_dx, grad_cols=None):
    """Returns (dx or None, [per-unit (dW, dbias, dgamma, dbeta)] in chain order)."""
    n = len(chain)
    per_elem = [None] * n
    # gradient entering the last element
    S_last, bn_last = _entry(chain, recs, n - 1)
    like = gout
    if not S_last.x.identity:
        stats = ops.zeros(2 * S_last.t.shape[-1], dtype=torch.float64, like=like)
        dz = ops.dact_stats(gout, S_last.t, S_last.x, stats)
    else:
        stats, dz = None, gout
        if isinstance(chain[-1], Unit) and chain[-1].act == "sigmoid":
            dz = ops.act_bwd(gout, recs[-1].sig, L.ACT_SIGMOID)
    dx = None
    for i in range(n - 1, -1, -1):
        e, r = chain[i], recs[i]
        if i > 0:
            prev_entry, _ = _entry(chain, recs, i - 1)
            prev_has_x = not prev_entry.x.identity
            prev_stats = ops.zeros(2 * prev_entry.t.shape[-1], dtype=torch.float64, like=like) if prev_has_x else None
            need = True
        else:
            prev_entry, prev_stats, need = None, None, need_dx
        g = []
        if isinstance(e, Unit):
            if i > 0 and isinstance(chain[i - 1], Unit) and chain[i - 1].act == "sigmoid":
                raise NotImplementedError("sigmoid is only supported as the last activation of a chain")
            d_in = _unit_bwd(e, r, dz, stats, prev_entry, prev_stats, need, None, g, grad_cols if i == 0 else None)
        else:
            r1, r2 = r
            g_r = dz                                   # gradient w.r.t. the residual sum
            S1, _ = r1.S_out, None
            stats1 = ops.zeros(2 * r1.y.shape[-1], dtype=torch.float64, like=like)
            g2 = []
            dz1 = _unit_bwd(e.u2, r2, g_r, stats, State(r1.y, r1.S_out.x), stats1, True, None, g2)
            g1 = []
            if i == 0 and prev_entry is None:
                d_in = _unit_bwd(e.u1, r1, dz1, stats1, None, None, need, g_r if need else None, g1)
            else:
                d_in = _unit_bwd(e.u1, r1, dz1, stats1, prev_entry, prev_stats, need, g_r, g1)
            g = g1 + g2
        per_elem[i] = g
        dz, stats = d_in, prev_stats
        if i == 0:
            dx = d_in
    flat = [t for g in per_elem for t in g]
    return dx, flat


class ChainFn(torch.autograd.Function):
    """autograd node for a whole chain.  args: x (NHWC), chain, train flag, grad_cols, *params."""

    @staticmethod
    def forward(ctx, x, chain, train, grad_cols, *params):
        keep = any(ctx.needs_input_grad)
        out, recs, final = chain_forward(chain, x, train, keep)
        ctx.chain, ctx.recs, ctx.final, ctx.grad_cols = chain, recs, final, grad_cols
        ctx.need_dx = ctx.needs_input_grad[0]
        ctx.x_cols = x.shape[-1]
        return out

    @staticmethod
    def backward(ctx, gout):
        gout = gout.contiguous()
        plan = ops._PLAN[0]
        if plan is not None:
            plan.phase = "bwd"
        try:
            dx, flat = chain_backward(ctx.chain, ctx.recs, ctx.final, gout, ctx.need_dx, ctx.grad_cols)
        finally:
            if plan is not None:
                plan.phase = "fwd"
        ctx.recs = ctx.final = None
        if dx is not None and ctx.grad_cols is not None and ctx.grad_cols != ctx.x_cols:
            full = ops.zeros(*dx.shape[:-1], ctx.x_cols, like=dx)
            rows = dx.numel() // ctx.grad_cols
            ops.copy_cols(dx, ctx.grad_cols, 0, full, ctx.x_cols, 0, rows, ctx.grad_cols)
            dx = full
        grads = []
        units = []
        for e in ctx.chain:
            units += [e] if isinstance(e, Unit) else [e.u1, e.u2]
        for u, (gw, gb, gg, gbeta) in zip(units, flat):
            grads += [gw, gb if u.mod.bias is not None else None]
            if u.bn is not None:
                grads += [gg, gbeta]
        return (dx, None, None, None, *grads)


def chain_params(chain):
    ps = []
    for e in chain:
        ps += e.params()
    return ps


def run_chain(chain, x_nhwc, train, grad_cols=None):
    return ChainFn.apply(x_nhwc, chain, train, grad_cols, *chain_params(chain))
