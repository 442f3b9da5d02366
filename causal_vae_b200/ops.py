"""Thin Python wrappers over the C ABI (one call = one enqueue on the current CUDA stream).

No autograd here: these functions take and return raw CUDA tensors.  Image activations are plain
contiguous [N, H, W, C] tensors (NHWC); matrices are [rows, cols].
"""
import os

import torch

from . import _lib as L

f32 = torch.float32


def empty(*shape, dtype=f32, like=None):
    dev = like.device if like is not None else torch.device("cuda", torch.cuda.current_device())
    return torch.empty(shape, dtype=dtype, device=dev)


# Per-step arena for the small fp64 accumulators (BatchNorm statistics, loss sums): a trainer brackets
# its forward + backward with arena_begin() / arena_end(); begin zeroes the whole arena with ONE launch
# and every zeros(..., dtype=float64) inside the bracket is a view of it (the vessel step used 41
# separate fill launches).  Outside a bracket zeros() allocates as usual.
_ARENA = {"buf": None, "off": 0, "on": False}
_ARENA_DOUBLES = 1 << 15


def arena_begin(device):
    a = _ARENA
    if a["buf"] is None or a["buf"].device != device:
        a["buf"] = torch.zeros(_ARENA_DOUBLES, dtype=torch.float64, device=device)
    else:
        fill(a["buf"].view(f32), 0.0)
    a["off"], a["on"] = 0, True


def arena_end():
    _ARENA["on"] = False


def zeros(*shape, dtype=f32, like=None):
    dev = like.device if like is not None else torch.device("cuda", torch.cuda.current_device())
    a = _ARENA
    if a["on"] and dtype == torch.float64 and a["buf"].device == dev:
        n = 1
        for d in shape:
            n *= int(d)
        if a["off"] + n <= _ARENA_DOUBLES:
            v = a["buf"][a["off"]:a["off"] + n].view(shape)
            a["off"] += (n + 1) // 2 * 2
            return v
    return torch.zeros(shape, dtype=dtype, device=dev)


class XF:
    """Deferred per-channel transform: (v-center[c])*scale[c]+shift[c] then leaky-relu(slope)."""
    __slots__ = ("scale", "shift", "slope", "center")

    def __init__(self, scale=None, shift=None, slope=1.0, center=None):
        self.scale, self.shift, self.slope, self.center = scale, shift, float(slope), center

    @property
    def identity(self):
        return self.scale is None and self.slope == 1.0

    def c(self):
        return L.xform(self.scale, self.shift, self.slope, self.center)


IDENT = XF()


# tensor-core (tcgen05, 3xTF32) path for the GEMM-shaped layers; CVAE_TC=0 forces the fp32 SIMT kernels
_TC = os.environ.get("CVAE_TC", "1") != "0"


_TILE = os.environ.get("CVAE_WGRAD_TILE", "1") != "0"   # smem-tiled fp32 weight gradient for few-channel 3x3 layers


_TC_MIN_ROWS = int(os.environ.get("CVAE_TC_MIN_ROWS", "1024"))   # below this a 128-row tile grid cannot fill the GPU


def tc_eligible(Cs, Cd, M):
    # few rows but a very wide output still make a full grid of tiles: decoder_input (512 -> 16384) at the 416 rows of a
    # counterfactual chunk ran 211 us on the fp32 gather kernel
    enough = M >= _TC_MIN_ROWS or (M >= 128 and M % 8 == 0 and M * Cd >= (1 << 21))
    return _TC and enough and bool(L.lib.cvae_tc_eligible(int(Cs), int(Cd), int(M)))


# Linear layers without an input transform: A operand pre-packed by cvae_tc_pack_rows and streamed by bulk copies (no
# producer warps).  Off by default: measured on the vessel step it neither gains nor loses (8.65 vs 8.66 ms) -- those
# launches are bound by the MMA's shared-memory operand traffic, not by staging; CVAE_APRE=1 enables it, the parity test
# tests/test_tc_gpu.py::test_linear_packed_operand always exercises it.
_APRE = os.environ.get("CVAE_APRE", "0") == "1"
FORCE_APRE = [False]
_APRE_MIN_ROWS = int(os.environ.get("CVAE_APRE_MIN_ROWS", "1024"))
_FEW = os.environ.get("CVAE_FEW", "1") != "0"   # fp32 tile kernels for image-sized 16 -> 16 stride-2 layers


def few_eligible(Cs, Cd, k, stride, pad, mode, N, Hs, Ws, Hd, Wd, epi):
    """True when cvae_conv_gather covers the layer with the few-channel tile kernels (then it must not go
    to the tensor-core entry point, and its weights are packed in the fp32 layout)."""
    return _FEW and bool(L.lib.cvae_conv_few_eligible(int(Cs), int(Cd), int(k), int(stride), int(pad), int(mode),
                                                      int(N), int(Hs), int(Ws), int(Hd), int(Wd), int(epi)))


class PackPlan:
    """Weight packings of a training step, recorded once and replayed as ONE launch (csrc/pack_batch.cu).

    While `recording`, pack_weight() packs as usual into a buffer the plan keeps and notes the job;
    finalize() uploads the job table.  Afterwards run() re-packs every recorded weight from its current
    values in one grid (call it before the forward pass, after the optimizer changed the weights) and
    pack_weight() just returns the kept buffers.  A request the plan has not seen falls back to a
    per-use pack, so a change of batch size or control flow stays correct."""

    def __init__(self):
        self.entries, self.jobs, self.recording = {}, [], True
        self.tables = {}            # phase ("fwd" | "bwd") -> (device job table, number of jobs, number of blocks)
        self.phase = "fwd"          # set to "bwd" by the chain backward: layouts first needed there can be packed
                                    # off the critical path (VesselTrainer runs them on its side stream)

    @staticmethod
    def key(w, A, A_pad, B, taps, src_bat, src_ld, tc):
        return (w.data_ptr(), int(A), int(A_pad), int(B), int(taps), int(src_bat), int(src_ld), bool(tc))

    def finalize(self):
        import numpy as np
        self.recording = False
        dt = np.dtype([("src", "<u8"), ("dst", "<u8"), ("A", "<i4"), ("A_pad", "<i4"), ("B", "<i4"), ("taps", "<i4"),
                       ("src_bat", "<i4"), ("src_ld", "<i4"), ("tc", "<i4"), ("block0", "<i4")])
        for phase in ("fwd", "bwd"):
            jobs = [j for j in self.jobs if j[3] == phase]
            if not jobs:
                continue
            arr = np.zeros(len(jobs), dtype=dt)
            b0 = 0
            for i, (w, out, (_, A, A_pad, B, taps, src_bat, src_ld, tc), _) in enumerate(jobs):
                arr[i] = (w.data_ptr(), out.data_ptr(), A, A_pad, B, taps, src_bat, src_ld, int(tc), b0)
                b0 += int(L.lib.cvae_pack_batch_blocks(A_pad, B, taps, int(tc)))
            self.tables[phase] = (torch.from_numpy(arr.view(np.uint8).copy()).to(jobs[0][1].device), len(jobs), b0)

    def run(self, phase=None):
        """re-pack the recorded layouts (all, or only those first requested in `phase`) on the current stream"""
        for ph in (("fwd", "bwd") if phase is None else (phase,)):
            if ph in self.tables:
                table, njobs, nblocks = self.tables[ph]
                L.check(L.lib.cvae_pack_batch(L.ptr(table), njobs, nblocks, L.stream()), "pack_batch")


_PLAN = [None]


def set_pack_plan(plan):
    _PLAN[0] = plan


def pack_weight(w, A, A_pad, B, taps, src_bat, src_ld, tc=False):
    plan = _PLAN[0]
    if plan is not None:
        k = PackPlan.key(w, A, A_pad, B, taps, src_bat, src_ld, tc)
        hit = plan.entries.get(k)
        if hit is not None:
            return hit
        if plan.recording:
            out = _pack_weight_now(w, A, A_pad, B, taps, src_bat, src_ld, tc)
            plan.entries[k] = out
            plan.jobs.append((w, out, k, plan.phase))
            return out
    return _pack_weight_now(w, A, A_pad, B, taps, src_bat, src_ld, tc)


def _pack_weight_now(w, A, A_pad, B, taps, src_bat, src_ld, tc=False):
    if tc:
        out = empty(L.lib.cvae_tc_pack_floats(A_pad, B, taps), like=w)
        L.check(L.lib.cvae_tc_pack_weight(L.ptr(w), L.ptr(out), A, A_pad, B, taps, int(src_bat), src_ld, L.stream()),
                "tc_pack_weight")
        return out
    out = empty(taps, A_pad, B, like=w)
    L.check(L.lib.cvae_pack_weight(L.ptr(w), L.ptr(out), A, A_pad, B, taps, int(src_bat), src_ld, L.stream()),
            "pack_weight")
    return out


def conv_gather(src, wt, bias, out_hw_c, k, stride, pad, mode, in_x=IDENT, epi=L.EPI_PLAIN, epi_ref=None,
                epi_add=None, epi_x=IDENT, stats=None, out=None, tc=False):
    N, Hs, Ws, Cs = src.shape
    Hd, Wd, Cd = out_hw_c
    dst = out if out is not None else empty(N, Hd, Wd, Cd, like=src)
    p = L.ConvParams(L.ptr(src), L.ptr(wt), L.ptr(bias), L.ptr(dst), in_x.c(), epi, L.ptr(epi_ref),
                     L.ptr(epi_add), epi_x.c(), L.ptr(stats), N, Hs, Ws, Cs, Hd, Wd, Cd, k, k, stride, pad, mode)
    M = N * Hs * Ws
    if tc and (_APRE or FORCE_APRE[0]) and k == 1 and stride == 1 and pad == 0 and in_x.identity and (Hd, Wd) == (Hs, Ws) and M % 8 == 0 \
            and M >= _APRE_MIN_ROWS and (epi == L.EPI_PLAIN or Cd <= 256) and mode == L.MODE_GATHER:
        # Linear layer on a plain matrix: split + swizzle the rows once (one small launch), then both GEMM operands are
        # streamed by bulk copies and no warp stages the A operand (csrc/conv_halo_tc.cu, a_pre mode)
        img = empty(int(L.lib.cvae_tc_pack_rows_floats(M, Cs)), like=src)
        L.check(L.lib.cvae_tc_pack_rows(L.ptr(src), L.ptr(img), M, Cs, L.stream()), "tc_pack_rows")
        L.check(L.lib.cvae_linear_tc_packed(p, L.ptr(img), L.stream()), f"linear_tc_packed M={M} K={Cs} N={Cd}")
        return dst
    fn = L.lib.cvae_conv_gather_tc if tc else L.lib.cvae_conv_gather
    L.check(fn(p, L.stream()), f"conv_gather tc={int(tc)} N={N} {Hs}x{Ws}x{Cs}->{Hd}x{Wd}x{Cd} k{k}s{stride} mode{mode}")
    return dst


_WG_DIRECT = os.environ.get("CVAE_WG_DIRECT", "1") != "0"


def conv_wgrad(ga, db, xa, xb, k, stride, pad, grad_out, ca_real=None, accumulate=False, tc=None, zeroed=False):
    """grad_out (torch layout [Cb][Ca_real][k*k]) = sum_pix xa(ga[gather]) (x) xb(db).
    zeroed: grad_out is known to hold zeros (the trainers' flat gradient buffer after zero_grad) - the tensor-core
    kernel then adds its split-K tiles straight into it (no partial buffer, no reduce launch)."""
    N, Ha, Wa, Ca = ga.shape
    _, Hq, Wq, Cb = db.shape
    taps = k * k
    rows = taps * Ca
    pixels = N * Hq * Wq
    tile_splits = L.lib.cvae_wgrad_tile_splits(pixels, Ca, Cb, k, stride, pad) if (_TILE and tc is None) else 0
    if tc is None:
        tc = _TC and bool(L.lib.cvae_wgrad_tc_eligible(pixels, rows, Cb))
    if tile_splits > 0:        # few channels x many pixels: shared-memory tiled fp32 kernel (wgrad_tile.cu)
        splits, fn, tc = tile_splits, L.lib.cvae_conv_wgrad_tile, 2
    else:
        splits = (L.lib.cvae_wgrad_tc_splits if tc else L.lib.cvae_wgrad_splits)(pixels, rows, Cb)
        fn = L.lib.cvae_conv_wgrad_tc if tc else L.lib.cvae_conv_wgrad
    if tc is True and (zeroed or accumulate) and _WG_DIRECT and grad_out.is_contiguous():
        p = L.WgradParams(L.ptr(ga), L.ptr(db), xa.c(), xb.c(), None, splits, N, Ha, Wa, Ca, Hq, Wq, Cb, k, k, stride, pad)
        L.check(L.lib.cvae_conv_wgrad_tc_direct(p, L.ptr(grad_out), Ca if ca_real is None else ca_real, L.stream()),
                f"conv_wgrad_tc_direct {Ha}x{Wa}x{Ca} / {Hq}x{Wq}x{Cb} k{k}s{stride}")
        return grad_out
    partial = empty(splits, rows, Cb, like=ga)
    p = L.WgradParams(L.ptr(ga), L.ptr(db), xa.c(), xb.c(), L.ptr(partial), splits, N, Ha, Wa, Ca, Hq, Wq, Cb,
                      k, k, stride, pad)
    L.check(fn(p, L.stream()), f"conv_wgrad tc={int(tc)} {Ha}x{Wa}x{Ca} / {Hq}x{Wq}x{Cb} k{k}s{stride}")
    L.check(L.lib.cvae_wgrad_reduce(L.ptr(partial), splits, taps, Ca, Ca if ca_real is None else ca_real, Cb,
                                    L.ptr(grad_out), int(accumulate), L.stream()), "wgrad_reduce")
    return grad_out


_HEAD_FUSED = os.environ.get("CVAE_HEAD_FUSED", "1") != "0"


def head_bwd_eligible(N, H, W, C):
    return _HEAD_FUSED and bool(L.lib.cvae_head_bwd_eligible(int(N), int(H), int(W), int(C)))


def head_bwd(g, y, xf, w, dw, stats):
    """Image-head backward in one pass (csrc/head_bwd.cu): returns dz = conv^T(g) * act'(xf(y)); writes the
    weight gradient `dw` (torch layout) and accumulates the BN-backward sums of the producer into `stats`."""
    N, H, W, C = y.shape
    dz = torch.empty_like(y)
    L.check(L.lib.cvae_head_bwd(L.ptr(g), L.ptr(y), xf.c(), L.ptr(w), L.ptr(dz), L.ptr(stats), L.ptr(dw), N, H, W, C,
                                L.stream()), f"head_bwd {N}x{H}x{W}x{C}")
    return dz


def bn_finalize(stats, C, count, bn, train_buffers=True):
    scale, shift, mean, rstd = (empty(C, like=stats) for _ in range(4))
    rm = bn.running_mean if (train_buffers and bn.track_running_stats) else None
    rv = bn.running_var if (train_buffers and bn.track_running_stats) else None
    nbt = bn.num_batches_tracked if (train_buffers and bn.track_running_stats) else None
    mom = 0.1 if bn.momentum is None else bn.momentum
    L.check(L.lib.cvae_bn_finalize(L.ptr(stats), C, float(count), L.ptr(bn.weight), L.ptr(bn.bias), bn.eps, mom,
                                   L.ptr(rm), L.ptr(rv), L.ptr(nbt), L.ptr(scale), L.ptr(shift), L.ptr(mean),
                                   L.ptr(rstd), L.stream()), "bn_finalize")
    return scale, shift, mean, rstd


def bn_eval_coeffs(bn):
    C = bn.num_features
    scale, shift = empty(C, like=bn.running_mean), empty(C, like=bn.running_mean)
    L.check(L.lib.cvae_bn_eval_coeffs(L.ptr(bn.running_mean), L.ptr(bn.running_var), L.ptr(bn.weight),
                                      L.ptr(bn.bias), bn.eps, C, L.ptr(scale), L.ptr(shift), L.stream()),
            "bn_eval_coeffs")
    return scale, shift


def col_stats(y2d_rows, C, y, stats):
    L.check(L.lib.cvae_col_stats(L.ptr(y), y2d_rows, C, L.ptr(stats), L.stream()), "col_stats")


def bn_bwd_finalize(stats, C, count, gamma, mean, rstd, want_dbias, outs=(None, None, None)):
    """outs: optional destination tensors for (dgamma, dbeta, dbias) (e.g. the parameters' .grad views)."""
    ca, cb, cc = (empty(C, like=mean) for _ in range(3))
    dg = outs[0] if outs[0] is not None else empty(C, like=mean)
    db = outs[1] if outs[1] is not None else empty(C, like=mean)
    dbias = None
    if want_dbias:
        dbias = outs[2] if outs[2] is not None else empty(C, like=mean)
    L.check(L.lib.cvae_bn_bwd_finalize(L.ptr(stats), C, float(count), L.ptr(gamma), L.ptr(mean), L.ptr(rstd),
                                       L.ptr(ca), L.ptr(cb), L.ptr(cc), L.ptr(dg), L.ptr(db), L.ptr(dbias),
                                       L.stream()), "bn_bwd_finalize")
    return ca, cb, cc, dg, db, dbias


def bn_bwd(dz, y, stats, count, gamma, mean, rstd, want_dbias, outs=(None, None, None)):
    """BatchNorm backward in one launch: returns (dy, dgamma, dbeta, dbias); falls back to finalize + apply for
    shapes the fused kernel does not cover.  outs as in bn_bwd_finalize."""
    C = y.shape[-1]
    if not L.lib.cvae_bn_bwd_fused_ok(int(C)):
        ca, cb, cc, dg, db, dbias = bn_bwd_finalize(stats, C, count, gamma, mean, rstd, want_dbias, outs)
        return bn_bwd_apply(dz, y, ca, cb, cc, mean), dg, db, dbias
    dg = outs[0] if outs[0] is not None else empty(C, like=mean)
    db = outs[1] if outs[1] is not None else empty(C, like=mean)
    dbias = None
    if want_dbias:
        dbias = outs[2] if outs[2] is not None else empty(C, like=mean)
    out = torch.empty_like(y)
    rows = y.numel() // C
    L.check(L.lib.cvae_bn_bwd(L.ptr(dz), L.ptr(y), L.ptr(stats), float(count), L.ptr(gamma), L.ptr(mean), L.ptr(rstd),
                              L.ptr(out), L.ptr(dg), L.ptr(db), L.ptr(dbias), rows, C, L.stream()), "bn_bwd")
    return out, dg, db, dbias


def affine_act(a, xa, b=None, xb=IDENT, out=None):
    C = a.shape[-1]
    rows = a.numel() // C
    out = out if out is not None else torch.empty_like(a)
    L.check(L.lib.cvae_affine_act(L.ptr(a), xa.c(), L.ptr(b), xb.c(), L.ptr(out), rows, C, L.stream()), "affine_act")
    return out


def bn_bwd_apply(dz, y, ca, cb, cc, mean, out=None):
    C = y.shape[-1]
    rows = y.numel() // C
    out = out if out is not None else torch.empty_like(y)
    L.check(L.lib.cvae_bn_bwd_apply(L.ptr(dz), L.ptr(y), L.ptr(ca), L.ptr(cb), L.ptr(cc), L.ptr(mean), L.ptr(out), rows, C,
                                    L.stream()), "bn_bwd_apply")
    return out


def dact_stats(g, ref, x, stats):
    C = ref.shape[-1]
    rows = ref.numel() // C
    dz = torch.empty_like(ref)
    L.check(L.lib.cvae_dact_stats(L.ptr(g), L.ptr(ref), x.c(), L.ptr(dz), L.ptr(stats), rows, C, L.stream()),
            "dact_stats")
    return dz


def col_sum(x, C, out=None, accumulate=False):
    rows = x.numel() // C
    out = out if out is not None else empty(C, like=x)
    L.check(L.lib.cvae_col_sum(L.ptr(x), rows, C, L.ptr(out), int(accumulate), L.stream()), "col_sum")
    return out


def layernorm_fwd(x, gamma, beta, rows, D, row_stride, eps):
    y = empty(rows, D, like=x)
    mean, rstd = empty(rows, like=x), empty(rows, like=x)
    L.check(L.lib.cvae_layernorm_fwd(L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(y), L.ptr(mean), L.ptr(rstd), rows,
                                     D, row_stride, eps, L.stream()), "layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, rows, D, x_row_stride, dx, dx_row_stride, accumulate_dx, dgamma, dbeta):
    L.check(L.lib.cvae_layernorm_bwd(L.ptr(dy), L.ptr(x), L.ptr(gamma), L.ptr(mean), L.ptr(rstd), L.ptr(dx),
                                     L.ptr(dgamma), L.ptr(dbeta), rows, D, x_row_stride, dx_row_stride,
                                     int(accumulate_dx), L.stream()), "layernorm_bwd")


def attention_fwd(qkv, B, S, H, d, p, seed, offset, counter=None):
    out = empty(B, S, H * d, like=qkv)
    probs = empty(B, H, S, S, like=qkv)
    L.check(L.lib.cvae_attention_fwd(L.ptr(qkv), L.ptr(out), L.ptr(probs), B, S, H, d, p, seed, offset, L.ptr(counter),
                                     L.stream()),
            f"attention_fwd S={S} d={d}")
    return out, probs


def attention_bwd(qkv, probs, dout, B, S, H, d, p, seed, offset, counter=None):
    dqkv = torch.empty_like(qkv)
    nws = int(L.lib.cvae_attention_ws_bytes(B, S, H, d))
    if nws:      # S > 128: strip kernels with a per-row workspace (csrc/attention_long.cu)
        ws = torch.empty(nws // 4, dtype=torch.float32, device=qkv.device)
        L.check(L.lib.cvae_attention_bwd_ws(L.ptr(qkv), L.ptr(probs), L.ptr(dout), L.ptr(dqkv), L.ptr(ws), nws, B, S, H, d, p,
                                            L.stream()), f"attention_bwd_ws S={S}")
        return dqkv
    L.check(L.lib.cvae_attention_bwd(L.ptr(qkv), L.ptr(probs), L.ptr(dout), L.ptr(dqkv), B, S, H, d, p, seed, offset,
                                     L.ptr(counter), L.stream()), "attention_bwd")
    return dqkv


def act_fwd(x, act, slope=0.0):
    y = torch.empty_like(x)
    L.check(L.lib.cvae_act_fwd(L.ptr(x), L.ptr(y), x.numel(), act, slope, L.stream()), "act_fwd")
    return y


def act_bwd(dy, x, act, slope=0.0):
    dx = torch.empty_like(x)
    L.check(L.lib.cvae_act_bwd(L.ptr(dy), L.ptr(x), L.ptr(dx), x.numel(), act, slope, L.stream()), "act_bwd")
    return dx


def add(a, b, out=None):
    out = out if out is not None else torch.empty_like(a)
    L.check(L.lib.cvae_add(L.ptr(a), L.ptr(b), L.ptr(out), a.numel(), L.stream()), "add")
    return out


def dropout(x, p, seed, offset, counter=None):
    y = torch.empty_like(x)
    L.check(L.lib.cvae_dropout(L.ptr(x), L.ptr(y), x.numel(), p, seed, offset, L.ptr(counter), L.stream()), "dropout")
    return y


def transpose_bc(src, B, rows, cols):
    """[B, rows, cols] -> [B, cols, rows]"""
    dst = empty(B, cols, rows, like=src)
    L.check(L.lib.cvae_transpose_bc(L.ptr(src), L.ptr(dst), B, rows, cols, L.stream()), "transpose_bc")
    return dst


def copy_cols(src, src_ld, src_col0, dst, dst_ld, dst_col0, rows, w, accumulate=False):
    L.check(L.lib.cvae_copy_cols(L.ptr(src), src_ld, src_col0, L.ptr(dst), dst_ld, dst_col0, rows, w,
                                 int(accumulate), L.stream()), "copy_cols")


def fill(t, v):
    L.check(L.lib.cvae_fill(L.ptr(t), t.numel(), float(v), L.stream()), "fill")
    return t


def finish_scalar(acc, mul=1.0):
    out = torch.empty((), dtype=f32, device=acc.device)
    L.check(L.lib.cvae_finish_scalar(L.ptr(acc), float(mul), L.ptr(out), L.stream()), "finish_scalar")
    return out
