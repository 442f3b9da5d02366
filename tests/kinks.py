"""Kink-conditioned gradient parity (test infrastructure).

Whole-network gradients are discontinuous where a pre-activation crosses a LeakyReLU / ReLU / |.| kink.  Two correct
fp32 implementations whose forward values differ by one rounding pick different sides for a unit sitting within
rounding distance of 0, and one such unit moves whole gradient tensors by 1e-3..1e-1 of their maximum (every loss of
the reference is sum-reduced, gradient sums cancel to ~sqrt(n) of their terms).  So the whole-model gradient check
is split into statements that are each tight:

  1. forward values match the fp64 oracle at the north-star forward tolerance (checked by the callers);
  2. the derivative sides the native kernels took (read back through chain.TRACE) differ from the fp64 oracle's
     only on units whose oracle pre-activation lies INSIDE the forward-tolerance band around the kink
     (|z| <= band * max|z| of that layer) -- i.e. only where the reference's own fp32 arithmetic may differ too;
  3. evaluated with the SAME sides (oracle KINKS["masks"]), the gradients agree to max(1e-4, 4 x the oracle's own
     fp32-vs-fp64 discrepancy under those sides) of each tensor's max |g| -- no perturbation-calibrated slack.
"""
import torch

from causal_vae_b200 import chain
from oracle import cvae_oracle as O


class NativeTrace:
    """with NativeTrace(model) as tr: native forward ...  -> tr.masks: {state_dict prefix: bool tensor (NCHW / [B, C])}"""

    def __init__(self, *models, prefix=""):
        self.names = {}
        for mdl in models:
            self.names.update({id(m): prefix + n for n, m in mdl.named_modules()})
        self.masks = {}

    def __enter__(self):
        chain.TRACE[0] = {}
        return self

    def __exit__(self, *exc):
        rec, chain.TRACE[0] = chain.TRACE[0], None
        for mid, (y, xf) in rec.items():
            name = self.names.get(mid)
            if name is None:
                continue
            v = y.detach()
            if xf.scale is not None:
                # the kernels evaluate fmaf(v - center, scale, shift): (v - center) rounded to fp32, then one fused
                # multiply-add.  In fp64 the product of two fp32 numbers is exact and the sum is rounded once, so the
                # SIGN below is the sign the fp32 fma sees.
                d = v - xf.center if xf.center is not None else v
                z = d.double() * xf.scale.double() + xf.shift.double()
            else:
                z = v.double()
            m = z > 0
            N, H, W, C = m.shape
            m = m.reshape(N, C) if H * W == 1 else m.permute(0, 3, 1, 2)
            self.masks[name] = m.contiguous().cpu()
        return False

    def add_sign(self, key, t):
        """a |.| kink (vessel sparsity term): side = sign of the native tensor"""
        self.masks[key] = (t.detach() > 0).cpu()


def oracle_sides(run64):
    """run64(): an fp64 oracle forward.  Returns {key: fp64 pre-activation} of every activation site it passed."""
    O.KINKS["log"] = {}
    try:
        with torch.no_grad():
            run64()
        return O.KINKS["log"]
    finally:
        O.KINKS["log"] = None


def check_sides(masks, pre64, band):
    """Statement 2: every unit whose native side differs from the fp64 oracle's lies inside the band.  Returns
    (number of differing units, number of units, worst |z| / max|z| among the differing ones)."""
    flips = total = 0
    worst = 0.0
    missing = [k for k in pre64 if k not in masks]
    assert not missing, f"activation sites the native trace did not report: {missing}"
    for k, z in pre64.items():
        m = masks[k].reshape(z.shape)
        diff = m != (z > 0)
        total += z.numel()
        n = int(diff.sum())
        if n:
            flips += n
            r = float(z[diff].abs().max() / z.abs().max())
            worst = max(worst, r)
            assert r <= band, f"{k}: a unit {r:.2e} of max|z| away from the kink took the other side (band {band:.0e})"
    return flips, total, worst


class with_masks:
    """with with_masks(masks): oracle forward/backward evaluated with the given derivative sides"""

    def __init__(self, masks):
        self.masks = masks

    def __enter__(self):
        O.KINKS["masks"] = self.masks

    def __exit__(self, *exc):
        O.KINKS["masks"] = None
        return False


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def check_grads(native, g64, g32, floor=1e-4, what="", statistical=False):
    """Statement 3.  native / g64 / g32: {name: gradient}; tolerance per tensor max(floor, 4 x rel(g32, g64)).
    statistical=True (only for shapes whose batch-of-4 BatchNorm1d amplifies last-bit differences ~10^3 x, see the caller):
    the tolerance is a multiple of ONE realisation of the reference's own fp32 rounding noise, and another equally valid
    fp32 evaluation order (ours varies with the order of fp32 atomics) lands around it -- no tensor beyond 1.5 x, at most
    5 % of the tensors beyond 1 x."""
    worst = []
    for k, g in g64.items():
        if g is None:
            continue
        noise = rel(g32[k], g)
        if noise > 1.0:          # a bias in front of a BatchNorm: its true gradient is 0, fp32 gives rounding noise
            continue
        e = rel(native[k], g)
        worst.append((e / max(floor, 4 * noise), k, e, noise))
    worst.sort(reverse=True)
    top = [(round(r, 2), k, f"{e:.1e}", f"{n:.1e}") for r, k, e, n in worst[:8]]
    print(f"kink-conditioned gradient error / tolerance {what}:", top)
    if statistical:
        over = [w for w in worst if w[0] > 1.0]
        assert worst and worst[0][0] <= 1.5 and len(over) <= max(1, len(worst) // 20), top
    else:
        assert worst and worst[0][0] <= 1.0, top
    return worst
