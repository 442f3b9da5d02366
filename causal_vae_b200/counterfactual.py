"""Batched do()-interventions and counterfactual decoding.

The reference performs interventions by hand in scripts: abduct z from (x, m, t), clone m, overwrite
or shift one concept, decode cat[m', z] (vessel_analysis/04_generate_counterfactual/
generate_counterfactual.py:50-99; analyze_vessel.py:90-115; visualize_diff.py:43-51).  Here the same
arithmetic is batched over all K concepts of S source samples: the do() scatter is one kernel, the
decode is the eval-mode decoder chain (BatchNorm folded into the consumer's operand load), and the
per-image effect size ||x_cf - x_base||_2 is reduced on device so the images need not be kept."""
import torch

from . import _lib as L
from . import ops


def do_expand(m, z, delta=None, value=None):
    """rows (s*K + k) = cat(do_k(m[s]), z[s]) with do_k: m_k += delta  or  m_k := value.  [S*K, K+Z]."""
    S, K = m.shape
    Z = z.shape[1]
    out = ops.empty(S * K, K + Z, like=m)
    setv = value is not None
    v = float(value if setv else (delta if delta is not None else 0.0))
    L.check(L.lib.cvae_do_expand(L.ptr(m.contiguous()), L.ptr(z.contiguous()), L.ptr(out), S, K, Z, int(setv), v,
                                 L.stream()), "do_expand")
    return out


def rowdiff_l2(a, b, group):
    """out[r] = || a[r] - b[r // group] ||_2 over flattened rows."""
    rows = a.shape[0]
    rowlen = a.numel() // rows
    out = ops.empty(rows, like=a)
    L.check(L.lib.cvae_rowdiff_l2(L.ptr(a.contiguous()), L.ptr(b.contiguous()), L.ptr(out), rows, rowlen, group,
                                  L.stream()), "rowdiff_l2")
    return out


@torch.no_grad()
def abduct(model, x, m, t, eps=None, use_mean=False):
    """z for each sample: reparameterised with `eps` (generate_counterfactual.py:54-55) or z = mu."""
    mu, logvar, z = model.encode(x, m, t, eps)
    return mu if use_mean else z


@torch.no_grad()
def counterfactual_sweep(model, m, z, delta=5.0, value=None, return_images=False):
    """All K single-concept interventions for every source: decode(cat[do_k(m), z]).
    Returns (l2 effect per (source, concept) [S, K], images [S*K,1,H,W] or None)."""
    S, K = m.shape
    base = model.decode(m, z)
    rows = do_expand(m, z, delta=delta, value=value)
    x_cf = model.backbone.decode(model.dec_adapter(rows))
    l2 = rowdiff_l2(x_cf, base, K).view(S, K)
    return l2, (x_cf if return_images else None), base
