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


class CounterfactualEngine:
    """High-throughput form of counterfactual_sweep for a frozen model: the whole sweep of one chunk of
    sources (base decode, do() scatter, decode of chunk*K rows, per-image L2) is captured once in a CUDA
    graph over static (m, z) buffers, and the weight layouts are packed once (ops.PackPlan; weights do not
    change in eval mode - call refresh() after loading new weights).  The eager sweep spends a third of
    its time in ~10^3 host-side enqueues per chunk; a replay is one launch.

        eng = CounterfactualEngine(model, chunk=32, delta=5.0)
        for i in range(0, S, 32):
            l2 = eng(m[i:i+32], z[i:i+32])        # [32, K]; valid until the next call
    """

    def __init__(self, model, chunk, delta=5.0, value=None):
        self.model, self.chunk, self.delta, self.value = model, int(chunk), delta, value
        model.eval()
        p0 = next(model.parameters())
        self.m = torch.zeros(self.chunk, model.m_dim, device=p0.device)
        self.z = torch.zeros(self.chunk, model.my_z_dim, device=p0.device)
        self.plan = ops.PackPlan()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):
                self._sweep()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.l2 = self._sweep()

    def _sweep(self):
        ops.set_pack_plan(self.plan)
        try:
            l2, _, _ = counterfactual_sweep(self.model, self.m, self.z, delta=self.delta, value=self.value)
        finally:
            ops.set_pack_plan(None)
        if self.plan.recording:
            self.plan.finalize()
        return l2

    def refresh(self):
        """re-pack every weight layout from the model's current parameters (one launch)"""
        self.plan.run()

    @torch.no_grad()
    def __call__(self, m, z):
        if m.shape[0] != self.chunk:
            raise RuntimeError(f"CounterfactualEngine was captured for chunks of {self.chunk} sources, got {m.shape[0]}")
        self.m.copy_(m, non_blocking=True)
        self.z.copy_(z, non_blocking=True)
        self.graph.replay()
        return self.l2
