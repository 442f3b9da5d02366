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
    if hasattr(model, "dec_adapter"):                      # the ViT / CNN switch of analyze_vessel.py:93-98
        x_cf = model.backbone.decode(model.dec_adapter(rows))
    else:
        x_cf = model.dec_conv(model.dec_fc(rows).view(-1, 512, *model.GRID))
    l2 = rowdiff_l2(x_cf, base, K).view(S, K)
    return l2, (x_cf if return_images else None), base


@torch.no_grad()
def feature_importance(model, m, z, delta=1.0):
    """Visual sensitivity of every concept: mean over samples of ||decode(m + delta*e_k, z) - decode(m, z)||_2
    (vessel_analysis/03_evaluate_vessel/analyze_vessel.py:68-115, which perturbs by +1 sigma on random (m, z)).
    One batched sweep instead of K decodes; returns [K]."""
    l2, _, _ = counterfactual_sweep(model, m, z, delta=delta)
    return l2.mean(dim=0)


@torch.no_grad()
def mediation_decomposition(model, m_a, z_a, m_b, z_b):
    """How much of the visual difference between (m_a, z_a) and (m_b, z_b) is carried by the measured concepts M, by
    the unmeasured style Z, and by each single concept (mnist_test/05_feature_analysis/analyze_mediation.py:128-173,
    batched over S pairs and written for any model with .decode(m, z)):

        total   = ||dec(m_b, z_b) - dec(m_a, z_a)||
        m_pct   = 100 * ||dec(m_b, z_a) - base|| / total          (all of M swapped, Z held)
        z_pct   = 100 * ||dec(m_a, z_b) - base|| / total          (Z swapped, M held)
        k_pct_k = 100 * ||dec(m_a with m_a[k] := m_b[k], z_a) - base|| / total

    All (K + 3) * S counterfactual rows go through ONE decode; the norms are reduced on device.
    Returns dict(total [S], m_pct [S], z_pct [S], feature_pct [S, K])."""
    S, K = m_a.shape
    eye = torch.eye(K, device=m_a.device, dtype=torch.bool)
    m_feat = torch.where(eye.unsqueeze(0), m_b.unsqueeze(1), m_a.unsqueeze(1))            # [S, K, K]: row k swaps concept k
    m_rows = torch.cat([m_b.unsqueeze(1), m_b.unsqueeze(1), m_a.unsqueeze(1), m_feat], dim=1)   # target, M-swap, Z-swap, K swaps
    z_rows = torch.cat([z_b.unsqueeze(1), z_a.unsqueeze(1), z_b.unsqueeze(1), z_a.unsqueeze(1).expand(S, K, -1)], dim=1)
    G = K + 3
    base = model.decode(m_a.contiguous(), z_a.contiguous())
    x = model.decode(m_rows.reshape(S * G, K).contiguous(), z_rows.reshape(S * G, -1).contiguous())
    d = rowdiff_l2(x, base, G).view(S, G)
    total = d[:, 0] + 1e-9
    return {"total": d[:, 0], "m_pct": 100.0 * d[:, 1] / total, "z_pct": 100.0 * d[:, 2] / total,
            "feature_pct": 100.0 * d[:, 3:] / total.unsqueeze(1)}


def ensemble_mean_std(preds, with_std=True):
    """torch.stack(preds).mean(0), .std(0) of up to 8 same-shaped CUDA tensors in one pass
    (ensemble_reconstruction.py:80-86)."""
    import ctypes as C
    preds = [p.contiguous() for p in preds]
    mean = torch.empty_like(preds[0])
    std = torch.empty_like(preds[0]) if with_std else None
    arr = (C.c_void_p * len(preds))(*[p.data_ptr() for p in preds])
    L.check(L.lib.cvae_ensemble_mean_std(C.cast(arr, C.c_void_p), len(preds), L.ptr(mean), L.ptr(std), mean.numel(),
                                         L.stream()), "ensemble_mean_std")
    return mean, std


@torch.no_grad()
def ensemble_reconstruction(models, x, m, t):
    """Mean and unbiased std of the eval-mode reconstructions of the fold models
    (vessel_analysis/04_generate_counterfactual/ensemble_reconstruction.py:58-89).  Returns (mean, std), [B,1,H,W]."""
    recons = []
    for model in models:
        model.eval()
        recons.append(model(x, m, t)[0])
    return ensemble_mean_std(recons)


@torch.no_grad()
def z_permutation_grid(models, x, m, t, scale=1.0):
    """The M x Z cross-product of vessel_analysis/03_evaluate_vessel/check_mechanism_z_perm.py:100-131: row i takes the
    measured concepts M of sample i, column j the style code z_j = mu(x_j, m_j, t_j) * scale of sample j; every cell is
    decoded by every fold model and the cells are averaged over the models.  All N*N rows of a model go through ONE
    decode.  Returns [N, N, 1, H, W] (cell (i, i) at scale 1 is the reconstruction of sample i)."""
    N, K = m.shape
    cells = []
    for model in models:
        model.eval()
        if hasattr(model, "encode"):
            mu = model.encode(x, m, t, torch.zeros(N, model.my_z_dim, device=x.device))[0]    # z = mu, no decode needed
        else:
            mu = model(x, m, t)[2]                                                             # the reference's own call
        Z = mu.shape[1]
        rows = ops.empty(N * N, K + Z, like=m)
        L.check(L.lib.cvae_pair_expand(L.ptr(m.contiguous()), L.ptr(mu.contiguous()), L.ptr(rows), N, K, Z, float(scale),
                                       L.stream()), "pair_expand")
        if hasattr(model, "dec_adapter"):                      # the ViT / CNN switch of check_mechanism_z_perm.py:119-124
            cells.append(model.backbone.decode(model.dec_adapter(rows)))
        else:
            cells.append(model.dec_conv(model.dec_fc(rows).view(-1, 512, *model.GRID)))
    mean = cells[0] if len(cells) == 1 else ensemble_mean_std(cells, with_std=False)[0]
    return mean.view(N, N, *mean.shape[1:])


class CounterfactualEngine:
    """High-throughput form of counterfactual_sweep for a frozen model: the whole sweep of one chunk of
    sources (base decode, do() scatter, decode of chunk*K rows, per-image L2) is captured once in a CUDA
    graph over static (m, z) buffers, and the weight layouts are packed once (ops.PackPlan; weights do not
    change in eval mode - call refresh() after loading new weights).  The eager sweep spends a third of
    its time in ~10^3 host-side enqueues per chunk; a replay is one launch.

    `lanes` > 1 keeps that many independent (buffers, graph, stream) sets and rotates chunks over them, so
    consecutive chunks overlap on the device (measured on B200 at chunk = 32: no gain - the image-sized
    layers fill the GPU on their own - hence the default of 1).  sweep_all() drives a whole source set.

        eng = CounterfactualEngine(model, chunk=32, delta=5.0)
        for i in range(0, S, 32):
            l2 = eng(m[i:i+32], z[i:i+32])        # [32, K]; valid until the lane is reused
        l2_all = eng.sweep_all(m, z)              # [S, K]
    """

    def __init__(self, model, chunk, delta=5.0, value=None, lanes=1):
        self.model, self.chunk, self.delta, self.value = model, int(chunk), delta, value
        model.eval()
        p0 = next(model.parameters())
        self.plan = ops.PackPlan()
        self._params = list(model.parameters())
        self._wver = self._weight_version()
        self.lanes = []
        self._next = 0
        for _ in range(max(1, int(lanes))):
            lane = {"m": torch.zeros(self.chunk, model.m_dim, device=p0.device),
                    "z": torch.zeros(self.chunk, model.my_z_dim, device=p0.device),
                    "stream": torch.cuda.Stream(), "done": torch.cuda.Event()}
            lane["stream"].wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(lane["stream"]), torch.no_grad():
                for _ in range(2):
                    self._sweep(lane)
            torch.cuda.synchronize()
            lane["graph"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(lane["graph"], stream=lane["stream"]), torch.no_grad():
                lane["l2"] = self._sweep(lane)
            self.lanes.append(lane)
        torch.cuda.current_stream().wait_stream(self.lanes[-1]["stream"])

    def _sweep(self, lane):
        ops.set_pack_plan(self.plan)
        try:
            l2, _, _ = counterfactual_sweep(self.model, lane["m"], lane["z"], delta=self.delta, value=self.value)
        finally:
            ops.set_pack_plan(None)
        if self.plan.recording:
            self.plan.finalize()
        return l2

    def _weight_version(self):
        """changes whenever a parameter is written in place (optimizer step, load_state_dict, copy_)"""
        return sum(p._version for p in self._params) + sum(p.data_ptr() for p in self._params[:1])

    def refresh(self):
        """re-pack every weight layout from the model's current parameters (one launch)"""
        self.plan.run()
        self._wver = self._weight_version()

    def _launch(self, m, z):
        if m.shape[0] != self.chunk:
            raise RuntimeError(f"CounterfactualEngine was captured for chunks of {self.chunk} sources, got {m.shape[0]}")
        lane = self.lanes[self._next]
        self._next = (self._next + 1) % len(self.lanes)
        st = lane["stream"]
        st.wait_stream(torch.cuda.current_stream())       # inputs are ready; the lane's previous result was consumed
        if self._weight_version() != self._wver:          # weights moved since the layouts were packed: repack first
            for ln in self.lanes:                         # (every lane's graph reads the same packed buffers)
                torch.cuda.current_stream().wait_stream(ln["stream"])
            self.refresh()
            st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            lane["m"].copy_(m, non_blocking=True)
            lane["z"].copy_(z, non_blocking=True)
            lane["graph"].replay()
            lane["done"].record(st)
        return lane

    @torch.no_grad()
    def __call__(self, m, z):
        lane = self._launch(m, z)
        torch.cuda.current_stream().wait_event(lane["done"])
        return lane["l2"]

    @torch.no_grad()
    def sweep_all(self, m, z):
        """[S, K] effect sizes of all sources; chunks rotate over the lanes so that they overlap on the device."""
        S, K = m.shape
        if S % self.chunk:
            raise RuntimeError(f"number of sources ({S}) must be a multiple of the chunk size ({self.chunk})")
        out = torch.empty(S, K, device=m.device)
        cur = torch.cuda.current_stream()
        pending = []
        for i in range(0, S, self.chunk):
            if len(pending) == len(self.lanes):            # the lane about to be reused: drain its result first
                j, ln = pending.pop(0)
                cur.wait_event(ln["done"])
                out[j:j + self.chunk].copy_(ln["l2"], non_blocking=True)
            pending.append((i, self._launch(m[i:i + self.chunk], z[i:i + self.chunk])))
        for j, ln in pending:
            cur.wait_event(ln["done"])
            out[j:j + self.chunk].copy_(ln["l2"], non_blocking=True)
        return out
