"""Vessel training step — mirrors vessel_analysis/01_train/train.py:18-98 (loss_function,
train_one_epoch's inner step) on the native kernels, plus a CUDA-graph-captured whole step."""
import os

import torch

from .. import functional as F
from .. import ops
from ..chain import direct_grads, side_wgrad
from ..optim import FlatParams, FusedClipAdam
from ..parallel import allreduce_gradients
from .models import CONFIG


def loss_function(recon_x, x, m_hat, m, mu, logvar, m_mu, m_logvar):
    """(recon_loss, kld_loss, morph_loss, sparsity_loss) — train.py:18-60.  One fused pass computes
    the weighted-MSE and sparsity sums (pos_weight from x.sum() on device), one each the KL and the
    Gaussian NLL; all accumulate in fp64."""
    recon_loss, sparsity_loss = F.vessel_recon_loss(recon_x, x)
    kld_loss = F.kld_loss(mu, logvar)
    morph_loss = F.gauss_nll_loss(m, m_mu, m_logvar)
    return recon_loss, kld_loss, morph_loss, sparsity_loss


def total_loss(recon, kld, morph, sparsity, beta=None, lambda_morph=1.0):
    """train.py:82 (lambda_morph = 1) / train_kfold.py:71 (lambda_morph = CONFIG['LAMBDA_MORPH'])."""
    beta = CONFIG["BETA"] if beta is None else beta
    if all(torch.is_tensor(v) and v.is_cuda and v.dim() == 0 for v in (recon, kld, morph, sparsity)):
        return F.weighted_sum((recon, kld, morph, sparsity), (1.0, beta, lambda_morph, 0.3))   # one kernel each way
    return recon + beta * kld + lambda_morph * morph + 0.3 * sparsity


class VesselTrainer:
    """fwd + loss + bwd + clip_grad_norm_(5.0) + Adam(lr) for CausalViTVAE (train.py:77-86).

    `step(x, m, t)` runs eagerly; `capture(B)` records the whole step into a CUDA graph over static
    input buffers so a replay costs one launch (the eager step is ~10^3 small enqueues).  With
    `world_size > 1` the flat gradient is all-reduced (SUM: every reference loss is sum-reduced, so
    the sum of shard gradients is the single-device gradient — SURVEY §8(e)) before the clip."""

    def __init__(self, model, lr=None, max_norm=5.0, beta=None, lambda_morph=1.0, process_group=None,
                 distributed=False, overlap_wgrad=True):
        self.model = model
        self.flat = FlatParams(model)
        self.opt = FusedClipAdam(self.flat, CONFIG["LEARNING_RATE"] if lr is None else lr, max_norm)
        self.beta, self.lambda_morph = beta, lambda_morph
        self.distributed, self.pg = distributed, process_group
        self.graph = None
        self.static = None
        self.pack_plan = ops.PackPlan()
        self.side = torch.cuda.Stream() if overlap_wgrad else None
        self._early_done, self._seg, self.comm = False, None, None
        self._mid, self._mid_done = None, False
        if distributed and os.environ.get("CVAE_DP_OVERLAP", "1") != "0":
            self._seg = self._decoder_segment()
            if self._seg is not None:
                self.comm = torch.cuda.Stream()
                model._decoder_grad_hook = self._early_allreduce
                self._mid = self._transformer_start()
                if self._mid is not None and hasattr(model, "backbone"):
                    model.backbone._tokens_grad_hook = self._mid_allreduce
        rank = 0
        if distributed:
            import torch.distributed as dist
            rank = dist.get_rank(process_group)
        # this trainer's own dropout generator: seeded from torch.manual_seed / F.manual_seed and the rank (shards of a
        # data-parallel batch draw independent masks), keyed on this trainer's step counter so graph replays differ
        self.rng = F.RngState(counter=self.opt.step_count, rank=rank)

    def _fwd_bwd(self, x, m, t, eps):
        ops.arena_begin(self.flat.data.device)
        ops.set_pack_plan(self.pack_plan)
        packed = not self.pack_plan.recording
        side_prep = self.side is not None and os.environ.get("CVAE_SIDE_PREP", "1") != "0"
        if side_prep:
            # off the critical path: the gradient buffer is zeroed and the weight layouts first needed in backward
            # are packed on the side stream while the forward pass runs (joined when backward starts)
            self.side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.side):
                self.opt.zero_grad()
                if packed:
                    self.pack_plan.run("bwd")
            if packed:
                self.pack_plan.run("fwd")
        else:
            self.opt.zero_grad()
            if packed:
                self.pack_plan.run()      # every weight layout of the step, one launch
        try:
            with F.use_rng(self.rng):
                out = self.model(x, m, t, eps)
            self.last_outputs = tuple(o.detach() for o in out)      # the step's forward tuple (views, no graph)
            recon, kld, morph, sp = loss_function(out[0], x, out[1], m, out[2], out[3], out[4], out[5])
            loss = total_loss(recon, kld, morph, sp, self.beta, self.lambda_morph)
            # zero_grad() above zeroed the flat buffer and every parameter is used once: gradients are written
            # in place, the weight-gradient kernels on a side stream (joined on exit, before the optimizer)
            with direct_grads(), side_wgrad(self.side, join_first=side_prep):
                loss.backward()
        finally:
            ops.arena_end()
            ops.set_pack_plan(None)
        if self.pack_plan.recording:
            self.pack_plan.finalize()
        return loss, recon, kld, morph, sp

    # ---- data-parallel gradient exchange ---------------------------------------------------------------------
    # The decoder (decoder_input + conv stack: 9.2 M of the 14.1 M parameters, contiguous in the flat buffer) is the
    # FIRST part of the model whose gradients are complete.  Its all-reduce starts on a communication stream as soon
    # as the decoder's backward has been enqueued and overlaps the rest of backward (adapters, ViT, stem); the
    # remaining 19 MB are reduced after the join.
    def _decoder_segment(self):
        names = {id(p): n for n, p in self.model.named_parameters()}
        idx = [i for i, p in enumerate(self.flat.params)
               if names[id(p)].startswith(("backbone.decoder_input.", "backbone.decoder."))]
        if not idx or idx != list(range(idx[0], idx[-1] + 1)):
            return None
        lo = self.flat.offsets[idx[0]]
        hi = self.flat.offsets[idx[-1] + 1] if idx[-1] + 1 < len(self.flat.params) else self.flat.numel
        return lo, hi

    def _transformer_start(self):
        """offset of the first transformer parameter when [transformer .. fc_var] directly precedes the decoder segment"""
        names = {id(p): n for n, p in self.model.named_parameters()}
        lo = self._seg[0]
        start = None
        for i, p in enumerate(self.flat.params):
            n = names[id(p)]
            if self.flat.offsets[i] >= lo:
                break
            inside = n.startswith(("backbone.transformer.", "backbone.to_latent.", "backbone.fc_mu.", "backbone.fc_var."))
            if inside and start is None:
                start = self.flat.offsets[i]
            if not inside and start is not None:
                return None                      # something else sits between the transformer and the decoder
        return start

    def _mid_allreduce(self, _grad):
        """Second bucket: fires when the transformer's backward has been enqueued (gradient of the token sequence), i.e.
        before the stem's.  Transformer / to_latent gradients and everything behind the decoder segment (adapters,
        morphology head: complete long before) are reduced while the stem backward runs; only the stem + embeddings
        (4 MB) remain for after the join."""
        import torch.distributed as dist
        if not self._early_done:
            return None
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)
        if self.side is not None:
            self.comm.wait_stream(self.side)
        lo, hi = self._seg
        with torch.cuda.stream(self.comm):
            dist.all_reduce(self.flat.grad[self._mid:lo], op=dist.ReduceOp.SUM, group=self.pg)
            if hi < self.flat.numel:
                dist.all_reduce(self.flat.grad[hi:], op=dist.ReduceOp.SUM, group=self.pg)
        self._mid_done = True
        return None

    def _early_allreduce(self, _grad):
        import torch.distributed as dist
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)                       # decoder input gradients (main stream) ...
        if self.side is not None:
            self.comm.wait_stream(self.side)             # ... and weight gradients (side stream) are complete
        lo, hi = self._seg
        with torch.cuda.stream(self.comm):
            dist.all_reduce(self.flat.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
        self._early_done = True
        return None

    def _allreduce(self):
        if not self.distributed:
            return
        if self._early_done:
            import torch.distributed as dist
            lo, hi = self._seg
            if self._mid_done:
                lo, hi = self._mid, self.flat.numel          # only the stem + embeddings are left
            if lo > 0:
                dist.all_reduce(self.flat.grad[:lo], op=dist.ReduceOp.SUM, group=self.pg)
            if hi < self.flat.numel:
                dist.all_reduce(self.flat.grad[hi:], op=dist.ReduceOp.SUM, group=self.pg)
            torch.cuda.current_stream().wait_stream(self.comm)
            self._early_done = self._mid_done = False
        else:
            allreduce_gradients(self.flat.grad, group=self.pg)

    def step(self, x, m, t, eps=None):
        self.model.train()
        losses = self._fwd_bwd(x, m, t, eps)
        self._allreduce()
        self.opt.step()
        return losses

    # ---- CUDA graph ------------------------------------------------------------------------------
    def capture(self, B, H=None, W=None, warmup=3):
        H = CONFIG["IMG_HEIGHT"] if H is None else H
        W = CONFIG["IMG_WIDTH"] if W is None else W
        dev = self.flat.data.device
        self.static = dict(
            x=torch.zeros(B, 1, H, W, device=dev), m=torch.zeros(B, CONFIG["M_DIM"], device=dev),
            t=torch.zeros(B, CONFIG["T_DIM"], device=dev), eps=torch.zeros(B, CONFIG["Z_DIM"], device=dev))
        self.static["t"][:, 0] = 1.0
        self.model.train()
        snap = (self.flat.data.clone(), {k: v.clone() for k, v in self.model.state_dict().items()})
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._fwd_bwd(**self.static)
                self._allreduce()
                self.opt.step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        # undo the warm-up updates (parameters, BN running stats, Adam state)
        with torch.no_grad():
            self.flat.data.copy_(snap[0])
            for k, v in self.model.state_dict().items():
                if not v.is_floating_point() or k.endswith(("running_mean", "running_var")):
                    v.copy_(snap[1][k])
            self.opt.exp_avg.zero_(); self.opt.exp_avg_sq.zero_(); self.opt.step_count.zero_()
        # Autograd graphs of the warm-up steps can survive in reference cycles; their AccumulateGrad nodes
        # remember the warm-up stream, and a leaf node that receives no gradient (direct_grads) would make
        # the capturing stream wait on that uncaptured stream.  Collect them so the capture pass builds
        # fresh nodes on the capture stream.
        import gc
        gc.collect()
        # CVAE_GRAPH_PRIO=1 (experiment, off by default): per-node priorities (csrc/graph_prio.cu) so that the main chain's
        # grids overtake queued weight-gradient CTAs of the side branch.  Measured: 7.885 ms against 7.797 ms for the plain
        # replay on the same box - the side kernels are single-wave grids that stay resident, so there is nothing queued
        # to overtake, and the GPU is busy either way (the step is bound by the sum of the two branches' work).
        prio = self.side is not None and os.environ.get("CVAE_GRAPH_PRIO", "0") == "1"
        self.graph = torch.cuda.CUDAGraph(keep_graph=True) if prio else torch.cuda.CUDAGraph()
        self._prio_exec = None
        from .. import _lib as _L
        n0 = _L.launch_count
        with torch.cuda.graph(self.graph):
            self.static_losses = self._fwd_bwd(**self.static)
            self._allreduce()
            self.opt.step()
        self.captured_launches = _L.launch_count - n0      # C-ABI calls recorded in the graph = kernels of ours per replay
        # the capture pass itself does not execute; nothing to undo
        if prio:
            from .. import _lib as L
            self._prio_exec = L.PriorityGraphExec(self.graph, prio_main=int(os.environ.get("CVAE_PRIO_MAIN", "-1")),
                                                  prio_side=int(os.environ.get("CVAE_PRIO_SIDE", "0")))
        return self

    def load_batch(self, x, m, t, eps=None):
        """host (pinned) or device tensors -> static graph inputs (async H2D on the current stream)."""
        self.static["x"].copy_(x, non_blocking=True)
        self.static["m"].copy_(m, non_blocking=True)
        self.static["t"].copy_(t, non_blocking=True)
        if eps is not None:
            self.static["eps"].copy_(eps, non_blocking=True)
        else:
            self.static["eps"].normal_()

    # ---- double-buffered input: the H2D copy of batch i+1 overlaps the replay of batch i ------------------
    def prefetch(self, x, m, t, eps=None):
        """Start the host -> device copy of the NEXT batch on a copy stream into staging buffers
        (pinned host tensors make it asynchronous).  commit_prefetched() hands it to the graph."""
        if getattr(self, "_stage", None) is None:
            self._stage = {k: torch.empty_like(v) for k, v in self.static.items()}
            self._copy_stream = torch.cuda.Stream()
            self._copy_done = torch.cuda.Event()
            self._stage_free = torch.cuda.Event()
            self._stage_free.record()
        self._copy_stream.wait_event(self._stage_free)          # the previous commit has read the staging buffers
        with torch.cuda.stream(self._copy_stream):
            self._stage["x"].copy_(x, non_blocking=True)
            self._stage["m"].copy_(m, non_blocking=True)
            self._stage["t"].copy_(t, non_blocking=True)
            if eps is not None:
                self._stage["eps"].copy_(eps, non_blocking=True)
            else:
                self._stage["eps"].normal_()
            self._copy_done.record()

    def commit_prefetched(self):
        """staging -> static graph inputs (device-to-device, ordered after the prefetch and before the replay)"""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._copy_done)
        for k, v in self.static.items():
            v.copy_(self._stage[k], non_blocking=True)
        self._stage_free.record(cur)

    def replay(self):
        if getattr(self, "_prio_exec", None) is not None:
            self._prio_exec.launch()
        else:
            self.graph.replay()
        return self.static_losses


# ---- epoch loops with the reference's names (vessel_analysis/01_train/train.py:62-133) ------------------------------
_TRAINERS = {}          # id(torch optimizer) -> VesselTrainer built for it (reference-signature calls)


def _trainer_for(vae, opt):
    """The fused trainer behind a stock `torch.optim.Adam(vae.parameters(), lr=...)` (train.py:152): same lr / betas /
    eps, clip_grad_norm_(5.0) as train.py:85.  Built once per optimizer object."""
    if isinstance(opt, VesselTrainer):
        return opt
    tr = _TRAINERS.get(id(opt))
    if tr is None or tr.model is not vae:
        if not isinstance(opt, torch.optim.Adam):
            raise RuntimeError("train_one_epoch takes a VesselTrainer or the torch.optim.Adam the reference builds "
                               f"(train.py:152); got {type(opt).__name__}")
        g = opt.param_groups[0]
        if len(opt.param_groups) != 1 or g.get("weight_decay", 0) != 0 or g.get("amsgrad", False):
            raise RuntimeError("only the reference's optimizer form is fused: one group, no weight decay, no amsgrad")
        tr = VesselTrainer(vae, lr=g["lr"], max_norm=5.0)
        tr.opt.betas, tr.opt.eps = tuple(g["betas"]), g["eps"]
        _TRAINERS[id(opt)] = tr
    return tr


def train_one_epoch(*args, **kw):
    """train.py:62-98.  Accepts the reference's call `train_one_epoch(epoch, vae, train_loader, opt_vae)` (opt_vae: the
    stock Adam of train.py:152 or a VesselTrainer) and the keyword form `train_one_epoch(vae, train_loader, trainer,
    epoch=0)`.  Returns the epoch loss per sample.  Unlike the reference the running totals stay on the device: one
    host synchronisation per epoch, not four per batch."""
    if args and isinstance(args[0], int):
        epoch, vae, train_loader, optimizer = args[:4]
    else:
        vae, train_loader, optimizer = args[:3]
    trainer = _trainer_for(vae, optimizer)
    dev = kw.get("device") or trainer.flat.data.device
    total = torch.zeros((), device=dev)
    parts = torch.zeros(2, device=dev)
    n = 0
    for x, m, t in train_loader:
        losses = trainer.step(x.to(dev, non_blocking=True), m.to(dev, non_blocking=True), t.to(dev, non_blocking=True))
        total += losses[0].detach()
        parts += torch.stack([losses[1].detach(), losses[2].detach()])
        n += x.shape[0]
    size = len(train_loader.dataset) if hasattr(train_loader, "dataset") else max(n, 1)
    recon, kld = (parts / size).tolist()
    print(f"   [Train Breakdown] Recon: {recon:.1f} | KLD: {kld:.1f}")
    return float(total) / size


@torch.no_grad()
def validate(vae, val_loader, device=None):
    """train.py:100-133: eval-mode forward + the four loss terms over a loader; returns the validation loss per
    sample (recon + BETA * kld + morph + 0.3 * sparsity).  validate.breakdown holds the per-sample recon / kld /
    morph averages the reference prints."""
    vae.eval()
    dev = next(vae.parameters()).device if device is None else device
    tot = torch.zeros(4, device=dev)
    n = 0
    for x, m, t in val_loader:
        x, m, t = (a.to(dev, non_blocking=True) for a in (x, m, t))
        out = vae(x, m, t)
        recon, kld, morph, sp = loss_function(out[0], x, out[1], m, out[2], out[3], out[4], out[5])
        tot += torch.stack([total_loss(recon, kld, morph, sp), recon, kld, morph])
        n += x.shape[0]
    size = len(val_loader.dataset) if hasattr(val_loader, "dataset") else max(n, 1)
    vals = (tot / size).tolist()
    validate.breakdown = {"recon": vals[1], "kld": vals[2], "morph": vals[3]}
    print(f"   [Val Breakdown] Recon: {vals[1]:.1f} | KLD: {vals[2]:.1f} | Morph: {vals[3]:.1f}")
    return vals[0]
