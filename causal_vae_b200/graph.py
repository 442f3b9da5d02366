"""Whole-step CUDA-graph capture shared by the small-model trainers (mnist / cascade / latent_translator).

These steps are launch-bound (whole-network activations of 0.17 / 1.0 / 11 MB per sample), so one graph
replay per step replaces a few hundred enqueues.  VesselTrainer keeps its own capture (side streams, early
all-reduce); this helper covers the plain single-stream steps."""
import gc
import os

import torch

from . import ops
from .chain import direct_grads, side_wgrad


class GraphedStep:
    """Capture `fn()` — a full training step reading `static` input buffers — into one CUDA graph.

    `state` lists every tensor the step mutates (flat parameters, Adam moments, step counters, BatchNorm running
    statistics): the eager warm-up steps that allocator and autograd need before capture are undone by restoring
    them, so the first replay is step 1 of the run."""

    def __init__(self, fn, static, state, warmup=3):
        self.static = static
        snap = [t.clone() for t in state]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.no_grad():
            for t, v in zip(state, snap):
                t.copy_(v)
        gc.collect()                       # stale autograd graphs remember the warm-up stream (see VesselTrainer.capture)
        from . import _lib as L
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count
        with torch.cuda.graph(self.graph):
            self.outputs = fn()
        self.captured_launches = L.launch_count - n0       # C-ABI calls recorded in the graph (bench.py: gpu_launches)

    def load(self, **tensors):
        for k, v in tensors.items():
            if v is not None:
                self.static[k].copy_(v, non_blocking=True)

    def replay(self):
        self.graph.replay()
        return self.outputs


def trainer_state(models, opts):
    """the tensors a step mutates: flat parameters + Adam state of each optimizer, module buffers of each model"""
    out = []
    for o in opts:
        out += [o.flat.data, o.exp_avg, o.exp_avg_sq, o.step_count]
    for m in models:
        out += [b for b in m.buffers()]
    return out


class StepScope:
    """What VesselTrainer does around its forward + backward, for the single-optimizer small trainers:

    * the step's small fp64 accumulators come from one arena zeroed by one launch (ops.arena_begin);
    * every weight re-layout of the step is recorded on the first step and replayed as ONE launch at the start of each
      later step (ops.PackPlan; the latent_translator step spent 80 launches / 0.34 ms on per-use packs);
    * in backward the weight-gradient kernels run on a side stream, forked per layer and joined before the optimizer
      (chain.side_wgrad): a layer's weight and input gradients are independent, and most grids of these models leave
      SMs idle.  The fork / join pattern is captured when the step is captured in a CUDA graph.

        with scope:                      # forward + loss + backward
            out = model(...); loss = ...
            with scope.backward():
                loss.backward()

    CVAE_SMALL_SCOPE=0 turns the plan and the side stream off (A/B runs)."""

    def __init__(self, device):
        on = os.environ.get("CVAE_SMALL_SCOPE", "1") != "0"
        self.device = device
        self.plan = ops.PackPlan() if on else None
        self.side = torch.cuda.Stream(device) if on and os.environ.get("CVAE_SMALL_SIDE", "1") != "0" else None

    def __enter__(self):
        ops.arena_begin(self.device)
        if self.plan is not None:
            ops.set_pack_plan(self.plan)
            if not self.plan.recording:
                self.plan.run()
        return self

    def __exit__(self, *exc):
        ops.arena_end()
        ops.set_pack_plan(None)
        if self.plan is not None and self.plan.recording and exc[0] is None:
            self.plan.finalize()
        return False

    def backward(self):
        import contextlib
        st = contextlib.ExitStack()
        st.enter_context(direct_grads())
        st.enter_context(side_wgrad(self.side))
        return st
