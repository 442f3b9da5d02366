"""Whole-step CUDA-graph capture shared by the small-model trainers (mnist / cascade / latent_translator).

These steps are launch-bound (whole-network activations of 0.17 / 1.0 / 11 MB per sample), so one graph
replay per step replaces a few hundred enqueues.  VesselTrainer keeps its own capture (side streams, early
all-reduce); this helper covers the plain single-stream steps."""
import gc

import torch


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
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn()

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
