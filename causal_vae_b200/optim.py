"""Flat parameter / gradient storage and the fused global-norm clip + Adam step
(torch.nn.utils.clip_grad_norm_(max_norm) followed by torch.optim.Adam defaults:
vessel_analysis/01_train/train.py:85-86,152)."""
import torch

from . import _lib as L
from . import ops


class FlatParams:
    """Re-homes every parameter of `model` into one contiguous fp32 buffer (each tensor 16-byte aligned)
    with a matching gradient buffer; `p.data` / `p.grad` become views, so stock PyTorch code
    (state_dict, torch.optim, autograd accumulation) keeps working while the optimizer, the gradient
    all-reduce and zero_grad become single passes over flat memory."""

    def __init__(self, model):
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self.data = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                k = p.numel()
                self.data[o:o + k].copy_(p.data.reshape(-1))
                p.data = self.data[o:o + k].view(p.shape)
                p.grad = self.grad[o:o + k].view(p.shape)
        self.offsets = offs

    def zero_grad(self):
        ops.fill(self.grad, 0.0)
        for p, o in zip(self.params, self.offsets):   # autograd may have replaced .grad; re-attach the views
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view(p.shape)


class FusedClipAdam:
    """clip_grad_norm_(max_norm) + Adam(lr, betas, eps) in two kernels over FlatParams.
    max_norm=None disables clipping (causal_cascade / latent_translator / mnist loops)."""

    def __init__(self, flat, lr, max_norm=None, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        self.flat, self.lr, self.max_norm, self.betas, self.eps = flat, lr, max_norm, betas, eps
        self.grad_scale = grad_scale
        dev = flat.data.device
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        self.step_count = torch.zeros((), dtype=torch.int64, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)

    def zero_grad(self, set_to_none=False):
        self.flat.zero_grad()

    def grad_norm(self):
        """total L2 norm of the last step's gradient (device scalar, fp64)."""
        return self.sumsq.sqrt()

    def step(self):
        f = self.flat
        s = L.stream()
        clip = self.max_norm is not None
        if clip:
            L.check(L.lib.cvae_fill(L.ptr(self.sumsq.view(torch.float32)), 2, 0.0, s), "fill")
            L.check(L.lib.cvae_sumsq(L.ptr(f.grad), f.numel, L.ptr(self.sumsq), s), "sumsq")
        L.check(L.lib.cvae_clip_adam(L.ptr(f.data), L.ptr(f.grad), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq),
                                     f.numel, L.ptr(self.sumsq) if clip else None,
                                     float(self.max_norm) if clip else 0.0, self.lr, self.betas[0], self.betas[1],
                                     self.eps, self.grad_scale, L.ptr(self.step_count), s), "clip_adam")
