"""latent_translator/engine.py:6-52 on the native kernels: the ViTVAE training step (MEAN-reduced
MSE + beta * MEAN-reduced KL) and latent extraction (mu of an eval-mode encode)."""
import numpy as np
import torch

from .. import functional as F
from ..graph import GraphedStep, StepScope, trainer_state
from ..optim import FlatParams, FusedClipAdam
from ..parallel import allreduce_gradients


def loss_function(recons, x, mu, log_var, beta=1.0):
    """(loss, recon_loss, kld_loss) — engine.py:25-27."""
    recon_loss = F.mse_mean(recons, x)
    kld_loss = F.kld_mean(mu, log_var)
    return recon_loss + beta * kld_loss, recon_loss, kld_loss


class ViTVAETrainer:
    """zero_grad, forward, loss, backward, Adam(lr=1e-4) step (engine.py:19-30; main.py:29).
    Data-parallel use averages the gradients (mean-reduced loss): grad_scale = 1/world."""

    def __init__(self, model, lr=1e-4, beta=1.0, grad_scale=1.0, distributed=False, process_group=None):
        self.model, self.beta = model, beta
        self.distributed, self.pg = distributed, process_group
        rank = 0
        if distributed:
            import torch.distributed as dist
            rank = dist.get_rank(process_group)
            grad_scale = grad_scale / dist.get_world_size(process_group)      # mean-reduced loss: average
        self.opt = FusedClipAdam(FlatParams(model), lr, grad_scale=grad_scale)
        self.rng = F.RngState(counter=self.opt.step_count, rank=rank)         # this trainer's dropout generator
        self.graphed = None
        self.scope = StepScope(self.opt.flat.data.device)

    def step(self, x, eps=None):
        self.model.train()
        self.opt.zero_grad()
        with self.scope:
            with F.use_rng(self.rng):
                recons, _, mu, log_var = self.model(x, eps)
            loss, rl, kl = loss_function(recons, x, mu, log_var, self.beta)
            with self.scope.backward():
                loss.backward()
        if self.distributed:
            allreduce_gradients(self.opt.flat.grad, group=self.pg)
        self.opt.step()
        return loss, rl, kl

    def capture(self, B, H, W, latent=512, warmup=3):
        """The whole step (engine.py:19-30) as one CUDA graph over static x and eps."""
        dev = self.opt.flat.data.device
        st = dict(x=torch.zeros(B, 1, H, W, device=dev), eps=torch.zeros(B, latent, device=dev))
        self.graphed = GraphedStep(lambda: self.step(st["x"], st["eps"]), st,
                                   trainer_state([self.model], [self.opt]), warmup)
        return self.graphed


def train_vit_vae(model, loader, optimizer, device, epochs, beta=1.0):
    """Same loop as engine.py:6-36; `optimizer` may be a ViTVAETrainer (fused) or a torch optimizer."""
    trainer = optimizer if isinstance(optimizer, ViTVAETrainer) else None
    model.train()
    for ep in range(1, epochs + 1):
        total_loss, n_samples = 0.0, 0
        for batch in loader:
            x = batch["x"].to(device)
            if trainer is not None:
                loss = trainer.step(x)[0]
            else:
                optimizer.zero_grad()
                recons, _, mu, log_var = model(x)
                loss = loss_function(recons, x, mu, log_var, beta)[0]
                loss.backward()
                optimizer.step()
            total_loss += loss.item() * x.size(0)
            n_samples += x.size(0)
        print(f"[ViTVAE] Epoch {ep:03d}/{epochs} | Loss: {total_loss / max(n_samples, 1):.6f}")


@torch.no_grad()
def extract_vit_latents(model, loader, device):
    """engine.py:38-52: mu of model.encode(x) over a loader -> numpy [N, latent]."""
    model.eval()
    zs = []
    for batch in loader:
        mu, _ = model.encode(batch["x"].to(device))
        zs.append(mu.detach().cpu().numpy())
    return np.concatenate(zs, axis=0)
