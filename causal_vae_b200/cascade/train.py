"""causal_cascade/train.py:5-40 on the native kernels: loss_function and the inner training step."""
from .. import functional as F
from ..graph import GraphedStep, StepScope, trainer_state
from ..optim import FlatParams, FusedClipAdam
from ..parallel import allreduce_gradients


def loss_function(recon_x, x, m_hat, m, mu, logvar, gamma=2000.0):
    """(loss, recon_loss, m_loss) — MSE_sum + gamma * MSE_sum(m_hat, m) + KL (train.py:5-17)."""
    recon_loss = F.mse_sum(recon_x, x)
    m_loss = F.mse_sum(m_hat, m)
    kld = F.kld_loss(mu, logvar)
    return recon_loss + gamma * m_loss + kld, recon_loss, m_loss


class CascadeTrainer:
    """zero_grad, forward, loss, backward, Adam(lr=1e-3) step (train.py:28-34; main.py:50)."""

    def __init__(self, model, lr=1e-3, gamma=2000.0, distributed=False, process_group=None):
        self.model, self.gamma = model, gamma
        self.opt = FusedClipAdam(FlatParams(model), lr)
        self.distributed, self.pg = distributed, process_group      # SUM of shard gradients (sum-reduced loss)
        self.graphed = None
        self.scope = StepScope(self.opt.flat.data.device)

    def step(self, x, m, t, eps=None):
        self.model.train()
        self.opt.zero_grad()
        with self.scope:
            recon_x, m_hat, mu, logvar = self.model(x, m, t, eps)
            loss, l_recon, l_m = loss_function(recon_x, x, m_hat, m, mu, logvar, self.gamma)
            with self.scope.backward():
                loss.backward()
        if self.distributed:
            allreduce_gradients(self.opt.flat.grad, group=self.pg)
        self.opt.step()
        return loss, l_recon, l_m

    def capture(self, B, H=64, W=64, m_dim=8, z_dim=64, warmup=3):
        """The whole step (train.py:28-34) as one CUDA graph over static inputs x, m, t (int64), eps."""
        import torch
        dev = self.opt.flat.data.device
        st = dict(x=torch.zeros(B, 1, H, W, device=dev), m=torch.zeros(B, m_dim, device=dev),
                  t=torch.zeros(B, dtype=torch.int64, device=dev), eps=torch.zeros(B, z_dim, device=dev))
        self.graphed = GraphedStep(lambda: self.step(st["x"], st["m"], st["t"], st["eps"]), st,
                                   trainer_state([self.model], [self.opt]), warmup)
        return self.graphed
