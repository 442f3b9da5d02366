"""causal_cascade/train.py:5-40 on the native kernels: loss_function and the inner training step."""
from .. import functional as F
from ..chain import direct_grads
from ..optim import FlatParams, FusedClipAdam


def loss_function(recon_x, x, m_hat, m, mu, logvar, gamma=2000.0):
    """(loss, recon_loss, m_loss) — MSE_sum + gamma * MSE_sum(m_hat, m) + KL (train.py:5-17)."""
    recon_loss = F.mse_sum(recon_x, x)
    m_loss = F.mse_sum(m_hat, m)
    kld = F.kld_loss(mu, logvar)
    return recon_loss + gamma * m_loss + kld, recon_loss, m_loss


class CascadeTrainer:
    """zero_grad, forward, loss, backward, Adam(lr=1e-3) step (train.py:28-34; main.py:50)."""

    def __init__(self, model, lr=1e-3, gamma=2000.0):
        self.model, self.gamma = model, gamma
        self.opt = FusedClipAdam(FlatParams(model), lr)

    def step(self, x, m, t, eps=None):
        self.model.train()
        self.opt.zero_grad()
        recon_x, m_hat, mu, logvar = self.model(x, m, t, eps)
        loss, l_recon, l_m = loss_function(recon_x, x, m_hat, m, mu, logvar, self.gamma)
        with direct_grads():
            loss.backward()
        self.opt.step()
        return loss, l_recon, l_m
