"""The adversarial training step of mnist_test/01_baseline_causal_vae/train.py:36-89 (06 variant:
mnist_test/06_model_experiment/train.py:41-96) on the native kernels: discriminator step
(cross-entropy on argmax(t)) then VAE step (BCE_sum + beta*KL + morph + confusion loss)."""
import torch

from .. import functional as F
from ..graph import GraphedStep, StepScope, trainer_state
from ..optim import FlatParams, FusedClipAdam
from ..parallel import allreduce_gradients
from .models import CONFIG


def vae_loss(vae, disc, x, m, t, eps=None, eps_adv=None, beta=None, lambda_adv=None):
    """(loss, recon, kld, morph, adv) of train.py:65-87; 6-tuple models use the Gaussian NLL morph term."""
    beta = CONFIG["BETA"] if beta is None else beta
    lambda_adv = CONFIG["LAMBDA_ADV"] if lambda_adv is None else lambda_adv
    out = vae(x, m, t, eps)
    recon_x, m_hat, mu, logvar = out[:4]
    l_rec = F.bce_sum(recon_x.reshape(-1, 784), x.reshape(-1, 784))
    l_kld = F.kld_loss(mu, logvar, beta)
    if len(out) == 4:
        l_m = F.mse_sum(m_hat, m, 100.0)
    else:
        l_m = F.gauss_nll_loss(m, out[4], out[5])
    logits = disc(vae.reparameterize(mu, logvar, eps_adv))
    l_adv = F.uniform_kl_batchmean(logits, lambda_adv * 100.0)
    return l_rec + l_kld + l_m + l_adv, l_rec, l_kld, l_m, l_adv


def disc_loss(vae, disc, x, m, t, eps=None):
    """train.py:41-55: z from a no-grad VAE pass, CE(D(z), argmax t)."""
    with torch.no_grad():
        _, _, z = vae.encode(x, m, t, eps)
    return F.cross_entropy(disc(z), F.argmax_rows(t))


class AdversarialTrainer:
    """opt_d / opt_vae = Adam(lr=CONFIG['LR']) over flat parameter buffers (train.py:21-22)."""

    def __init__(self, vae, disc, lr=None, distributed=False, process_group=None, fused=True):
        lr = CONFIG["LR"] if lr is None else lr
        self.vae, self.disc = vae, disc
        self.distributed, self.pg = distributed, process_group
        self.graphed = None
        self.opt_vae = FusedClipAdam(FlatParams(vae), lr)
        self.opt_d = FusedClipAdam(FlatParams(disc), lr)
        # Each half of the step has its own scope (graph.StepScope: fp64 arena, one-launch weight packing, parameter
        # gradients written in place, weight-gradient kernels on a side stream): the discriminator's weights change
        # between the halves, so the layouts of the second half are packed after opt_d.step().  In both halves every
        # parameter is used once; the second half also leaves gradients in the discriminator's buffer, as the reference's
        # loss_vae.backward() does -- opt_d.zero_grad() clears them at the start of the next step.  fused=False keeps the
        # plain autograd accumulation (the checker of tests/test_families_gpu.py).
        dev = self.opt_vae.flat.data.device
        self.scope_d = StepScope(dev) if fused else None
        self.scope_v = StepScope(dev) if fused else None

    def step(self, x, m, t, eps_d=None, eps=None, eps_adv=None):
        self.vae.train(); self.disc.train()
        self.opt_d.zero_grad()
        if self.scope_d is not None:
            with self.scope_d:
                loss_d = disc_loss(self.vae, self.disc, x, m, t, eps_d)
                with self.scope_d.backward():
                    loss_d.backward()
        else:
            loss_d = disc_loss(self.vae, self.disc, x, m, t, eps_d)
            loss_d.backward()
        if self.distributed:        # CE is batch-MEAN reduced (train.py:55): average the shard gradients
            allreduce_gradients(self.opt_d.flat.grad, group=self.pg, mean=True)
        self.opt_d.step()
        self.opt_vae.zero_grad()
        if self.scope_v is not None:
            with self.scope_v:
                losses = vae_loss(self.vae, self.disc, x, m, t, eps, eps_adv)
                with self.scope_v.backward():
                    losses[0].backward()
        else:
            losses = vae_loss(self.vae, self.disc, x, m, t, eps, eps_adv)
            losses[0].backward()
        if self.distributed:        # sum-reduced terms: SUM (the batchmean confusion term is left per shard)
            allreduce_gradients(self.opt_vae.flat.grad, group=self.pg)
        self.opt_vae.step()
        return loss_d, losses

    def capture(self, B, warmup=3):
        """Both halves of the adversarial step (train.py:41-89) as one CUDA graph over static x, m, t and the three
        reparameterisation noises (discriminator pass, VAE pass, confusion-loss sample)."""
        dev = self.opt_vae.flat.data.device
        Z, M, T = CONFIG["Z_DIM"], CONFIG["M_DIM"], CONFIG["T_DIM"]
        st = dict(x=torch.zeros(B, 1, 28, 28, device=dev), m=torch.zeros(B, M, device=dev),
                  t=torch.zeros(B, T, device=dev), eps_d=torch.zeros(B, Z, device=dev),
                  eps=torch.zeros(B, Z, device=dev), eps_adv=torch.zeros(B, Z, device=dev))
        st["t"][:, 0] = 1.0
        self.graphed = GraphedStep(lambda: self.step(st["x"], st["m"], st["t"], st["eps_d"], st["eps"], st["eps_adv"]),
                                   st, trainer_state([self.vae, self.disc], [self.opt_vae, self.opt_d]), warmup)
        return self.graphed
