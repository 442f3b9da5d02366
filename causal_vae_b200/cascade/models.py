"""CausalBioVAE — drop-in for causal_cascade/models.py:5-89 (constructor, encode / reparameterize /
forward signatures, submodule names, state_dict keys) on libcvae_b200."""
import torch
import torch.nn as tnn

from .. import functional as F
from .. import nn


class CausalBioVAE(tnn.Module):
    def __init__(self, img_channels=1, m_dim=12, t_dim=19, latent_dim=64):
        super().__init__()
        self.m_dim, self.t_dim, self.latent_dim = m_dim, t_dim, latent_dim
        self.enc_conv = nn.Sequential(
            nn.Conv2d(img_channels, 32, 4, 2, 1), nn.ReLU(),
            nn.Conv2d(32, 64, 4, 2, 1), nn.ReLU(),
            nn.Conv2d(64, 128, 4, 2, 1), nn.ReLU(),
            nn.Conv2d(128, 256, 4, 2, 1), nn.ReLU(),
            nn.AdaptiveAvgPool2d((4, 4)),
            nn.Flatten(),
        )
        self.flatten_dim = 256 * 4 * 4
        self.enc_fc = nn.Sequential(
            nn.Linear(self.flatten_dim + m_dim + t_dim, 512), nn.ReLU(),
            nn.Linear(512, 256), nn.ReLU(),
        )
        self.fc_mu = nn.Linear(256, latent_dim)
        self.fc_logvar = nn.Linear(256, latent_dim)
        self.mechanism_net = nn.Sequential(
            nn.Linear(t_dim, 64), nn.BatchNorm1d(64), nn.ReLU(),
            nn.Linear(64, 64), nn.ReLU(),
            nn.Linear(64, m_dim),
        )
        self.dec_input = nn.Linear(latent_dim + m_dim, self.flatten_dim)
        self.dec_conv = nn.Sequential(
            nn.ConvTranspose2d(256, 128, 4, 2, 1), nn.ReLU(),
            nn.ConvTranspose2d(128, 64, 4, 2, 1), nn.ReLU(),
            nn.ConvTranspose2d(64, 32, 4, 2, 1), nn.ReLU(),
            nn.ConvTranspose2d(32, img_channels, 4, 2, 1),
        )

    def encode(self, x, m, t_onehot):
        h = self.enc_fc(F.cat_pad([self.enc_conv(x), m, t_onehot]))
        return self.fc_mu(h), self.fc_logvar(h)

    def reparameterize(self, mu, logvar, eps=None):
        return F.reparameterize(mu, logvar, eps)

    def decode(self, z, m_hat, out_hw=None):
        """dec_input(cat[z, m_hat]) (z first) -> view(256,4,4) -> dec_conv (models.py:81-86)."""
        out = self.dec_conv(self.dec_input(F.cat_pad([z, m_hat])).view(-1, 256, 4, 4))
        if out_hw is not None and tuple(out.shape[2:]) != tuple(out_hw):
            # the reference's bilinear resize is the identity when the sizes already agree (64x64 config)
            raise RuntimeError(f"bilinear resize {tuple(out.shape[2:])}->{tuple(out_hw)} is outside the B200 hot path; "
                               "feed 64x64 inputs (BASELINE config) ")
        return out

    def forward(self, x, m, t, eps=None):
        t_onehot = F.one_hot(t, self.t_dim)
        mu, logvar = self.encode(x, m, t_onehot)
        z = self.reparameterize(mu, logvar, eps)
        m_hat = self.mechanism_net(t_onehot)
        return self.decode(z, m_hat, x.shape[2:]), m_hat, mu, logvar
