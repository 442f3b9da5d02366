"""CausalMorphVAE12 / LatentDiscriminator — drop-ins for mnist_test/01_baseline_causal_vae/models.py:6-111
and the probabilistic-morphology variant mnist_test/06_model_experiment/models.py:6-85 (same
constructors, forward tuple arities, submodule names and state_dict keys), on libcvae_b200.

`forward` and `reparameterize` accept an optional explicit `eps` (the reference draws it with
torch.randn_like inside; parity tests inject it)."""
import torch
import torch.nn as tnn

from .. import functional as F
from .. import nn

try:                                  # when dropped into the reference's script directories
    from config import CONFIG        # noqa: F401  (mnist_test/*/config.py)
    if "Z_DIM" not in CONFIG or "IMG_HEIGHT" in CONFIG:
        raise ImportError
except ImportError:
    from .config import CONFIG


class _MorphVAEBase(tnn.Module):
    def __init__(self):
        super().__init__()
        self.m_dim, self.t_dim, self.z_dim = CONFIG["M_DIM"], CONFIG["T_DIM"], CONFIG["Z_DIM"]
        self.enc_conv = nn.Sequential(
            nn.Conv2d(1, 32, 4, 2, 1), nn.ReLU(),
            nn.Conv2d(32, 64, 4, 2, 1), nn.ReLU(),
            nn.Flatten(),
        )
        self.enc_flat_dim = 64 * 7 * 7
        self.enc_fc = nn.Sequential(
            nn.Linear(self.enc_flat_dim + self.m_dim + self.t_dim, 512), nn.ReLU(),
            nn.Linear(512, self.z_dim * 2),
        )

    def _decoder(self):
        self.dec_fc = nn.Sequential(nn.Linear(self.m_dim + self.z_dim, self.enc_flat_dim), nn.ReLU())
        self.dec_conv = nn.Sequential(
            nn.ConvTranspose2d(64, 32, 4, 2, 1), nn.ReLU(),
            nn.ConvTranspose2d(32, 1, 4, 2, 1), nn.Sigmoid(),
        )

    def reparameterize(self, mu, logvar, eps=None):
        return F.reparameterize(mu, logvar, eps)

    def encode(self, x, m, t, eps=None):
        """enc_conv -> cat[x_feat, m, t] -> enc_fc -> chunk -> reparameterise (models.py:58-61)."""
        h = self.enc_fc(F.cat_pad([self.enc_conv(x), m, t]))
        if eps is None:
            eps = torch.randn(h.shape[0], self.z_dim, device=h.device, dtype=h.dtype)
        return F.latent(h, eps)

    def decode(self, m, z):
        """dec_fc(cat[m, z]) -> view(64,7,7) -> dec_conv (models.py:66-70)."""
        h = self.dec_fc(F.cat_pad([m, z]))
        return self.dec_conv(h.view(-1, 64, 7, 7))


class CausalMorphVAE12(_MorphVAEBase):
    """T -> M -> X with a deterministic morphology predictor; decodes from m_hat (models.py:6-72)."""

    def __init__(self):
        super().__init__()
        self.morph_predictor = nn.Sequential(nn.Linear(self.t_dim, 128), nn.ReLU(), nn.Linear(128, self.m_dim))
        self._decoder()

    def forward(self, x, m, t, eps=None):
        mu, logvar, z = self.encode(x, m, t, eps)
        m_hat = self.morph_predictor(t)
        return self.decode(m_hat, z), m_hat, mu, logvar


class CausalMorphVAE12Prob(_MorphVAEBase):
    """06_model_experiment variant: Gaussian P(M|T) heads, decoder fed the REAL m, 6-tuple output
    (mnist_test/06_model_experiment/models.py:34-85).  Exported there as `CausalMorphVAE12`."""

    def __init__(self):
        super().__init__()
        self.morph_predictor_shared = nn.Sequential(nn.Linear(self.t_dim, 128), nn.ReLU())
        self.morph_predictor_mu = nn.Linear(128, self.m_dim)
        self.morph_predictor_logvar = nn.Linear(128, self.m_dim)
        self._decoder()

    def morph_predictor(self, t):
        return self.morph_predictor_mu(self.morph_predictor_shared(t))

    def forward(self, x, m, t, eps=None):
        mu, logvar, z = self.encode(x, m, t, eps)
        h = self.morph_predictor_shared(t)
        m_mu, m_logvar = self.morph_predictor_mu(h), self.morph_predictor_logvar(h)
        return self.decode(m, z), m_mu, mu, logvar, m_mu, m_logvar


class LatentDiscriminator(tnn.Module):
    """z -> treatment logits (models.py:93-111)."""

    def __init__(self):
        super().__init__()
        self.z_dim, self.t_dim = CONFIG["Z_DIM"], CONFIG["T_DIM"]
        self.net = nn.Sequential(
            nn.Linear(self.z_dim, 64), nn.LeakyReLU(0.2),
            nn.Linear(64, 64), nn.LeakyReLU(0.2),
            nn.Linear(64, self.t_dim),
        )

    def forward(self, z):
        return self.net(z)


class SimpleClassifier(tnn.Module):
    """Name kept so that `from models import CausalMorphVAE12, SimpleClassifier, LatentDiscriminator`
    (mnist_test/01_baseline_causal_vae/train.py:9) imports.  The class itself — the external evaluation classifier
    (models.py:74-91: two 5x5 convolutions with max-pooling, trained with SGD on real MNIST by
    `train_external_classifier`, train.py:105-128) — is outside the accelerated path (SURVEY 8a lists a11-a13 only) and
    has no native kernels yet (5x5 taps, max-pool, NLL); constructing it fails loudly instead of silently running on
    ATen."""

    def __init__(self):
        super().__init__()
        raise RuntimeError("SimpleClassifier (external evaluation classifier, mnist_test/*/models.py:74-91) is not part of "
                           "the B200 hot path and has no native implementation; use the reference's class for it")
