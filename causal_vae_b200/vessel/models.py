"""CausalViTVAE — drop-in for vessel_analysis/00_core/models.py:181-307 (same constructor, forward
6-tuple, submodule names and state_dict keys), forward/backward on libcvae_b200."""
import torch
import torch.nn as tnn

from .. import functional as F
from .. import nn
from .vit_backbone import ViTVAE

try:                                  # when dropped into the reference's script directories
    from config import CONFIG        # noqa: F401  (vessel_analysis/00_core/config.py)
    if "IMG_HEIGHT" not in CONFIG:
        raise ImportError
except ImportError:
    from .config import CONFIG


class CausalViTVAE(tnn.Module):
    """(X, M, T) -> Z -> (M, Z) -> X with a hybrid-ViT backbone (models.py:181-250)."""

    def __init__(self, pretrained_path=None):
        super().__init__()
        self.backbone = ViTVAE(img_size=(CONFIG["IMG_HEIGHT"], CONFIG["IMG_WIDTH"]), patch_size=32,
                               embed_dim=256, depth=6, heads=8, mlp_dim=512, latent_dim=512)
        if pretrained_path:
            print(f"[CausalViTVAE] Loading backbone weights from {pretrained_path}")
            state_dict = torch.load(pretrained_path, map_location=CONFIG["DEVICE"])
            self.backbone.load_state_dict(state_dict, strict=False)
        self.vit_embed_dim, self.vit_latent_dim = 256, 512
        self.my_z_dim, self.m_dim, self.t_dim = CONFIG["Z_DIM"], CONFIG["M_DIM"], CONFIG["T_DIM"]
        self.enc_adapter = nn.Sequential(
            nn.Linear(self.vit_embed_dim + self.m_dim + self.t_dim, 512),
            nn.BatchNorm1d(512),
            nn.LeakyReLU(0.2),
            nn.Linear(512, self.my_z_dim * 2),
        )
        self.dec_adapter = nn.Sequential(
            nn.Linear(self.my_z_dim + self.m_dim, 256),
            nn.BatchNorm1d(256),
            nn.LeakyReLU(0.2),
            nn.Linear(256, self.vit_latent_dim),
        )
        self.morph_predictor_shared = nn.Sequential(
            nn.Linear(self.t_dim, 64), nn.LeakyReLU(0.2), nn.Linear(64, 64), nn.LeakyReLU(0.2))
        self.morph_predictor_mu = nn.Linear(64, self.m_dim)
        self.morph_predictor_logvar = nn.Linear(64, self.m_dim)

    def reparameterize(self, mu, logvar, eps=None):
        return F.reparameterize(mu, logvar, eps)

    def encode(self, x, m, t, eps=None):
        """backbone CLS feature -> enc_adapter(cat[cls, m, t]) -> chunk, clamp, reparameterise
        (models.py:262-288).  Returns (mu, logvar, z)."""
        cls_out = self.backbone.encode_cls(x)
        h = self.enc_adapter(F.cat_pad([cls_out, m, t]))
        if eps is None:
            eps = torch.randn(h.shape[0], self.my_z_dim, device=h.device, dtype=h.dtype)
        return F.latent(h, eps, mu_clamp=100.0, lv_clamp=10.0)

    def morph_head(self, t):
        h = self.morph_predictor_shared(t)
        return self.morph_predictor_mu(h), F.clamp(self.morph_predictor_logvar(h), -10.0, 10.0)

    def decode(self, m, z):
        """backbone.decode(dec_adapter(cat[m, z])) — m first (models.py:299-305)."""
        h = self.dec_adapter(F.cat_pad([m, z]))
        hook = getattr(self, "_decoder_grad_hook", None)
        if hook is not None and h.requires_grad:
            h.register_hook(hook)      # fires when the whole decoder's backward has been enqueued (data-parallel overlap)
        return self.backbone.decode(h)

    def forward(self, x, m, t, eps=None):
        mu, logvar, z = self.encode(x, m, t, eps)
        m_mu, m_logvar = self.morph_head(t)
        recon_x = self.decode(m, z)
        return recon_x, m_mu, mu, logvar, m_mu, m_logvar
