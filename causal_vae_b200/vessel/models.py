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


class CausalVesselVAE(tnn.Module):
    """The CNN variant — drop-in for vessel_analysis/00_core/models.py:9-166: seven Conv4x4/s2 + BN + LeakyReLU(0.2)
    stages (768x1280 -> 6x10, hard-coded like the reference's `enc_flat_dim`), `enc_fc` / `dec_fc` adapters with
    BatchNorm1d, the shared Gaussian P(M|T) head, and a decoder of seven [nearest x2 -> Conv3x3 -> BN -> ReLU] stages
    ending in Conv3x3 -> Sigmoid.  Same attribute names (`enc_conv`, `enc_fc`, `dec_fc`, `dec_conv`,
    `morph_predictor_*`), Sequential indices and state_dict keys; `analyze_vessel.py:93-98` tells the two vessel
    models apart by `hasattr(model, 'dec_adapter')` and reaches into `dec_fc` / `dec_conv` directly."""

    GRID = (6, 10)                                   # models.py:44,163: fixed by the 768x1280 input

    def __init__(self):
        super().__init__()
        self.m_dim, self.t_dim, self.z_dim = CONFIG["M_DIM"], CONFIG["T_DIM"], CONFIG["Z_DIM"]
        widths = (1, 32, 64, 128, 256, 512, 512, 512)
        enc = []
        for cin, cout in zip(widths[:-1], widths[1:]):
            enc += [nn.Conv2d(cin, cout, 4, 2, 1), nn.BatchNorm2d(cout), nn.LeakyReLU(0.2)]
        self.enc_conv = nn.Sequential(*enc, nn.Flatten())
        self.enc_flat_dim = 512 * self.GRID[0] * self.GRID[1]
        self.enc_fc = nn.Sequential(nn.Linear(self.enc_flat_dim + self.m_dim + self.t_dim, 1024), nn.BatchNorm1d(1024),
                                    nn.LeakyReLU(0.2), nn.Linear(1024, self.z_dim * 2))
        self.morph_predictor_shared = nn.Sequential(
            nn.Linear(self.t_dim, 64), nn.LeakyReLU(0.2), nn.Linear(64, 64), nn.LeakyReLU(0.2))
        self.morph_predictor_mu = nn.Linear(64, self.m_dim)
        self.morph_predictor_logvar = nn.Linear(64, self.m_dim)
        self.dec_fc = nn.Sequential(nn.Linear(self.m_dim + self.z_dim, 1024), nn.BatchNorm1d(1024), nn.LeakyReLU(0.2),
                                    nn.Linear(1024, self.enc_flat_dim), nn.ReLU())
        up = (512, 512, 512, 256, 128, 64, 32)
        dec = []
        for cin, cout in zip(up, up[1:]):
            dec += [nn.Upsample(scale_factor=2, mode="nearest"), nn.Conv2d(cin, cout, 3, 1, 1), nn.BatchNorm2d(cout),
                    nn.ReLU()]
        self.dec_conv = nn.Sequential(*dec, nn.Upsample(scale_factor=2, mode="nearest"), nn.Conv2d(32, 1, 3, 1, 1),
                                      nn.Sigmoid())

    def reparameterize(self, mu, logvar, eps=None):
        return F.reparameterize(mu, logvar, eps)

    def morph_head(self, t):
        h = self.morph_predictor_shared(t)
        return self.morph_predictor_mu(h), F.clamp(self.morph_predictor_logvar(h), -10.0, 10.0)

    def decode(self, m, z):
        """dec_conv(dec_fc(cat[m, z]).view(-1, 512, 6, 10)) — models.py:161-164."""
        return self.dec_conv(self.dec_fc(F.cat_pad([m, z])).view(-1, 512, *self.GRID))

    def forward(self, x, m, t, eps=None):
        h = self.enc_fc(F.cat_pad([self.enc_conv(x), m, t]))
        if eps is None:
            eps = torch.randn(h.shape[0], self.z_dim, device=h.device, dtype=h.dtype)
        mu, logvar, z = F.latent(h, eps, mu_clamp=100.0, lv_clamp=10.0)
        m_mu, m_logvar = self.morph_head(t)
        return self.decode(m, z), m_mu, mu, logvar, m_mu, m_logvar
