"""ViTVAE of latent_translator/models.py:40-126 — the vessel backbone family with a ResBlock after
every one of the first FOUR up-stages (decoder indices {3,7,11,15}, final conv at 19), 512-d
fc_mu / fc_var heads and the 384x640 default image size.  Same constructor and state_dict keys."""
from ..vessel import vit_backbone as _vb

ResBlock = _vb.ResBlock
ViTBlock = _vb.ViTBlock   # attn(norm1(x), norm1(x), norm1(x)) (models.py:36) == one LayerNorm + self-attention


class ViTVAE(_vb.ViTVAE):
    def __init__(self, in_channels=1, latent_dim=512, img_size=(384, 640), patch_size=32, embed_dim=256,
                 depth=6, heads=8, mlp_dim=512):
        super().__init__(in_channels=in_channels, latent_dim=latent_dim, img_size=img_size, patch_size=patch_size,
                         embed_dim=embed_dim, depth=depth, heads=heads, mlp_dim=mlp_dim, res_after=4)
