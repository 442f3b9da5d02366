"""Hybrid CNN-stem ViT VAE backbone — same classes, constructor arguments, attribute tree and
state_dict keys as vessel_analysis/00_core/vit_backbone.py:7-199, on the native kernels."""
import torch
import torch.nn as tnn

from .. import _lib as L
from .. import functional as F
from .. import nn


ResBlock = nn.ResBlock


class ViTBlock(tnn.Module):
    """LN -> MHA -> +res ; LN -> MLP(GELU) -> +res  (vit_backbone.py:22-47)."""

    def __init__(self, dim, heads, mlp_dim, dropout=0.1):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = nn.MultiheadAttention(embed_dim=dim, num_heads=heads, dropout=dropout, batch_first=True)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = nn.Sequential(
            nn.Linear(dim, mlp_dim),
            nn.GELU(),
            nn.Dropout(dropout),
            nn.Linear(mlp_dim, dim),
            nn.Dropout(dropout),
        )

    def forward(self, x):
        q = self.norm1(x)
        attn_out, _ = self.attn(q, q, q)
        x, n2 = F.add_layer_norm(x, attn_out, self.norm2.weight, self.norm2.bias, self.norm2.eps)   # residual + norm2
        # mlp = Linear, GELU, Dropout, Linear, Dropout (vit_backbone.py:31-37) with GELU+Dropout and
        # Dropout+residual each fused into one kernel (same masks: the random stream is consumed in the same order)
        lin1, act, drop1, lin2, drop2 = self.mlp[0], self.mlp[1], self.mlp[2], self.mlp[3], self.mlp[4]
        h = F.act_dropout(lin1(n2), L.ACT_GELU, drop1.p, drop1.training)
        return F.dropout_add(lin2(h), x, drop2.p, drop2.training)


class ViTVAE(tnn.Module):
    def __init__(self, in_channels=1, latent_dim=128, img_size=(768, 1280), patch_size=32, embed_dim=256,
                 depth=6, heads=8, mlp_dim=512, res_after=3):
        super().__init__()
        self.latent_dim, self.embed_dim = latent_dim, embed_dim
        self.img_height, self.img_width = img_size
        self.patch_size = patch_size
        chans = [in_channels, 32, 64, 128, embed_dim, embed_dim]
        stem = []
        for i in range(5):
            stem += [nn.Conv2d(chans[i], chans[i + 1], kernel_size=3, stride=2, padding=1),
                     nn.BatchNorm2d(chans[i + 1]), nn.LeakyReLU()]
        self.stem = nn.Sequential(*stem)
        self.grid_h, self.grid_w = self.img_height // 32, self.img_width // 32
        self.num_patches = self.grid_h * self.grid_w
        self.pos_embedding = tnn.Parameter(torch.randn(1, self.num_patches + 1, embed_dim))
        self.cls_token = tnn.Parameter(torch.randn(1, 1, embed_dim))
        self.dropout = nn.Dropout(0.1)
        self.transformer = nn.Sequential(*[ViTBlock(embed_dim, heads, mlp_dim) for _ in range(depth)])
        self.to_latent = nn.LayerNorm(embed_dim)
        self.fc_mu = nn.Linear(embed_dim, latent_dim)
        self.fc_var = nn.Linear(embed_dim, latent_dim)
        self.decoder_input = nn.Linear(latent_dim, embed_dim * self.grid_h * self.grid_w)
        dch = [embed_dim, 128, 64, 32, 16, 16]
        dec = []
        for s in range(5):
            dec += [nn.ConvTranspose2d(dch[s], dch[s + 1], kernel_size=3, stride=2, padding=1, output_padding=1),
                    nn.BatchNorm2d(dch[s + 1]), nn.LeakyReLU()]
            if s < res_after:
                dec.append(ResBlock(dch[s + 1]))
        dec.append(nn.Conv2d(16, in_channels, kernel_size=3, padding=1))
        self.decoder = nn.Sequential(*dec)

    def tokens(self, x):
        """stem -> `b c h w -> b (h w) c` -> prepend cls -> += pos_embedding[:, :n+1] -> dropout
        (vit_backbone.py:161-170)."""
        feat = self.stem(x)
        tok = F.tokens(F.to_nhwc(feat), self.cls_token, self.pos_embedding)
        return self.dropout(tok)

    def encode_cls(self, x):
        tok = self.tokens(x)
        hook = getattr(self, "_tokens_grad_hook", None)
        if hook is not None and tok.requires_grad:
            tok.register_hook(hook)        # fires when the transformer's backward has been enqueued (data-parallel bucket)
        tok = self.transformer(tok)
        return self.to_latent(tok[:, 0])

    def encode(self, x):
        cls_out = self.encode_cls(x)
        return self.fc_mu(cls_out), self.fc_var(cls_out)

    def reparameterize(self, mu, log_var, eps=None):
        return F.reparameterize(mu, log_var, eps)

    def decode(self, z):
        result = self.decoder_input(z)
        result = result.view(-1, self.embed_dim, self.grid_h, self.grid_w)
        return self.decoder(result)

    def forward(self, input, eps=None):
        mu, log_var = self.encode(input)
        z = self.reparameterize(mu, log_var, eps)
        return self.decode(z), input, mu, log_var
