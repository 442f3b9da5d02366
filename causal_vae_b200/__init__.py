"""causal_vae_b200 — B200-native (sm_100a CUDA behind a C ABI) CausalVAE training and counterfactual
hot path with the module API of bjo5029/causal-vae's models.py files.

Sub-packages mirror the reference's experiment directories:
    vessel/            vessel_analysis/00_core   (CausalViTVAE, ViTVAE backbone, loss_function)
    latent_translator/ latent_translator         (ViTVAE, train step)
    cascade/           causal_cascade            (CausalBioVAE, loss_function)
    mnist/             mnist_test/01 + 06        (CausalMorphVAE12, LatentDiscriminator)
Importing the package loads libcvae_b200.so and fails loudly if it has not been built.
"""
from . import _lib  # noqa: F401  (raises ImportError when the native library is missing)
from . import nn, functional  # noqa: F401

__all__ = ["nn", "functional"]
