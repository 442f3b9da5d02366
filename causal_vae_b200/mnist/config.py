"""Hyper-parameters of the MNIST experiments (values of mnist_test/01_baseline_causal_vae/config.py:6-17).
A mutable module-level dict read by the models at construction, like the reference's."""
import torch

CONFIG = {
    "BATCH_SIZE": 128, "EPOCHS": 100, "LR": 1e-3, "Z_DIM": 10, "M_DIM": 12, "T_DIM": 10,
    "DEVICE": torch.device("cuda" if torch.cuda.is_available() else "cpu"), "SEED": 42,
    "BETA": 1.0, "LAMBDA_ADV": 10.0,
}
