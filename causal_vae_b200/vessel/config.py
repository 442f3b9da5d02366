"""Default hyper-parameters of the vessel experiment (values of vessel_analysis/00_core/config.py:3-39).
Like the reference, `CONFIG` is a mutable module-level dict read by the models at construction."""
import torch


def _defaults():
    dev = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
    return dict(GPU_ID=0, DEVICE=dev, EPOCHS=150, BATCH_SIZE=8, LEARNING_RATE=1e-4, BETA=0.5,
                LAMBDA_MORPH=10000, IMG_HEIGHT=768, IMG_WIDTH=1280, T_DIM=19, M_DIM=12, Z_DIM=128)


CONFIG = _defaults()
