from .models import CausalViTVAE, ViTVAE  # noqa: F401
from .config import CONFIG  # noqa: F401
