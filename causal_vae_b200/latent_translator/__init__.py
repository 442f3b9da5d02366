from .models import ViTVAE  # noqa: F401
