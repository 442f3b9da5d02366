from .models import CausalBioVAE  # noqa: F401
