from .config import CONFIG  # noqa: F401
from .models import CausalMorphVAE12, CausalMorphVAE12Prob, LatentDiscriminator  # noqa: F401
