from .base import NormalizingFlow  # noqa: F401
from .realnvp import RealNVP1d, RealNVP2d  # noqa: F401
from .maf import MAF  # noqa: F401
