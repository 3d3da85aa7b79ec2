from .ratspn import RatSpn, GaussianRatSpn, BernoulliRatSpn  # noqa: F401
from .dgcspn import DgcSpn  # noqa: F401
