from .ratspn import RatSpn, GaussianRatSpn, BernoulliRatSpn  # noqa: F401
