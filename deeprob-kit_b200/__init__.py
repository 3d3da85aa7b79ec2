"""deeprob_kit_b200 -- B200-native (sm_100a) log-likelihood hot path of deeprob-kit.

Same nn.Module API surface as `deeprob.spn.models` / `deeprob.flows` (constructor signatures,
parameter names and shapes, state_dict keys); the arithmetic runs in hand-written CUDA kernels
behind the C ABI declared in include/deeprob_b200.h (libdeeprob_b200.so, loaded with ctypes).
There is no CPU fallback: calling a model on a non-CUDA tensor raises.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (does not load the shared library until first use)
