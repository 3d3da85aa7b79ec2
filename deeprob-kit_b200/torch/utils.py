"""Small torch helpers (interface of deeprob/torch/utils.py:11-135)."""
from typing import Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn, optim

_ACTIVATIONS = {'relu': nn.ReLU, 'leaky-relu': nn.LeakyReLU, 'softplus': nn.Softplus, 'tanh': nn.Tanh,
                'sigmoid': nn.Sigmoid}
_OPTIMIZERS = {'sgd': optim.SGD, 'rmsprop': optim.RMSprop, 'adagrad': optim.Adagrad, 'adam': optim.Adam}


def get_activation_class(name: str):
    if name not in _ACTIVATIONS:
        raise ValueError("Unknown activation function {}".format(name))
    return _ACTIVATIONS[name]


def get_optimizer_class(name: str):
    if name not in _OPTIMIZERS:
        raise ValueError("Unknown optimizer {}".format(name))
    return _OPTIMIZERS[name]


class ScaledTanh(nn.Module):
    """weight * tanh(x); the coupling kernels read `weight` directly (csrc/flows.cu), this forward is the
    stand-alone elementwise form."""

    def __init__(self, weight_size: Union[int, tuple, list] = 1):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(weight_size), requires_grad=True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.weight * torch.tanh(x)


class MaskedLinear(nn.Linear):
    """Linear layer whose weight is multiplied by a fixed 0/1 mask (MADE)."""

    def __init__(self, in_features: int, out_features: int, mask: np.ndarray):
        super().__init__(in_features, out_features)
        if mask.shape[0] != out_features or mask.shape[1] != in_features:
            raise ValueError("Inconsistent mask shape")
        self.register_buffer('mask', torch.tensor(mask, dtype=torch.float32))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.linear(x, self.mask * self.weight, self.bias)


class WeightNormConv2d(nn.Module):
    """Conv2d with weight normalisation (state_dict keys conv.weight_g / conv.weight_v like the reference)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: Union[int, Tuple[int, int]],
                 stride: Union[int, Tuple[int, int]] = 1, padding: Union[int, Tuple[int, int]] = 0, bias: bool = True):
        super().__init__()
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.conv = nn.utils.weight_norm(
                nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=bias)
            )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.conv(x)
