"""Bijector base class and the parameter-light bijectors (interface of deeprob/flows/utils.py:
squeeze/unsqueeze :11-38, Bijector :41-88, BatchNormLayer1d :91-153, BatchNormLayer2d :156-221,
DequantizeLayer :224-254, LogitLayer :257-294).  Density direction (`apply_backward`) and the inverse
both run on csrc/flows.cu."""
import abc
from typing import Tuple, Union

import numpy as np
import torch
from torch import nn

from . import _engine


def squeeze_depth2d(x: torch.Tensor) -> torch.Tensor:
    """[N, C, H, W] -> [N, 4C, H/2, W/2] (RealNVP squeeze): pure index permutation."""
    n, c, h, w = x.size()
    return x.reshape(n, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(n, c * 4, h // 2, w // 2)


def unsqueeze_depth2d(x: torch.Tensor) -> torch.Tensor:
    """[N, 4C, H/2, W/2] -> [N, C, H, W]: inverse of squeeze_depth2d."""
    n, c, h, w = x.size()
    return x.reshape(n, c // 4, 2, 2, h, w).permute(0, 1, 4, 2, 5, 3).reshape(n, c // 4, h * 2, w * 2)


class Bijector(abc.ABC, nn.Module):
    def __init__(self, in_features: Union[int, Tuple[int, int, int]]):
        if isinstance(in_features, torch.Size):
            in_features = tuple(in_features)
        if not isinstance(in_features, int):
            if not isinstance(in_features, tuple) or len(in_features) != 3:
                raise ValueError("The number of input features must be either an int or a (C, H, W) tuple")
        super().__init__()
        self.in_features = in_features
        self.out_features = in_features

    def forward(self, x: torch.Tensor, backward: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.apply_backward(x) if backward else self.apply_forward(x)

    @abc.abstractmethod
    def apply_backward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """x -> (u, inv-log-det-jacobian): the density-evaluation direction."""

    @abc.abstractmethod
    def apply_forward(self, u: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """u -> (x, log-det-jacobian): the sampling direction."""


class _BatchNormBijector(Bijector):
    _param_shape = None
    _unbiased = True

    def __init__(self, in_features: int, momentum: float = 0.9, eps: float = 1e-5):
        if momentum <= 0.0 or momentum >= 1.0:
            raise ValueError("The momentum value must be in (0, 1)")
        if eps <= 0.0:
            raise ValueError("The epsilon value must be positive")
        super().__init__(in_features)
        self.momentum = momentum
        self.eps = eps
        shape = self._param_shape(in_features)
        self.weight = nn.Parameter(torch.zeros(*shape), requires_grad=True)
        self.bias = nn.Parameter(torch.zeros(*shape), requires_grad=True)
        self.register_buffer('running_var', torch.ones(*shape))
        self.register_buffer('running_mean', torch.zeros(*shape))

    def apply_backward(self, x):
        return _engine.batch_norm(x, self.weight, self.bias, self.running_mean, self.running_var, self.training,
                                  self.momentum, self.eps, self._unbiased, direction=0)

    def apply_forward(self, u):
        return _engine.batch_norm(u, self.weight, self.bias, self.running_mean, self.running_var, False,
                                  self.momentum, self.eps, self._unbiased, direction=1)


class BatchNormLayer1d(_BatchNormBijector):
    """Batch statistics over dim 0 with the unbiased variance (torch.var_mean), flows/utils.py:122-128."""
    _param_shape = staticmethod(lambda f: (1, f))
    _unbiased = True


class BatchNormLayer2d(_BatchNormBijector):
    """Per-channel statistics over (N, H, W) with the biased variance, flows/utils.py:188-195."""
    _param_shape = staticmethod(lambda f: (1, f, 1, 1))
    _unbiased = False


class DequantizeLayer(Bijector):
    def __init__(self, in_features: Union[int, Tuple[int, int, int]], n_bits: int = 8):
        if n_bits <= 0:
            raise ValueError("The number of bits must be positive")
        super().__init__(in_features)
        self.n_bits = n_bits
        self.bins = 2 ** n_bits
        dims = np.prod(self.in_features)
        self.register_buffer('ldj', torch.tensor(dims * np.log(self.bins), dtype=torch.float32))

    def apply_backward(self, x):
        u, _ = _engine.preprocess(x, torch.rand_like(x), self.bins, -1.0)
        return u, -self.ldj.expand(x.shape[0])

    def apply_forward(self, u):
        x = torch.clamp(torch.floor(u * self.bins), min=0, max=self.bins - 1) / (self.bins - 1)
        return x, self.ldj.expand(u.shape[0])


class LogitLayer(Bijector):
    def __init__(self, in_features: Union[int, Tuple[int, int, int]], alpha: float = 0.05):
        if alpha <= 0.0 or alpha >= 1.0:
            raise ValueError("The alpha logit parameter must be in (0, 1)")
        super().__init__(in_features)
        self.alpha = alpha
        dims = np.prod(self.in_features)
        self.register_buffer('ldj', torch.tensor(-dims * np.log(1.0 - 2.0 * alpha), dtype=torch.float32))

    def apply_backward(self, x):
        u, ildj = _engine.preprocess(x, None, 0, self.alpha)      # ildj = -sum(log y + log(1 - y))
        return u, ildj - self.ldj

    def apply_forward(self, u):
        batch = u.shape[0]
        s = torch.sigmoid(u)
        x = (s - self.alpha) / (1.0 - 2.0 * self.alpha)
        v = torch.log(s) + torch.log(1.0 - s)
        return x, torch.sum(v.view(batch, -1), dim=1) + self.ldj
