"""Masked autoregressive layer (interface of deeprob/flows/layers/autoregressive.py:13-181): MADE conditioner
(tcgen05 GEMMs over the cached masked weights in inference, library GEMMs with gradients) + the fused affine/log-det kernel of csrc/flows.cu in the density direction;
the sampling direction is the reference's D-step sequential loop."""
from typing import List, Optional, Tuple

import numpy as np
import torch
from torch import nn

from ...torch.utils import MaskedLinear, ScaledTanh, get_activation_class
from .. import _engine
from ..utils import Bijector


class AutoregressiveLayer(Bijector):
    def __init__(self, in_features: int, depth: int, units: int, activation: str, reverse: bool = False,
                 sequential: bool = True, random_state: Optional[np.random.RandomState] = None):
        if depth <= 0:
            raise ValueError("The depth value must be positive")
        if units <= 0:
            raise ValueError("The units value must be positive")
        if not sequential and not isinstance(random_state, np.random.RandomState):
            raise ValueError("A Numpy RandomState is required if sequential is False")
        activation_cls = get_activation_class(activation)
        super().__init__(in_features)
        self.layers = nn.ModuleList()
        self.scale_act = ScaledTanh()
        if sequential:
            degrees = self.build_degrees_sequential(depth, units, reverse)
        else:
            degrees = self.build_degrees_random(depth, units, random_state)
        masks = self.build_masks(degrees)
        self.ordering = degrees[0]
        self.inv_ordering = np.argsort(self.ordering)
        layers, fan_in = [], in_features
        for mask in masks[:-1]:
            layers += [MaskedLinear(fan_in, units, mask), activation_cls()]
            fan_in = units
        layers.append(MaskedLinear(fan_in, self.in_features * 2, np.tile(masks[-1], reps=(2, 1))))
        self.network = nn.Sequential(*layers)

    def apply_backward(self, x):
        # inference on a large batch: the MADE layers (mask * weight cached) run on the tcgen05 GEMM
        z = _engine.mlp(self.network, x)
        return _engine.coupling(x, z, self.scale_act.weight, None, self.in_features, 0, True, 0, 1)

    def _conditioner(self, x):
        t, s = torch.chunk(self.network(x), chunks=2, dim=1)
        return t, self.scale_act(s)

    def apply_forward(self, u):
        if torch.is_grad_enabled():      # differentiable (slower) variant, like the reference
            cols = list(torch.unbind(torch.zeros_like(u), dim=1))
            ldj = list(torch.unbind(torch.zeros_like(u), dim=1))
            for i in self.inv_ordering:
                t, s = self._conditioner(torch.stack(cols, dim=1))
                cols[i] = u[:, i] * torch.exp(s[:, i]) + t[:, i]
                ldj[i] = s[:, i]
            return torch.stack(cols, dim=1), torch.sum(torch.stack(ldj, dim=1), dim=1)
        x = torch.zeros_like(u)
        ldj = torch.zeros_like(u)
        for i in self.inv_ordering:
            t, s = self._conditioner(x)
            x[:, i] = u[:, i] * torch.exp(s[:, i]) + t[:, i]
            ldj[:, i] = s[:, i]
        return x, torch.sum(ldj, dim=1)

    def build_degrees_sequential(self, depth: int, units: int, reverse: bool) -> List[np.ndarray]:
        first = np.arange(self.in_features - 1, -1, -1) if reverse else np.arange(self.in_features)
        return [first] + [np.arange(units) % (self.in_features - 1) for _ in range(depth)]

    def build_degrees_random(self, depth: int, units: int, random_state: np.random.RandomState) -> List[np.ndarray]:
        ordering = np.arange(self.in_features)
        random_state.shuffle(ordering)
        degrees = [ordering]
        for _ in range(depth):
            degrees.append(random_state.randint(np.min(degrees[-1]), self.in_features - 1, units))
        return degrees

    @staticmethod
    def build_masks(degrees: List[np.ndarray]) -> List[np.ndarray]:
        masks = [np.less_equal(d1[None, :], d2[:, None]) for d1, d2 in zip(degrees[:-1], degrees[1:])]
        masks.append(np.less(degrees[-1][None, :], degrees[0][:, None]))
        return masks
