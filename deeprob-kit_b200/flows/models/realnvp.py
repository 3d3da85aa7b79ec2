"""RealNVP models (interface of deeprob/flows/models/realnvp.py: RealNVP1d :16-72, RealNVP2d :75-220)."""
from typing import Optional, Tuple

import numpy as np
import torch
from torch import nn

from ...torch.base import DensityEstimator
from ..layers.coupling import CouplingBlock2d, CouplingLayer1d
from ..utils import BatchNormLayer1d
from .base import NormalizingFlow


class RealNVP1d(NormalizingFlow):
    def __init__(self, in_features: int, dequantize: bool = False, logit: Optional[float] = None,
                 in_base: Optional[DensityEstimator] = None, n_flows: int = 5, depth: int = 1, units: int = 128,
                 batch_norm: bool = True, affine: bool = True):
        if n_flows <= 0:
            raise ValueError("The number of coupling flow layers must be positive")
        if depth <= 0:
            raise ValueError("The number of hidden layers of conditioners must be positive")
        if units <= 0:
            raise ValueError("The number of hidden units per layer must be positive")
        super().__init__(in_features, dequantize=dequantize, logit=logit, in_base=in_base)
        self.n_flows, self.depth, self.units = n_flows, depth, units
        self.batch_norm, self.affine = batch_norm, affine
        for i in range(n_flows):                       # masks alternate between consecutive couplings
            self.layers.append(CouplingLayer1d(self.in_features, depth, units, affine=affine, reverse=bool(i % 2)))
            if batch_norm:
                self.layers.append(BatchNormLayer1d(self.in_features))


class RealNVP2d(NormalizingFlow):
    def __init__(self, in_features: Tuple[int, int, int], dequantize: bool = False, logit: Optional[float] = None,
                 in_base: Optional[DensityEstimator] = None, network: str = 'resnet', n_flows: int = 1,
                 n_blocks: int = 2, channels: int = 32, affine: bool = True):
        if n_flows <= 0:
            raise ValueError("The number of coupling flow layers must be positive")
        if n_blocks <= 0:
            raise ValueError("The number of conditioners blocks must be positive")
        if channels <= 0:
            raise ValueError("The number of channels must be positive")
        super().__init__(in_features, dequantize=dequantize, logit=logit, in_base=in_base)
        self.n_flows, self.network, self.n_blocks, self.channels, self.affine = n_flows, network, n_blocks, channels, affine
        self.perm_matrices = torch.nn.ParameterList()
        shape, width = self.in_features, channels
        for _ in range(n_flows):
            self.layers.append(CouplingBlock2d(shape, network, n_blocks, width, affine=affine, last_block=False))
            self.perm_matrices.append(nn.Parameter(self.build_permutation_matrix(shape[0]), requires_grad=False))
            shape = (shape[0] * 2, shape[1] // 2, shape[2] // 2)   # multi-scale: half of the squeezed channels go on
            width *= 2
        self.layers.append(CouplingBlock2d(shape, network, n_blocks, width, affine=affine, last_block=True))

    @staticmethod
    def build_permutation_matrix(channels: int) -> torch.Tensor:
        """0/1 stride-2 conv weights (4C, C, 2, 2) that squeeze with RealNVP's pixel ordering and interleave the
        channels so that `chunk(2)` splits every input channel in two (realnvp.py:142-165)."""
        taps = [(0, 0), (1, 1), (0, 1), (1, 0)]
        weights = np.zeros([channels * 4, channels, 2, 2], dtype=np.float32)
        for c in range(channels):
            for j, (r, s) in enumerate(taps):
                weights[4 * c + j, c, r, s] = 1.0
        order = np.array([4 * c + j for j in range(4) for c in range(channels)])
        return torch.tensor(weights[order], dtype=torch.float32)

    # The reference applies `perm_matrices` with F.conv2d / F.conv_transpose2d (realnvp.py:184,191); the weights
    # are a fixed 0/1 permutation, so the same result is an exact strided gather / scatter (no multiplies,
    # no TF32 rounding): out[:, j*C + c, h, w] = x[:, c, 2h + r_j, 2w + s_j] with taps (0,0),(1,1),(0,1),(1,0).
    _TAPS = ((0, 0), (1, 1), (0, 1), (1, 0))

    @classmethod
    def downscale(cls, x: torch.Tensor) -> torch.Tensor:
        return torch.cat([x[:, :, r::2, s::2] for r, s in cls._TAPS], dim=1)

    @classmethod
    def upscale(cls, y: torch.Tensor) -> torch.Tensor:
        n, c4, h, w = y.shape
        c = c4 // 4
        x = y.new_empty(n, c, 2 * h, 2 * w)
        for j, (r, s) in enumerate(cls._TAPS):
            x[:, :, r::2, s::2] = y[:, j * c:(j + 1) * c]
        return x

    def apply_backward(self, x):
        total, slices = 0.0, []
        last = len(self.layers) - 1
        for i, layer in enumerate(self.layers):
            x, ildj = layer.apply_backward(x)
            total = total + ildj
            if i != last:
                x, z = torch.chunk(self.downscale(x), chunks=2, dim=1)
                slices.append(z)
        for i in range(last - 1, -1, -1):
            x = self.upscale(torch.cat([x, slices[i]], dim=1))
        return x, total

    def apply_forward(self, x):
        total, slices = 0.0, []
        last = len(self.layers) - 1
        for i in range(last):
            x, z = torch.chunk(self.downscale(x), chunks=2, dim=1)
            slices.append(z)
        for i in range(last, -1, -1):
            if i != last:
                x = self.upscale(torch.cat([x, slices[i]], dim=1))
            x, ldj = self.layers[i].apply_forward(x)
            total = total + ldj
        return x, total
