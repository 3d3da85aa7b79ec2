"""Masked Autoregressive Flow (interface of deeprob/flows/models/maf.py:12-78)."""
from typing import Optional

from ...torch.base import DensityEstimator
from ...utils.random import RandomState, check_random_state
from ..layers.autoregressive import AutoregressiveLayer
from ..utils import BatchNormLayer1d
from .base import NormalizingFlow


class MAF(NormalizingFlow):
    def __init__(self, in_features: int, dequantize: bool = False, logit: Optional[float] = None,
                 in_base: Optional[DensityEstimator] = None, n_flows: int = 5, depth: int = 1, units: int = 128,
                 batch_norm: bool = True, activation: str = 'relu', sequential: bool = True,
                 random_state: Optional[RandomState] = None):
        if n_flows <= 0:
            raise ValueError("The number of autoregressive flow layers must be positive")
        if depth <= 0:
            raise ValueError("The number of hidden layers of conditioners must be positive")
        if units <= 0:
            raise ValueError("The number of hidden units per layer must be positive")
        super().__init__(in_features, dequantize=dequantize, logit=logit, in_base=in_base)
        self.n_flows, self.depth, self.units = n_flows, depth, units
        self.batch_norm, self.activation, self.sequential = batch_norm, activation, sequential
        if not sequential:
            random_state = check_random_state(random_state)
        for i in range(n_flows):
            self.layers.append(AutoregressiveLayer(self.in_features, depth, units, activation, reverse=bool(i % 2),
                                                   sequential=sequential, random_state=random_state))
            if batch_norm:
                self.layers.append(BatchNormLayer1d(self.in_features))
