"""Model API boundary (interface of deeprob/torch/base.py:11-49)."""
import abc
from typing import Optional, Union

import torch
from torch import distributions, nn


class ProbabilisticModel(abc.ABC, nn.Module):
    """`log_prob(x)` == `__call__(x)` == `forward(x)`; subclasses implement forward/sample/loss."""

    has_rsample = False

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        return self(x)

    @abc.abstractmethod
    def sample(self, n_samples: int, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        ...

    @abc.abstractmethod
    def loss(self, x: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        ...

    def apply_constraints(self):
        """Project the parameters back on their domain after an optimiser step (no-op by default)."""


DensityEstimator = Union[ProbabilisticModel, distributions.Distribution]
