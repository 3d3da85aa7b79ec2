"""Parameter constraints (interface of deeprob/torch/constraints.py:7-27)."""
import torch
from torch import nn


class ScaleClipper(nn.Module):
    """Clamp `module.scale` to >= eps in place (called from `apply_constraints`)."""

    def __init__(self, eps: float = 1e-5):
        if eps <= 0.0:
            raise ValueError("The epsilon value must be positive")
        super().__init__()
        self.register_buffer('eps', torch.tensor(eps))

    def forward(self, module: nn.Module):
        with torch.no_grad():
            module.scale.clamp_(min=self.eps)
