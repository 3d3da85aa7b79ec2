"""Normalizing-flow base model (interface of deeprob/flows/models/base.py:13-210)."""
from typing import Optional, Tuple

import torch
from torch import distributions, nn

from ...torch.base import DensityEstimator, ProbabilisticModel
from .. import _engine
from ..layers.coupling import CouplingLayer1d
from ..utils import BatchNormLayer1d, DequantizeLayer, LogitLayer


class NormalizingFlow(ProbabilisticModel):
    has_rsample = True

    def __init__(self, in_features, dequantize: bool = False, logit: Optional[float] = None,
                 in_base: Optional[DensityEstimator] = None):
        if isinstance(in_features, torch.Size):
            in_features = tuple(in_features)
            if len(in_features) == 1:
                in_features = in_features[0]
        if not isinstance(in_features, int):
            if not isinstance(in_features, tuple) or len(in_features) != 3:
                raise ValueError("The number of input features must be either an int or a (C, H, W) tuple")
        super().__init__()
        self.in_features = in_features
        self.dequantize = DequantizeLayer(in_features) if dequantize else None
        if logit is not None:
            if logit <= 0.0 or logit >= 1.0:
                raise ValueError("The logit factor must be in (0, 1)")
            self.logit = LogitLayer(in_features, alpha=logit)
        else:
            self.logit = None
        if in_base is None:
            self.in_base_loc = nn.Parameter(torch.zeros(in_features), requires_grad=False)
            self.in_base_scale = nn.Parameter(torch.ones(in_features), requires_grad=False)
            self.in_base = distributions.Normal(self.in_base_loc, self.in_base_scale)
            self._default_base = True
        else:
            self.in_base = in_base
            self._default_base = False
        self.layers = nn.ModuleList()

    def _apply(self, fn, *args, **kwargs):
        # nn.Module._apply swaps the Parameter objects' data; the Normal distribution holds expanded copies
        # made at construction, so rebuild it after .to()/.cuda() (the reference has the same hazard)
        out = super()._apply(fn, *args, **kwargs)
        if self._default_base:
            self.in_base = distributions.Normal(self.in_base_loc, self.in_base_scale)
        return out

    def train(self, mode: bool = True, base_mode: bool = True):
        self.training = mode
        self.layers.train(mode)
        if isinstance(self.in_base, torch.nn.Module):
            self.in_base.train(base_mode)
        return self

    def eval(self):
        return self.train(False, False)

    def preprocess(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        ildj = 0.0
        if self.dequantize is not None and self.logit is not None:
            # both stages in one pass over x (csrc/flows.cu preprocess kernel)
            u, part = _engine.preprocess(x, torch.rand_like(x), self.dequantize.bins, self.logit.alpha)
            return u, part - self.logit.ldj - self.dequantize.ldj
        if self.dequantize is not None:
            x, part = self.dequantize.apply_backward(x)
            ildj = ildj + part
        if self.logit is not None:
            x, part = self.logit.apply_backward(x)
            ildj = ildj + part
        return x, ildj

    def unpreprocess(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        ldj = 0.0
        if self.logit is not None:
            x, part = self.logit.apply_forward(x)
            ldj = ldj + part
        if self.dequantize is not None:
            x, part = self.dequantize.apply_forward(x)
            ldj = ldj + part
        return x, ldj

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Log-likelihood (B,) of complete evidence x."""
        batch = x.shape[0]
        x, ildj = self.preprocess(x)
        x, part = self.apply_backward(x)
        ildj = ildj + part
        if self._default_base:
            # prior log-density, per-sample sum and the log-det add in one kernel
            il = ildj if isinstance(ildj, torch.Tensor) else None
            ll = _engine.normal_prior(x, il, self.in_base_loc.reshape(-1), self.in_base_scale.reshape(-1))
            return ll if il is not None else ll + ildj
        base = self.in_base.log_prob(x)
        return torch.sum(base.view(batch, -1), dim=1) + ildj

    @torch.no_grad()
    def sample(self, n_samples: int, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        n = [n_samples] if isinstance(self.in_base, distributions.Distribution) else n_samples
        x = self.in_base.sample(n)
        x, _ = self.apply_forward(x)
        x, _ = self.unpreprocess(x)
        return x

    def rsample(self, n_samples: int, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not self.in_base.has_rsample:
            raise NotImplementedError("Base distribution must support parametrized sampling")
        n = [n_samples] if isinstance(self.in_base, distributions.Distribution) else n_samples
        x = self.in_base.rsample(n)
        x, _ = self.apply_forward(x)
        x, _ = self.unpreprocess(x)
        return x

    def apply_backward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        total = 0.0
        layers = list(self.layers)
        i = 0
        while i < len(layers):
            layer = layers[i]
            nxt = layers[i + 1] if i + 1 < len(layers) else None
            # coupling + eval-mode batch-norm as one pass (inference fast path, flows/_engine.py coupling1d_infer)
            if isinstance(layer, CouplingLayer1d) and isinstance(nxt, BatchNormLayer1d) and not nxt.training:
                fused = layer.apply_backward_with(x, nxt)
                if fused is not None:
                    x, ildj = fused
                    total = total + ildj
                    i += 2
                    continue
            x, ildj = layer.apply_backward(x)
            total = total + ildj
            i += 1
        return x, total

    def apply_forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        total = 0.0
        for layer in reversed(self.layers):
            x, ldj = layer.apply_forward(x)
            total = total + ldj
        return x, total

    def loss(self, x: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        return -torch.mean(x)
