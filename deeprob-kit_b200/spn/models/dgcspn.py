"""DGC-SPN model (interface of deeprob/spn/models/dgcspn.py:15-201): same constructor, ValueErrors,
layer schedule (:100-125) and state_dict keys; layers run on csrc/dgcspn.cu."""
from typing import List, Optional, Tuple, Union

import numpy as np
import os

import torch
import torch.nn.functional as F
from torch import autograd

from ...torch.base import ProbabilisticModel
from ...torch.constraints import ScaleClipper
from .. import _dgc_engine
from ..layers.dgcspn import SpatialGaussianLayer, SpatialProductLayer, SpatialRootLayer, SpatialSumLayer


class DgcSpn(ProbabilisticModel):
    def __init__(
        self,
        in_features: Tuple[int, int, int],
        out_classes: int = 1,
        n_batch: int = 8,
        sum_channels: int = 8,
        depthwise: Union[bool, List[bool]] = False,
        n_pooling: int = 0,
        optimize_scale: bool = False,
        in_dropout: Optional[float] = None,
        sum_dropout: Optional[float] = None,
        quantiles_loc: Optional[np.ndarray] = None,
        uniform_loc: Optional[Tuple[float, float]] = None
    ):
        if in_features[1] != in_features[2]:
            raise ValueError("The height and width of input size must be the same")
        if out_classes <= 0:
            raise ValueError("The number of output classes must be positive")
        if n_batch <= 0:
            raise ValueError("The number of base distribution batches must be positive")
        if sum_channels <= 0:
            raise ValueError("The number of output channels of spatial sum layers must be positive")
        if in_dropout is not None and not 0.0 < in_dropout < 1.0:
            raise ValueError("The dropout rate at base distribution must be in (0, 1)")
        if sum_dropout is not None and not 0.0 < sum_dropout < 1.0:
            raise ValueError("The dropout rate at spatial sum layers must be in (0, 1)")
        if quantiles_loc is not None and uniform_loc is not None:
            raise ValueError("At least one between quantiles_loc and uniform_loc must be None")
        if quantiles_loc is not None and len(quantiles_loc.shape) != 4:
            raise ValueError("The mean quantiles must be a 4D Numpy array")
        if uniform_loc is not None and (len(uniform_loc) != 2 or uniform_loc[0] >= uniform_loc[1]):
            raise ValueError("The uniform range must be a pair (A, B) with A < B")

        depth = int(np.ceil(np.log2(in_features[1])))
        if isinstance(depthwise, bool):
            depthwise = [depthwise] * (depth + 1)
        else:
            if len(depthwise) == 0 or len(depthwise) > depth + 1:
                raise ValueError("The length of depthwise argument must be in [1, ceil(log2(D)) + 1]")
            depthwise = list(depthwise) + [depthwise[-1]] * (depth + 1 - len(depthwise))
        if n_pooling < 0 or n_pooling > depth:
            raise ValueError("The number of initial pooling spatial product layers must be in [0, ceil(log2(D))]")

        super().__init__()
        self.in_features = in_features
        self.out_classes = out_classes
        self.n_batch = n_batch
        self.sum_channels = sum_channels
        self.depthwise = depthwise
        self.n_pooling = n_pooling
        self.optimize_scale = optimize_scale
        self.in_dropout = in_dropout
        self.sum_dropout = sum_dropout
        self.layers = torch.nn.ModuleList()

        self.base_layer = SpatialGaussianLayer(
            in_features, n_batch, optimize_scale=optimize_scale, dropout=in_dropout,
            quantiles_loc=quantiles_loc, uniform_loc=uniform_loc
        )
        shape = self.base_layer.out_features
        for level in range(depth + 1):
            if level < n_pooling:      # pooling products halve the resolution
                spec = dict(padding='valid', stride=(2, 2), dilation=(1, 1))
            else:                      # then dilations 1, 2, 4, ... ; the last level collapses to one root pixel grid
                k = 2 ** (level - n_pooling)
                spec = dict(padding='final' if level == depth else 'full', stride=(1, 1), dilation=(k, k))
            prod = SpatialProductLayer(shape, kernel_size=(2, 2), depthwise=depthwise[level], **spec)
            self.layers.append(prod)
            shape = prod.out_features
            if level != depth:
                mix = SpatialSumLayer(shape, sum_channels, sum_dropout)
                self.layers.append(mix)
                shape = mix.out_features
        self.root_layer = SpatialRootLayer(shape, out_classes)
        if optimize_scale:
            self.scale_clipper = ScaleClipper()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Log-likelihood (B, out_classes) of x (B, C, H, W); NaN = marginalised variable."""
        h = self.base_layer(x)
        layers = list(self.layers)
        # a depthwise product layer and the sum layer behind it run as one kernel (its output, the largest tensor of
        # the model, is never written); with gradients the pair is one autograd node that recomputes the product
        # output in its backward (DPK_DGC_FUSE_TRAIN=0: every layer its own node)
        grad = torch.is_grad_enabled() and (h.requires_grad or any(p.requires_grad for p in self.parameters()))
        fuse = os.environ.get("DPK_DGC_FUSE", "1") != "0" and (not grad or os.environ.get("DPK_DGC_FUSE_TRAIN", "1") != "0")
        i = 0
        while i < len(layers):
            layer = layers[i]
            nxt = layers[i + 1] if i + 1 < len(layers) else None
            if (fuse and isinstance(layer, SpatialProductLayer) and isinstance(nxt, SpatialSumLayer)
                    and not (nxt.training and nxt.dropout is not None)
                    and _dgc_engine.can_fuse_product_mixture(layer, nxt)):
                if grad:
                    h = _dgc_engine.product_mixture_train(h, layer._desc, nxt.weight)
                else:
                    h = _dgc_engine.product_mixture(h, layer._desc, nxt.weight)
                i += 2
            else:
                h = layer(h)
                i += 1
        return self.root_layer(h)

    def mpe(self, x: torch.Tensor) -> torch.Tensor:
        """Gradient-based MPE completion of the NaN entries (needs grad mode, like the reference)."""
        z = self.base_layer(x)
        if not z.requires_grad:
            z.requires_grad = True
        y = z
        for layer in self.layers:
            y = layer(y)
        y = self.root_layer(y)
        z_grad, = autograd.grad(y, z, grad_outputs=torch.ones_like(y), only_inputs=True)
        with torch.no_grad():
            estimates = torch.sum(z_grad.unsqueeze(2) * self.base_layer.loc, dim=1)
            return torch.where(torch.isnan(x), estimates, x)

    def sample(self, n_samples: int, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        raise NotImplementedError("Sampling is not implemented for DGC-SPNs")

    def loss(self, x: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        """models/dgcspn.py:189-196 as one kernel (csrc/train_glue.cu)."""
        from .._engine import nll_loss
        return nll_loss(x, None if self.out_classes == 1 else y)

    def apply_constraints(self):
        if self.optimize_scale:
            self.scale_clipper(self.base_layer)
