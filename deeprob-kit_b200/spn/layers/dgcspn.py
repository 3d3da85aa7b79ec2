"""DGC-SPN layers with the reference's names, shapes and state_dict keys (interface of
deeprob/spn/layers/dgcspn.py: SpatialGaussianLayer :14-120, SpatialProductLayer :123-236,
SpatialSumLayer :239-304, SpatialRootLayer :307-355).  Arithmetic: csrc/dgcspn.cu through
deeprob_kit_b200.spn._dgc_engine (one autograd node per layer)."""
from itertools import product as iter_product
from typing import Optional, Tuple, Union

import numpy as np
import torch
from torch import nn

from ... import _lib
from ...torch.initializers import dirichlet_
from .. import _dgc_engine


class _SpatialShape:
    """(C, H, W) accessors shared by the spatial layers."""

    @property
    def in_channels(self) -> int:
        return self.in_features[0]

    @property
    def in_height(self) -> int:
        return self.in_features[1]

    @property
    def in_width(self) -> int:
        return self.in_features[2]

    @property
    def out_channels(self) -> int:
        return self.out_features[0]

    @property
    def out_height(self) -> int:
        return self.out_features[1]

    @property
    def out_width(self) -> int:
        return self.out_features[2]


class SpatialGaussianLayer(_SpatialShape, nn.Module):
    """Per-pixel Gaussian leaves: (B, C_in, H, W) -> (B, out_channels, H, W); NaN inputs are marginalised."""

    def __init__(self, in_features: Tuple[int, int, int], out_channels: int, optimize_scale: bool = False,
                 dropout: Optional[float] = None, quantiles_loc: Optional[np.ndarray] = None,
                 uniform_loc: Optional[Tuple[float, float]] = None):
        if quantiles_loc is not None and uniform_loc is not None:
            raise ValueError("At most one between quantiles_loc and uniform_loc can be specified")
        super().__init__()
        self.in_features = in_features
        self.out_features = (out_channels, in_features[1], in_features[2])
        self.dropout = dropout
        shape = (out_channels, *in_features)
        if quantiles_loc is not None:
            loc = torch.tensor(quantiles_loc, dtype=torch.float32)
        elif uniform_loc is not None:
            low, high = uniform_loc
            loc = torch.linspace(low, high, steps=out_channels).view(-1, 1, 1, 1).repeat(1, *in_features)
        else:
            loc = torch.randn(*shape)
        self.loc = nn.Parameter(loc, requires_grad=True)
        if optimize_scale:
            self.scale = nn.Parameter(0.5 + 0.1 * torch.tanh(torch.randn(*shape)), requires_grad=True)
        else:
            self.scale = nn.Parameter(torch.ones(*shape), requires_grad=False)
        self.distribution = torch.distributions.Normal(self.loc, self.scale, validate_args=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training and self.dropout is not None:
            # layers/dgcspn.py:113-115: NaN dropout on the per-channel log-densities (B, K, C_in, H, W) before
            # nan_to_num and the sum over the input channels, i.e. a dropped (b, k, c, h, w) term contributes 0.
            # One kernel call per input channel (C_in is 1 or 3) keeps every term separately maskable; the
            # Bernoulli draws come from torch's generator exactly like the reference's `torch.rand_like`.
            out = None
            for c in range(x.shape[1]):
                term = _dgc_engine.leaf(x[:, c:c + 1].contiguous(), self.loc[:, c:c + 1], self.scale[:, c:c + 1])
                term = term.masked_fill(torch.rand_like(term) < self.dropout, 0.0)
                out = term if out is None else out + term
            return out
        return _dgc_engine.leaf(x, self.loc, self.scale)


class SpatialProductLayer(_SpatialShape, nn.Module):
    """2x2 (dilated / strided) products of neighbouring pixels in the log domain."""

    def __init__(self, in_features: Tuple[int, int, int], kernel_size: Union[int, Tuple[int, int]], padding: str,
                 stride: Union[int, Tuple[int, int]], dilation: Union[int, Tuple[int, int]], depthwise: bool = True):
        super().__init__()
        self.in_features = in_features
        self.depthwise = depthwise
        self.groups = in_features[0] if depthwise else 1
        as_pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)  # noqa: E731
        kh, kw = as_pair(kernel_size)
        self.stride, self.dilation = as_pair(stride), as_pair(dilation)
        if (kh, kw) != (2, 2):
            raise NotImplementedError("the CUDA product layer implements the 2x2 kernels DGC-SPNs use")
        keh, kew = (kh - 1) * self.dilation[0] + 1, (kw - 1) * self.dilation[1] + 1
        if padding == 'valid':
            self.pad = [0, 0, 0, 0]
        elif padding == 'full':
            self.pad = [kew - 1, kew - 1, keh - 1, keh - 1]
        elif padding == 'final':
            self.pad = [0, (kew - 1) * 2 - self.in_width, 0, (keh - 1) * 2 - self.in_height]
        else:
            raise ValueError("Padding mode must be either 'valid', 'full' or 'final'")
        out_h = int(np.ceil((self.pad[2] + self.pad[3] + self.in_height - keh + 1) / self.stride[0]))
        out_w = int(np.ceil((self.pad[0] + self.pad[1] + self.in_width - kew + 1) / self.stride[1]))
        c = self.in_channels
        out_c = c if depthwise else c ** (kh * kw)
        self.out_features = (out_c, out_h, out_w)

        # The `weight` buffer of the reference (state_dict compatibility); the kernel does not read it
        if depthwise:
            weight = torch.ones(out_c, 1, kh, kw)
        else:
            ids = np.array(list(iter_product(range(c), repeat=kh * kw))).reshape(out_c, 1, kh, kw)
            weight = torch.tensor(np.arange(c).reshape(1, c, 1, 1) == ids, dtype=torch.float32)
        self.register_buffer('weight', weight)

        d = _lib.DgcProductDesc()
        d.channels, d.height, d.width = c, self.in_height, self.in_width
        d.out_channels, d.out_height, d.out_width = out_c, out_h, out_w
        d.pad_top, d.pad_left = self.pad[2], self.pad[0]
        d.stride_h, d.stride_w = self.stride
        d.dilation_h, d.dilation_w = self.dilation
        d.depthwise = 1 if depthwise else 0
        self._desc = d

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _dgc_engine.product(x, self._desc)


class SpatialSumLayer(_SpatialShape, nn.Module):
    """Per-pixel mixtures over the input channels; weight (C_out, C_in, H, W) holds raw logits."""

    def __init__(self, in_features: Tuple[int, int, int], out_channels: int, dropout: Optional[float] = None):
        super().__init__()
        self.in_features = in_features
        self.out_features = (out_channels, in_features[1], in_features[2])
        self.dropout = dropout
        self.weight = nn.Parameter(torch.empty(out_channels, *in_features), requires_grad=True)
        dirichlet_(self.weight, alpha=1.0, dim=1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training and self.dropout is not None:
            # layers/dgcspn.py:297-299: x[torch.rand_like(x) < dropout] = -inf in front of the mixture (out of place
            # here: autograd-friendly; a dropped input gets no gradient, an all-dropped pixel gives -inf)
            x = x.masked_fill(torch.rand_like(x) < self.dropout, float("-inf"))
        return _dgc_engine.mixture(x, self.weight)


class SpatialRootLayer(nn.Module):
    """Class mixtures over every (channel, pixel): (B, C, H, W) -> (B, out_channels)."""

    def __init__(self, in_features: Tuple[int, int, int], out_channels: int):
        super().__init__()
        self.in_features = in_features
        self.out_channels = out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, int(np.prod(in_features))), requires_grad=True)
        dirichlet_(self.weight, alpha=1.0)

    @property
    def in_channels(self) -> int:
        return self.in_features[0]

    @property
    def in_height(self) -> int:
        return self.in_features[1]

    @property
    def in_width(self) -> int:
        return self.in_features[2]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _dgc_engine.root(x, self.weight)
