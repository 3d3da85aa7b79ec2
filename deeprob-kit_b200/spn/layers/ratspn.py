"""RAT-SPN layers with the reference's names, shapes and state_dict keys
(interface of deeprob/spn/layers/ratspn.py: RegionGraphLayer :14-157, GaussianLayer :160-213,
BernoulliLayer :216-247, ProductLayer :250-330, SumLayer :333-417, RootLayer :420-490).

The modules own the parameters/buffers; the arithmetic of the whole stack runs in the fused CUDA
path driven by `deeprob_kit_b200.spn.models.ratspn.RatSpn.forward` (csrc/ratspn_*.cu).  Calling a
single layer runs the same kernels level by level through the stand-alone C-ABI entry points
(inference only: gradients flow through the model-level call).
"""
import abc
from typing import List, Optional, Tuple

import numpy as np
import torch
from torch import distributions, nn

from ... import _lib
from ...torch.initializers import dirichlet_
from .. import _engine


class RegionGraphLayer(abc.ABC, nn.Module):
    """Leaf layer over the regions of a region graph: (B, D) -> (B, n_regions, out_channels)."""

    leaf_kind = None  # _lib.LEAF_*

    def __init__(self, in_features: int, out_channels: int, regions: List[tuple], rg_depth: int,
                 dropout: Optional[float] = None, **kwargs):
        super().__init__()
        self.in_features = in_features
        self.in_regions = len(regions)
        self.out_channels = out_channels
        self.rg_depth = rg_depth
        self.dropout = dropout
        self.distribution = None

        n_slots = 2 ** rg_depth
        self.pad = -in_features % n_slots
        self.dimension = (in_features + self.pad) // n_slots

        # Gather table: short regions repeat their last variable, the repeats are flagged in pad_mask
        table = np.empty((self.in_regions, self.dimension), dtype=np.int64)
        lengths = np.empty(self.in_regions, dtype=np.int32)
        for g, region in enumerate(regions):
            lengths[g] = len(region)
            table[g, :len(region)] = region
            table[g, len(region):] = region[-1]
        if self.pad > 0:
            flags = np.arange(self.dimension)[None, None, :] >= lengths[:, None, None]
            self.register_buffer('pad_mask', torch.from_numpy(flags))
        self.register_buffer('mask', torch.from_numpy(table))

        # Per-repetition inverse permutation (used when sampling)
        inv_mask = torch.argsort(self.mask.reshape(-1, in_features + self.pad), dim=1)
        self.register_buffer('inv_mask', inv_mask)
        if self.pad > 0:
            flat = self.pad_mask.reshape(-1, in_features + self.pad)
            self.register_buffer('inv_pad_mask', torch.gather(flat, 1, self.inv_mask))

        # Device-side copies the kernels read (not part of the state_dict)
        self.register_buffer('_mask_i32', torch.from_numpy(table.astype(np.int32)), persistent=False)
        self.register_buffer('_region_len', torch.from_numpy(lengths), persistent=False)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        # keep the int32 gather table in sync with a loaded `mask`
        self._mask_i32.copy_(self.mask.to(torch.int32))

    @abc.abstractmethod
    def leaf_parameters(self) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """(p0, p1): (loc, scale) or (logits, None)."""

    @abc.abstractmethod
    def distribution_mode(self) -> torch.Tensor:
        ...

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training and self.dropout is not None:
            raise NotImplementedError("stand-alone leaf layer call with dropout: call the model instead")
        return _engine.ratspn_leaf_forward(self, x)

    def unpad_samples(self, x: torch.Tensor, idx_group: torch.Tensor) -> torch.Tensor:
        n = idx_group.shape[0]
        rep = torch.div(idx_group[:, 0], 2 ** self.rg_depth, rounding_mode='floor')
        out = torch.gather(x, 1, self.inv_mask[rep])
        if self.pad > 0:
            out = out[~self.inv_pad_mask[rep]].view(n, self.in_features)
        return out

    @torch.no_grad()
    def mpe(self, x: torch.Tensor, idx_group: torch.Tensor, idx_offset: torch.Tensor) -> torch.Tensor:
        mode = self.distribution_mode()
        picked = mode[idx_group, idx_offset].flatten(1)
        picked = self.unpad_samples(picked, idx_group)
        return torch.where(torch.isnan(x), picked, x)

    @torch.no_grad()
    def sample(self, idx_group: torch.Tensor, idx_offset: torch.Tensor) -> torch.Tensor:
        n = idx_group.shape[0]
        draws = self.distribution.sample([n])
        rows = torch.arange(n, device=idx_group.device).unsqueeze(1)
        picked = draws[rows, idx_group, idx_offset].flatten(1)
        return self.unpad_samples(picked, idx_group)


class GaussianLayer(RegionGraphLayer):
    leaf_kind = _lib.LEAF_GAUSSIAN

    def __init__(self, in_features: int, out_channels: int, regions: List[tuple], rg_depth: int,
                 dropout: Optional[float] = None, uniform_loc: Optional[Tuple[float, float]] = None,
                 optimize_scale: bool = False):
        super().__init__(in_features, out_channels, regions, rg_depth, dropout)
        shape = (self.in_regions, self.out_channels, self.dimension)
        if uniform_loc is None:
            loc = torch.randn(*shape)
        else:
            low, high = uniform_loc
            loc = low + (high - low) * torch.rand(*shape)
        self.loc = nn.Parameter(loc, requires_grad=True)
        if optimize_scale:
            self.scale = nn.Parameter(0.5 + 0.1 * torch.tanh(torch.randn(*shape)), requires_grad=True)
        else:
            self.scale = nn.Parameter(torch.ones(*shape), requires_grad=False)
        self.distribution = distributions.Normal(self.loc, self.scale, validate_args=False)

    def leaf_parameters(self):
        return self.loc, self.scale

    def unit_scale(self) -> bool:
        """True when `scale` is frozen at 1 everywhere (the default, `optimize_scale=False`): the kernels
        then take the cheaper t = x - mu path.  One device->host sync per parameter version, cached."""
        s = self.scale
        if s.requires_grad:
            return False
        key = (s._version, s.data_ptr(), str(s.device))
        if getattr(self, "_unit_key", None) != key:
            self._unit_val = bool((s == 1).all())
            self._unit_key = key
        return self._unit_val

    def distribution_mode(self) -> torch.Tensor:
        return self.loc


class BernoulliLayer(RegionGraphLayer):
    leaf_kind = _lib.LEAF_BERNOULLI

    def __init__(self, in_features: int, out_channels: int, regions: List[tuple], rg_depth: int,
                 dropout: Optional[float] = None):
        super().__init__(in_features, out_channels, regions, rg_depth, dropout)
        self.logits = nn.Parameter(
            torch.randn(self.in_regions, self.out_channels, self.dimension), requires_grad=True
        )
        self.distribution = distributions.Bernoulli(logits=self.logits, validate_args=False)

    def leaf_parameters(self):
        return self.logits, None

    def distribution_mode(self) -> torch.Tensor:
        return (self.logits >= 0.0).float()   # sigmoid(l) >= 0.5  <=>  l >= 0


class ProductLayer(nn.Module):
    """Pairs sibling regions (2p, 2p+1): (B, G, K) -> (B, G/2, K*K), output node i*K+j."""

    def __init__(self, in_regions: int, in_nodes: int):
        super().__init__()
        self.in_regions = in_regions
        self.in_nodes = in_nodes
        self.out_partitions = in_regions // 2
        self.out_nodes = in_nodes ** 2
        self.register_buffer('mask', torch.tensor([True, False] * self.out_partitions))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _engine.outer_sum_forward(x, self.out_partitions, self.in_nodes)

    @torch.no_grad()
    def mpe(self, x, idx_group, idx_offset):
        return self.sample(idx_group, idx_offset)

    @torch.no_grad()
    def sample(self, idx_group: torch.Tensor, idx_offset: torch.Tensor):
        left = torch.div(idx_offset, self.in_nodes, rounding_mode='floor')
        right = idx_offset - left * self.in_nodes
        groups = torch.stack([2 * idx_group, 2 * idx_group + 1], dim=2).flatten(1)
        offsets = torch.stack([left, right], dim=2).flatten(1)
        return groups, offsets


class SumLayer(nn.Module):
    """Mixtures per partition: (B, P, K_in) -> (B, P, O); weight (P, O, K_in) holds raw logits."""

    def __init__(self, in_partitions: int, in_nodes: int, out_nodes: int, dropout: Optional[float] = None):
        super().__init__()
        self.in_partitions = in_partitions
        self.in_nodes = in_nodes
        self.out_regions = in_partitions
        self.out_nodes = out_nodes
        self.dropout = dropout
        self.weight = nn.Parameter(torch.empty(self.out_regions, self.out_nodes, self.in_nodes), requires_grad=True)
        dirichlet_(self.weight, alpha=1.0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training and self.dropout is not None:   # layers/ratspn.py:370-372 (the stand-alone layer has no backward)
            x = x.masked_fill(torch.rand_like(x) < self.dropout, float("-inf"))
        return _engine.mixture_forward(x, self.weight)

    @torch.no_grad()
    def mpe(self, x, idx_group, idx_offset):
        rows = torch.arange(x.shape[0], device=x.device).unsqueeze(1)
        w = torch.log_softmax(self.weight[idx_group, idx_offset], dim=2)
        return idx_group, torch.argmax(x[rows, idx_group] + w, dim=2)

    @torch.no_grad()
    def sample(self, idx_group, idx_offset):
        w = torch.log_softmax(self.weight[idx_group, idx_offset], dim=2)
        return idx_group, distributions.Categorical(logits=w).sample()


class RootLayer(nn.Module):
    """Class mixtures over every (partition, node): (B, P, K_in) -> (B, C); weight (C, P*K_in)."""

    def __init__(self, in_partitions: int, in_nodes: int, out_classes: int):
        super().__init__()
        self.in_partitions = in_partitions
        self.in_nodes = in_nodes
        self.out_classes = out_classes
        self.weight = nn.Parameter(torch.empty(self.out_classes, in_partitions * in_nodes), requires_grad=True)
        dirichlet_(self.weight, alpha=1.0)

    def forward(self, x):
        flat = x.flatten(1).unsqueeze(1)                       # (B, 1, P*K_in)
        return _engine.mixture_forward(flat, self.weight.unsqueeze(0)).squeeze(1)

    def _split(self, idx):
        group = torch.div(idx, self.in_nodes, rounding_mode='floor')
        return group, idx - group * self.in_nodes

    @torch.no_grad()
    def mpe(self, x: torch.Tensor, y: torch.Tensor):
        w = torch.log_softmax(self.weight, dim=1)
        return self._split(torch.argmax(x.flatten(1) + w[y], dim=1, keepdim=True))

    @torch.no_grad()
    def sample(self, y: torch.Tensor):
        w = torch.log_softmax(self.weight, dim=1)
        return self._split(distributions.Categorical(logits=w[y]).sample().unsqueeze(1))
