"""RAT-SPN models (interface of deeprob/spn/models/ratspn.py: RatSpn :16-191, GaussianRatSpn
:194-239, BernoulliRatSpn :242-273).

Same constructor signatures, ValueErrors, attribute names and state_dict keys
(`base_layer.*`, `layers.{i}.mask|weight`, `root_layer.weight`), so checkpoints of the reference load
unchanged.  `forward` (= `log_prob`) is ONE autograd node over the fused CUDA path
(csrc/ratspn_fwd.cu, csrc/ratspn_bwd.cu) instead of a chain of per-layer tensor ops.
"""
from typing import Optional, Tuple, Type

import os

import torch
import torch.nn.functional as F

from ...torch.base import ProbabilisticModel
from ...torch.constraints import ScaleClipper
from ...utils.random import RandomState
from ...utils.region import RegionGraph
from .. import _engine
from ..layers.ratspn import (BernoulliLayer, GaussianLayer, ProductLayer, RegionGraphLayer, RootLayer,
                             SumLayer)


class RatSpn(ProbabilisticModel):
    def __init__(
        self,
        in_features: int,
        base_cls: Type[RegionGraphLayer],
        base_kwargs: Optional[dict] = None,
        out_classes: int = 1,
        rg_depth: int = 2,
        rg_repetitions: int = 1,
        rg_batch: int = 2,
        rg_sum: int = 2,
        in_dropout: Optional[float] = None,
        sum_dropout: Optional[float] = None,
        random_state: Optional[RandomState] = None
    ):
        if not issubclass(base_cls, RegionGraphLayer):
            raise ValueError("The base distribution's class must be a sub-class of RegionGraphLayer")
        if in_features <= 0:
            raise ValueError("The number of input features must be positve")
        if out_classes <= 0:
            raise ValueError("The number of output classes must be positive")
        if rg_batch <= 0:
            raise ValueError("The number of base distribution batches must be positive")
        if rg_sum <= 0:
            raise ValueError("The number of sum nodes per region must be positive")
        if in_dropout is not None and not 0.0 < in_dropout < 1.0:
            raise ValueError("The dropout rate at base distribution must be in (0, 1)")
        if sum_dropout is not None and not 0.0 < sum_dropout < 1.0:
            raise ValueError("The dropout rate at sum layers must be in (0, 1)")

        super().__init__()
        self.in_features = in_features
        self.out_classes = out_classes
        self.rg_depth = rg_depth
        self.rg_repetitions = rg_repetitions
        self.rg_batch = rg_batch
        self.rg_sum = rg_sum
        self.in_dropout = in_dropout
        self.sum_dropout = sum_dropout
        self.layers = torch.nn.ModuleList()

        # Structure: index 0 = leaf regions, last = root region
        graph = RegionGraph(in_features, rg_depth, random_state)
        self.rg_layers = graph.make_layers(rg_repetitions)[::-1]

        self.base_layer = base_cls(
            in_features, rg_batch, regions=self.rg_layers[0], rg_depth=rg_depth, dropout=in_dropout,
            **(base_kwargs or {})
        )

        # Product, Sum, Product, ..., Product (2*depth - 1 inner layers), then the root
        groups, nodes = self.base_layer.in_regions, self.base_layer.out_channels
        for level in range(1, len(self.rg_layers) - 1):
            if level % 2 == 1:
                layer = ProductLayer(groups, nodes)
                groups, nodes = layer.out_partitions, layer.out_nodes
            else:
                layer = SumLayer(groups, nodes, rg_sum, sum_dropout)
                groups, nodes = layer.out_regions, layer.out_nodes
            self.layers.append(layer)
        self.root_layer = RootLayer(groups, nodes, out_classes)
        self._ws_cache = {}
        self._ws_sig = {}
        self.cache_tables = True

    # ---- kernel plumbing ---------------------------------------------------------------------
    def _sum_layers(self):
        return [layer for layer in self.layers if isinstance(layer, SumLayer)]

    def _kernel_parameters(self) -> Tuple[torch.Tensor, ...]:
        p0, p1 = self.base_layer.leaf_parameters()
        leaf = (p0,) if p1 is None else (p0, p1)
        return (*leaf, *(layer.weight for layer in self._sum_layers()), self.root_layer.weight)

    def _make_call(self, device, force_scale: bool = False) -> "_engine.RatSpnCall":
        return _engine.RatSpnCall(self.base_layer, [layer.weight for layer in self._sum_layers()],
                                  self.root_layer.weight, self.out_classes, self.rg_sum, self.rg_repetitions, device,
                                  force_scale)

    def _workspace(self, call, batch: int, flags: int, device, private: bool):
        """(workspace, extra flags).  Inference reuses one grow-only buffer per (device, stream); when neither the
        buffer nor the parameters (data pointer + version counter, like `unit_scale`) nor the batch size changed
        since the previous call, the parameter-derived tables inside it are still valid and the kernels that
        rebuild them are skipped (DPK_F_TABLES_VALID).  In-place updates through autograd-visible ops (optimizers,
        `load_state_dict`, the constraint modules) bump the version counter; writes through `.data` do not --
        set `model.cache_tables = False` (or DPK_TABLE_CACHE=0) if you do that."""
        if private:  # activations must survive until backward: one buffer per autograd node
            return call.workspace(batch, flags, device), 0
        # inference: grow-only buffer per (device, stream), reused call after call (stream-ordered)
        key = (str(device), flags, torch.cuda.current_stream(device).cuda_stream)
        ws = self._ws_cache.get(key)
        nbytes = call.workspace_bytes(batch, flags)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._ws_cache[key] = ws
            self._ws_sig.pop(key, None)
        sig = (ws.data_ptr(), batch, call.n_leaf, call.keep[1] is None, os.environ.get("DPK_LEAF_MMA"), os.environ.get("DPK_TREE_MMA"),
               os.environ.get("DPK_LEAF_STREAM"),
               tuple((t.data_ptr(), t._version) for t in self._kernel_parameters()))
        valid = self.cache_tables and os.environ.get("DPK_TABLE_CACHE", "1") != "0" and self._ws_sig.get(key) == sig
        self._ws_sig[key] = sig
        return ws, (_engine._lib.F_TABLES_VALID if valid else 0)

    # ---- API ---------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Log-likelihood (B, out_classes) given evidence x (B, in_features); NaN = marginalised.  In training mode
        with `in_dropout` / `sum_dropout` set, the probabilistic dropout of the reference layers
        (layers/ratspn.py:98-100, :370-372) is applied with draws from a counter-based generator (csrc/ratspn_dropout.cu)."""
        return _engine.ratspn_log_prob(self, x)

    def em_statistics(self, x: torch.Tensor):
        """E-step sufficient statistics of a batch (extension)."""
        return _engine.ratspn_em_statistics(self, x)

    @torch.no_grad()
    def mpe(self, x: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Maximum-a-posteriori completion of the NaN entries of x (models/ratspn.py:124-160): one forward that keeps
        every level's log-likelihoods, then one top-down kernel (csrc/ratspn_topdown.cu)."""
        if self.out_classes == 1:
            y = None
        return _engine.ratspn_mpe(self, x, y)

    @torch.no_grad()
    def sample(self, n_samples: int, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Ancestral samples (models/ratspn.py:162-182), one kernel (csrc/ratspn_topdown.cu)."""
        device = self.root_layer.weight.device
        if self.out_classes == 1:
            y = None
        elif y is None:
            y = torch.randint(self.out_classes, [n_samples])
        return _engine.ratspn_sample(self, n_samples, y, device)

    def mpe_layerwise(self, x: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The reference's layer-by-layer index walk (same result as `mpe`; kept for the stand-alone layer API)."""
        with torch.no_grad():
            inputs, n = x, x.shape[0]
            lls = []
            h = self.base_layer(x)
            for layer in self.layers:
                lls.append(h)
                h = layer(h)
            if self.out_classes == 1:
                y = torch.zeros(n, dtype=torch.long, device=x.device)
            elif y is None:
                y = torch.argmax(self.root_layer(h), dim=1)
            group, offset = self.root_layer.mpe(h, y)
            for i in range(len(self.layers) - 1, -1, -1):
                group, offset = self.layers[i].mpe(lls[i], group, offset)
            return self.base_layer.mpe(inputs, group, offset)

    def loss(self, x: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        """models/ratspn.py:184-191 as one kernel (value + gradient w.r.t. the log-likelihoods, no host sync)."""
        return _engine.nll_loss(x, None if self.out_classes == 1 else y)


class GaussianRatSpn(RatSpn):
    def __init__(
        self,
        in_features: int,
        out_classes: int = 1,
        rg_depth: int = 2,
        rg_repetitions: int = 1,
        rg_batch: int = 2,
        rg_sum: int = 2,
        in_dropout: Optional[float] = None,
        sum_dropout: Optional[float] = None,
        random_state: Optional[RandomState] = None,
        uniform_loc: Optional[Tuple[float, float]] = None,
        optimize_scale: bool = False
    ):
        super().__init__(
            in_features, GaussianLayer, {'uniform_loc': uniform_loc, 'optimize_scale': optimize_scale},
            out_classes, rg_depth, rg_repetitions, rg_batch, rg_sum, in_dropout, sum_dropout, random_state
        )
        self.optimize_scale = optimize_scale
        if optimize_scale:
            self.scale_clipper = ScaleClipper()

    def apply_constraints(self):
        if self.optimize_scale:
            self.scale_clipper(self.base_layer)


class BernoulliRatSpn(RatSpn):
    def __init__(
        self,
        in_features: int,
        out_classes: int = 1,
        rg_depth: int = 2,
        rg_repetitions: int = 1,
        rg_batch: int = 2,
        rg_sum: int = 2,
        in_dropout: Optional[float] = None,
        sum_dropout: Optional[float] = None,
        random_state: Optional[RandomState] = None
    ):
        super().__init__(
            in_features, BernoulliLayer, None,
            out_classes, rg_depth, rg_repetitions, rg_batch, rg_sum, in_dropout, sum_dropout, random_state
        )
