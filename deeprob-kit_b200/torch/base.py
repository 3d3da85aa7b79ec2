"""Model API boundary (interface of deeprob/torch/base.py:11-49)."""
import abc
import copy
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

    # ---- kernel-side caches ----------------------------------------------------------------------
    # The CUDA path keeps parameter-derived tables, operand images and workspaces as plain (non-buffer) attributes
    # of the modules, keyed on the parameters' (data_ptr, _version).  In-place updates through autograd-visible
    # ops (optimisers, load_state_dict, constraints) bump the version and refresh them automatically.  A write
    # through `.data`, or a re-assigned Parameter that lands on the same address, does NOT: call
    # `invalidate_caches()` after such an edit.  The caches are dropped before pickling / deep-copying so that
    # `torch.save(model)` and `copy.deepcopy(model)` carry parameters and buffers only.
    _CACHE_ATTRS = ("_ws_cache", "_ws_sig", "_host_pipeline", "_unit_key", "_unit_val", "_tree_sig")

    def invalidate_caches(self) -> None:
        """Forget every derived table / workspace; the next call rebuilds them from the current parameters."""
        for m in self.modules():
            d = m.__dict__
            for k in [k for k in d if k in self._CACHE_ATTRS or k.startswith("_dpk_")]:
                if k in ("_ws_cache", "_ws_sig"):
                    d[k] = {}
                else:
                    del d[k]

    def __getstate__(self):
        self.invalidate_caches()
        return self.__dict__.copy()

    def __deepcopy__(self, memo):
        self.invalidate_caches()
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new


DensityEstimator = Union[ProbabilisticModel, distributions.Distribution]
