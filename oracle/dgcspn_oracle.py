"""CPU oracle for the DGC-SPN path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see ratspn_oracle.py).

Functional torch-CPU restatement of deeprob/spn/layers/dgcspn.py and deeprob/spn/models/dgcspn.py,
pinned against the live reference by tests/golden/dgcspn_*.npz.
"""
import math
from itertools import product as iter_product
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

_LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


def spatial_gaussian(x, loc, scale, clean_nan=False, drop=None):
    """SpatialGaussianLayer.forward (layers/dgcspn.py:101-120): (B,Cin,H,W)->(B,K,H,W).  `drop` (bool,
    (B,K,Cin,H,W)) restates the training-mode dropout of :113-115 (NaN on the per-channel log-densities) with
    injected Bernoulli draws."""
    v = x.unsqueeze(1)
    missing = torch.isnan(v)
    if clean_nan:           # same values; autograd then gives the gradient of the marginalised LL (see ratspn_oracle)
        v = torch.where(missing, torch.zeros_like(v), v)
    ll = -((v - loc) ** 2) / (2 * scale ** 2) - scale.log() - _LOG_SQRT_2PI   # Normal.log_prob  (:110)
    if drop is not None:
        ll = torch.where(drop, torch.full_like(ll, float("nan")), ll)          # :113-115
    ll = torch.nan_to_num(ll)                                                  # :117
    if clean_nan:
        ll = torch.where(missing, torch.zeros_like(ll), ll)
    return ll.sum(2)                                                           # :120


def product_geometry(in_features, padding, stride, dilation, depthwise):
    """Pad list, output size and conv weight of SpatialProductLayer (layers/dgcspn.py:160-198), 2x2 kernels."""
    c, h, w = in_features
    keh, kew = dilation[0] + 1, dilation[1] + 1
    if padding == "valid":
        pad = [0, 0, 0, 0]
    elif padding == "full":
        pad = [kew - 1, kew - 1, keh - 1, keh - 1]
    else:  # final
        pad = [0, (kew - 1) * 2 - w, 0, (keh - 1) * 2 - h]
    oh = int(np.ceil((pad[2] + pad[3] + h - keh + 1) / stride[0]))
    ow = int(np.ceil((pad[0] + pad[1] + w - kew + 1) / stride[1]))
    if depthwise:
        weight, oc = torch.ones(c, 1, 2, 2), c
    else:
        oc = c ** 4
        ids = np.array(list(iter_product(range(c), repeat=4))).reshape(oc, 1, 2, 2)
        weight = torch.tensor(np.arange(c).reshape(1, c, 1, 1) == ids, dtype=torch.float32)
    return pad, (oc, oh, ow), weight


def spatial_product(x, pad, weight, stride, dilation, groups):
    """SpatialProductLayer.forward (layers/dgcspn.py:232-236)."""
    return F.conv2d(F.pad(x, pad), weight.to(x.dtype), stride=stride, dilation=dilation, groups=groups)


def spatial_sum(x, weight, drop=None):
    """SpatialSumLayer.forward (layers/dgcspn.py:289-303); `drop` (bool, like x) = the -inf dropout of :297-299."""
    if drop is not None:
        x = torch.where(drop, torch.full_like(x, float("-inf")), x)
    return torch.logsumexp(x.unsqueeze(1) + torch.log_softmax(weight, dim=1), dim=2)


def spatial_root(x, weight):
    """SpatialRootLayer.forward (layers/dgcspn.py:351-354)."""
    return torch.logsumexp(x.flatten(1).unsqueeze(1) + torch.log_softmax(weight, dim=1), dim=2)


class DgcSpnOracle:
    """Layer schedule of DgcSpn.__init__ (models/dgcspn.py:66-128) + forward (:134-151)."""

    def __init__(self, in_features, out_classes=1, n_batch=8, sum_channels=8, depthwise=False, n_pooling=0):
        depth = int(np.ceil(np.log2(in_features[1])))
        if isinstance(depthwise, bool):
            depthwise = [depthwise] * (depth + 1)
        else:
            depthwise = list(depthwise) + [depthwise[-1]] * (depth + 1 - len(depthwise))
        self.in_features, self.out_classes = tuple(in_features), out_classes
        self.n_batch, self.sum_channels = n_batch, sum_channels
        self.products: List[dict] = []
        shape = (n_batch, in_features[1], in_features[2])
        self.sum_shapes = []
        for i in range(depth + 1):
            if i < n_pooling:
                padding, stride, dil = "valid", (2, 2), (1, 1)
            else:
                padding, stride = ("final" if i == depth else "full"), (1, 1)
                dil = (2 ** (i - n_pooling),) * 2
            pad, out, weight = product_geometry(shape, padding, stride, dil, depthwise[i])
            self.products.append(dict(pad=pad, weight=weight, stride=stride, dilation=dil,
                                      groups=shape[0] if depthwise[i] else 1, in_shape=shape, out_shape=out))
            shape = out
            if i != depth:
                self.sum_shapes.append((sum_channels, *shape))
                shape = (sum_channels, shape[1], shape[2])
        self.root_in = shape
        self.loc = self.scale = self.root_weight = None
        self.sum_weights: List[torch.Tensor] = []
        self.clean_nan = False

    def load_reference_state(self, state: Dict[str, torch.Tensor]):
        self.loc, self.scale = state["base_layer.loc"].float(), state["base_layer.scale"].float()
        self.sum_weights = [state["layers.%d.weight" % (2 * i + 1)].float() for i in range(len(self.sum_shapes))]
        self.root_weight = state["root_layer.weight"].float()
        return self

    def double(self):
        self.loc, self.scale, self.root_weight = self.loc.double(), self.scale.double(), self.root_weight.double()
        self.sum_weights = [w.double() for w in self.sum_weights]
        return self

    leaf_drop = None    # injected training-mode dropout draws (see spatial_gaussian / spatial_sum)
    sum_drops = None

    def log_prob(self, x, keep=None):
        h = spatial_gaussian(x, self.loc, self.scale, self.clean_nan, self.leaf_drop)
        if keep is not None:
            keep.append(h)
        for i, p in enumerate(self.products):
            h = spatial_product(h, p["pad"], p["weight"], p["stride"], p["dilation"], p["groups"])
            if i < len(self.sum_weights):
                h = spatial_sum(h, self.sum_weights[i], self.sum_drops[i] if self.sum_drops is not None else None)
        return spatial_root(h, self.root_weight)

    def grads(self, x, grad_out, clean_nan=True):
        params = [self.loc, self.scale, *self.sum_weights, self.root_weight]
        leaves = [p.clone().requires_grad_(True) for p in params]
        saved = (self.loc, self.scale, self.sum_weights, self.root_weight, self.clean_nan)
        self.loc, self.scale, self.sum_weights, self.root_weight = leaves[0], leaves[1], leaves[2:-1], leaves[-1]
        self.clean_nan = clean_nan
        xx = x.clone().requires_grad_(True)
        try:
            with torch.enable_grad():
                out = self.log_prob(xx)
                (out * grad_out).sum().backward()
        finally:
            self.loc, self.scale, self.sum_weights, self.root_weight, self.clean_nan = saved
        return {"out": out.detach(), "x": xx.grad, "loc": leaves[0].grad, "scale": leaves[1].grad,
                "sums": [t.grad for t in leaves[2:-1]], "root": leaves[-1].grad}
