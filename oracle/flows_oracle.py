"""CPU oracle for the flow bijectors -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see ratspn_oracle.py).

Functional torch-CPU restatement of the density direction (`apply_backward`) of the reference flows, driven
by a reference-keyed state_dict.  The conv conditioners of RealNVP2d are library code on both sides and are
not restated: 2D models are pinned directly by golden vectors of the reference (tests/golden/flows_*.npz).
"""
import math
from typing import Dict

import torch
import torch.nn.functional as F

_LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))

# When set to a list, `mlp` appends the per-sample distance of the hidden pre-activations from the ReLU kink
# (min_j |h_j|) of every hidden layer it evaluates.  Gradients of a ReLU network are discontinuous there, so a
# gradient comparison between two correctly rounded evaluations is only meaningful on samples that keep a margin.
KINK_MARGINS = None


def mlp(x, state, prefix, masked=False, activation=torch.relu):
    """nn.Sequential of Linear(+mask)/activation pairs ending in a Linear (coupling.py:45-56, autoregressive.py:58-70)."""
    idx = sorted({int(k[len(prefix):].split(".")[0]) for k in state if k.startswith(prefix) and k.endswith(".weight")})
    h = x
    for n, i in enumerate(idx):
        w = state["%s%d.weight" % (prefix, i)]
        if masked:
            w = state["%s%d.mask" % (prefix, i)] * w                 # MaskedLinear (torch/utils.py:96)
        h = F.linear(h, w, state["%s%d.bias" % (prefix, i)])
        if n != len(idx) - 1:
            if KINK_MARGINS is not None and activation is torch.relu:
                KINK_MARGINS.append(h.detach().abs().min(dim=1).values)
            h = activation(h)
    return h


def coupling1d_backward(x, state, p, affine=True):
    """CouplingLayer1d.apply_backward (flows/layers/coupling.py:72-87)."""
    mask, inv_mask = state[p + "mask"], state[p + "inv_mask"]
    z = mlp(mask * x, state, p + "network.")
    if not affine:
        return x - inv_mask * z, 0.0
    t, s = torch.chunk(z, 2, dim=1)
    s = state[p + "scale_act.weight"] * torch.tanh(s)               # ScaledTanh (torch/utils.py:70)
    t, s = inv_mask * t, inv_mask * s
    return (x - t) * torch.exp(-s), -s.sum(1)


def made_backward(x, state, p, activation):
    """AutoregressiveLayer.apply_backward (flows/layers/autoregressive.py:72-79)."""
    z = mlp(x, state, p + "network.", masked=True, activation=activation)
    t, s = torch.chunk(z, 2, dim=1)
    s = state[p + "scale_act.weight"] * torch.tanh(s)
    return (x - t) * torch.exp(-s), -s.sum(1)


def batchnorm_backward(x, state, p, training=False, eps=1e-5):
    """BatchNormLayer1d/2d.apply_backward (flows/utils.py:118-139, 183-206); returns the batch statistics too."""
    w, b = state[p + "weight"], state[p + "bias"]
    if training:
        if x.dim() == 2:
            var, mean = torch.var_mean(x, dim=0, keepdim=True)
        else:
            mean = x.mean(dim=[0, 2, 3], keepdim=True)
            var = ((x - mean) ** 2.0).mean(dim=[0, 2, 3], keepdim=True)
    else:
        mean, var = state[p + "running_mean"], state[p + "running_var"]
    v = var + eps
    u = (x - mean) / torch.sqrt(v) * torch.exp(w) + b
    grid = 1 if x.dim() == 2 else x.shape[2] * x.shape[3]
    return u, (torch.sum(w - 0.5 * torch.log(v)) * grid).expand(x.shape[0]), (mean, var)


def logit_backward(x, alpha, ldj_const):
    """LogitLayer.apply_backward (flows/utils.py:276-284)."""
    y = alpha + (1.0 - 2.0 * alpha) * x
    lx, rx = torch.log(y), torch.log(1.0 - y)
    return lx - rx, -((lx + rx).flatten(1).sum(1) + ldj_const)


def dequantize_backward(x, noise, bins, ldj_const):
    """DequantizeLayer.apply_backward with the uniform noise made explicit (flows/utils.py:244-248)."""
    return (x * (bins - 1) + noise) / bins, -ldj_const.expand(x.shape[0])


def normal_prior(z, loc, scale):
    """in_base.log_prob summed per sample (flows/models/base.py:139-140)."""
    ll = -((z - loc) ** 2) / (2 * scale ** 2) - scale.log() - _LOG_SQRT_2PI
    return ll.flatten(1).sum(1)


def flow1d_log_prob(x, state: Dict[str, torch.Tensor], kind: str, kw: dict, training=False):
    """RealNVP1d / MAF forward = log-likelihood (flows/models/base.py:123-143)."""
    ildj = torch.zeros(x.shape[0], dtype=x.dtype)
    if kw.get("logit") is not None:
        x, part = logit_backward(x, kw["logit"], state["logit.ldj"])
        ildj = ildj + part
    act = {"relu": torch.relu, "tanh": torch.tanh}[kw.get("activation", "relu")]
    n_layers = len({k.split(".")[1] for k in state if k.startswith("layers.")})
    stats = []
    for i in range(n_layers):
        p = "layers.%d." % i
        if p + "running_var" in state:
            x, part, st = batchnorm_backward(x, state, p, training)
            stats.append(st)
        elif kind == "MAF":
            x, part = made_backward(x, state, p, act)
        else:
            x, part = coupling1d_backward(x, state, p, kw.get("affine", True))
        ildj = ildj + part
    return normal_prior(x, state["in_base_loc"], state["in_base_scale"]) + ildj, stats
