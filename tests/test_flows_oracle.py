"""CPU: flow oracle pinned against reference vectors; host-side mirror of the flow constructors."""
import numpy as np
import pytest
import torch

import param_gen as pg
from conftest import load_golden, norm_err, rel_err
from helpers import flow_reference_state
from oracle.flows_oracle import flow1d_log_prob
from deeprob_kit_b200.flows.models import MAF, RealNVP1d, RealNVP2d
from deeprob_kit_b200.flows.utils import (BatchNormLayer1d, DequantizeLayer, LogitLayer, squeeze_depth2d,
                                          unsqueeze_depth2d)
from deeprob_kit_b200.torch.utils import MaskedLinear, ScaledTanh

ONE_D = [n for n, c in pg.FLOW_CASES.items() if c["model"] != "RealNVP2d"]


@pytest.mark.parametrize("name", ONE_D)
def test_flow1d_oracle_matches_reference_golden(name):
    cfg = pg.FLOW_CASES[name]
    gold = load_golden("flows_" + name)
    _, state = flow_reference_state(cfg)
    x, g = pg.flow_inputs(cfg)
    with torch.enable_grad():
        xx = x.clone().requires_grad_(True)
        ll, _ = flow1d_log_prob(xx, state, cfg["model"], cfg["kw"])
        (ll * g).sum().backward()
    assert rel_err(ll.detach(), gold["ll"]) < 2e-5
    assert norm_err(xx.grad[:4], gold["grad.x"]) < 1e-4
    if "train.ll" in gold:
        ll_t, stats = flow1d_log_prob(x, state, cfg["model"], cfg["kw"], training=True)
        assert rel_err(ll_t, gold["train.ll"]) < 2e-5


def test_flow_constructors_and_errors():
    # deeprob-kit tests/test_flows.py:65-131 (constructor ValueErrors)
    for bad in (dict(n_flows=0), dict(depth=0), dict(units=0)):
        with pytest.raises(ValueError):
            RealNVP1d(10, **bad)
        with pytest.raises(ValueError):
            MAF(10, **bad)
    for bad in (dict(n_flows=0), dict(n_blocks=0), dict(channels=0)):
        with pytest.raises(ValueError):
            RealNVP2d((3, 8, 8), **bad)
    with pytest.raises(ValueError):
        RealNVP1d(10, logit=1.5)
    with pytest.raises(ValueError):
        RealNVP1d((3, 8), n_flows=1)
    with pytest.raises(NotImplementedError):
        RealNVP2d((3, 8, 8), network="unknown")
    with pytest.raises(ValueError):
        BatchNormLayer1d(4, momentum=1.0)
    with pytest.raises(ValueError):
        DequantizeLayer(4, n_bits=0)
    with pytest.raises(ValueError):
        LogitLayer(4, alpha=0.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        RealNVP1d(10, n_flows=1)(torch.zeros(3, 10))


def test_helpers_host():
    # deeprob-kit tests/test_torch.py:136-155
    x = torch.randn(5, 4)
    act = ScaledTanh()
    assert torch.all(act(x) == 0.0)
    act.weight.data.fill_(2.0)
    assert torch.allclose(act(x), 2.0 * torch.tanh(x))
    mask = np.tril(np.ones((3, 4)))
    lin = MaskedLinear(4, 3, mask)
    assert torch.allclose(lin(x), torch.nn.functional.linear(x, lin.weight * torch.tensor(mask, dtype=torch.float32), lin.bias))
    with pytest.raises(ValueError):
        MaskedLinear(4, 3, np.ones((2, 2)))
    img = torch.randn(2, 3, 8, 8)
    assert torch.equal(unsqueeze_depth2d(squeeze_depth2d(img)), img)
    assert squeeze_depth2d(img).shape == (2, 12, 4, 4)
    perm = RealNVP2d.build_permutation_matrix(3)
    assert perm.shape == (12, 3, 2, 2) and float(perm.sum()) == 12.0
    # the index form of the multi-scale permutation equals the reference's conv / conv_transpose with those weights
    down = torch.nn.functional.conv2d(img, perm, stride=2)
    assert torch.equal(RealNVP2d.downscale(img), down)
    assert torch.equal(RealNVP2d.upscale(down), torch.nn.functional.conv_transpose2d(down, perm, stride=2))
    assert torch.equal(RealNVP2d.upscale(down), img)
