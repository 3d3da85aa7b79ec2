"""CPU: DGC-SPN oracle pinned against reference vectors; host-side mirror of the reference interface."""
import numpy as np
import pytest
import torch

import param_gen as pg
from conftest import load_golden, norm_err, rel_err
from helpers import dgc_oracle_for, subsample_like
from deeprob_kit_b200.spn.layers.dgcspn import (SpatialGaussianLayer, SpatialProductLayer, SpatialRootLayer,
                                                SpatialSumLayer)
from deeprob_kit_b200.spn.models import DgcSpn


@pytest.mark.parametrize("name", sorted(pg.DGCSPN_CASES))
def test_dgcspn_oracle_matches_reference_golden(name):
    cfg = pg.DGCSPN_CASES[name]
    gold = load_golden("dgcspn_" + name)
    orc, _ = dgc_oracle_for(cfg)
    x, g = pg.dgcspn_inputs(cfg)
    res = orc.grads(x, g, clean_nan=False)
    assert rel_err(res["out"], gold["ll"]) < 1e-5
    sum_keys = sorted((k for k in gold if k.startswith("grad.layers.")), key=lambda k: int(k.split(".")[2]))
    named = {"grad.base_layer.loc": res["loc"], "grad.base_layer.scale": res["scale"], "grad.root_layer.weight": res["root"]}
    for k, t in zip(sum_keys, res["sums"]):
        named[k] = t
    for k, mine in named.items():
        assert norm_err(subsample_like(mine, gold[k].size), gold[k].reshape(-1)) < 2e-4, k
    assert norm_err(torch.nan_to_num(res["x"])[:4], gold["grad.x"]) < 2e-4


# ---- host-side mirror: the reference's own layer tests (deeprob-kit tests/test_dgcspn.py:26-86) -------
def test_gaussian_layer_host():
    layer = SpatialGaussianLayer((3, 32, 32), out_channels=16, optimize_scale=False, uniform_loc=(-1.0, 1.0))
    assert layer.out_features == (16, 32, 32)
    assert torch.all(layer.loc >= -1.0) and torch.all(layer.loc <= 1.0)
    assert torch.all(layer.scale == 1.0) and not layer.scale.requires_grad
    layer = SpatialGaussianLayer((3, 32, 32), out_channels=16, optimize_scale=True)
    assert torch.all(layer.scale > 0.0) and layer.scale.requires_grad
    with pytest.raises(ValueError):
        SpatialGaussianLayer((3, 32, 32), 16, quantiles_loc=np.zeros([16, 3, 32, 32]), uniform_loc=(-1.0, 1.0))


def test_product_layer_host():
    p = SpatialProductLayer((3, 32, 32), kernel_size=2, padding='full', stride=1, dilation=4, depthwise=True)
    assert p.pad == [4, 4, 4, 4] and p.out_features == (3, 36, 36)
    assert torch.all(p.weight == torch.ones(3, 1, 2, 2))
    p = SpatialProductLayer((3, 32, 32), kernel_size=2, padding='valid', stride=2, dilation=1, depthwise=True)
    assert p.pad == [0, 0, 0, 0] and p.out_features == (3, 16, 16)
    p = SpatialProductLayer((3, 32, 32), kernel_size=2, padding='full', stride=1, dilation=8, depthwise=False)
    assert p.pad == [8, 8, 8, 8] and tuple(p.weight.shape) == (81, 3, 2, 2)
    assert torch.allclose(p.weight.sum(1), torch.tensor(1.0))
    assert p.out_features == (81, 40, 40)
    with pytest.raises(ValueError):
        SpatialProductLayer((3, 32, 32), kernel_size=2, padding='same', stride=1, dilation=1)


def test_sum_root_and_model_host():
    s = SpatialSumLayer((3, 32, 32), out_channels=8)
    assert s.out_features == (8, 32, 32) and tuple(s.weight.shape) == (8, 3, 32, 32)
    assert torch.allclose(s.weight.exp().sum(1), torch.ones(8, 32, 32), atol=1e-5)
    r = SpatialRootLayer((3, 32, 32), out_channels=8)
    assert r.out_channels == 8 and tuple(r.weight.shape) == (8, 3 * 32 * 32)
    m = DgcSpn((1, 28, 28), n_batch=8, sum_channels=8, depthwise=True)
    assert [tuple(l.out_features) for l in m.layers][-1] == (8, 32, 32)
    assert tuple(m.root_layer.weight.shape) == (1, 8192)
    for bad in (dict(in_features=(1, 8, 9)), dict(in_features=(1, 8, 8), out_classes=0),
                dict(in_features=(1, 8, 8), n_batch=0), dict(in_features=(1, 8, 8), sum_channels=0),
                dict(in_features=(1, 8, 8), in_dropout=1.5), dict(in_features=(1, 8, 8), n_pooling=9),
                dict(in_features=(1, 8, 8), depthwise=[]), dict(in_features=(1, 8, 8), uniform_loc=(1.0, 0.0))):
        with pytest.raises(ValueError):
            DgcSpn(**bad)
    with pytest.raises(NotImplementedError):
        m.sample(3)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(2, 1, 28, 28))
