"""CPU: host-side mirror of the reference interface (constructors, state_dict layout, errors)."""
import pytest
import torch

import param_gen as pg
from conftest import load_golden
from deeprob_kit_b200.spn.layers.ratspn import GaussianLayer
from deeprob_kit_b200.spn.models import BernoulliRatSpn, GaussianRatSpn, RatSpn
from deeprob_kit_b200.torch.constraints import ScaleClipper
from deeprob_kit_b200.torch.initializers import dirichlet_


def test_state_dict_layout_matches_reference():
    m = GaussianRatSpn(784, rg_depth=3, rg_repetitions=16, rg_batch=10, rg_sum=10, random_state=42)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes == {
        'layers.0.mask': (128,), 'layers.1.weight': (64, 10, 100), 'layers.2.mask': (64,),
        'layers.3.weight': (32, 10, 100), 'layers.4.mask': (32,), 'base_layer.loc': (128, 10, 98),
        'base_layer.scale': (128, 10, 98), 'base_layer.mask': (128, 98), 'base_layer.inv_mask': (16, 784),
        'root_layer.weight': (1, 1600),
    }
    assert sum(p.numel() for p in m.parameters()) == 348480
    assert not m.base_layer.scale.requires_grad and m.base_layer.loc.requires_grad
    assert m.base_layer.distribution.loc is m.base_layer.loc


@pytest.mark.parametrize("name", sorted(pg.RATSPN_CASES))
def test_gather_table_matches_reference(name):
    cfg = pg.RATSPN_CASES[name]
    cls = GaussianRatSpn if cfg["kind"] == "gaussian" else BernoulliRatSpn
    m = cls(**pg.ratspn_ctor_kwargs(cfg))
    gold = load_golden("ratspn_" + name)
    assert torch.equal(m.base_layer.mask.int(), torch.from_numpy(gold["mask"]))
    assert torch.equal(m.base_layer._mask_i32, torch.from_numpy(gold["mask"]))
    if cfg["in_features"] % (2 ** cfg["rg_depth"]):
        assert 'base_layer.pad_mask' in m.state_dict() and 'base_layer.inv_pad_mask' in m.state_dict()
        lens = m.base_layer._region_len
        assert torch.equal((~m.base_layer.pad_mask[:, 0, :]).sum(1).int(), lens)


def test_constructor_errors():
    kw = dict(rg_depth=2, rg_repetitions=2, rg_batch=2, rg_sum=2)
    with pytest.raises(ValueError):
        RatSpn(8, torch.nn.Linear, **kw)
    with pytest.raises(ValueError):
        GaussianRatSpn(0, **kw)
    with pytest.raises(ValueError):
        GaussianRatSpn(8, out_classes=0, **kw)
    with pytest.raises(ValueError):
        GaussianRatSpn(8, rg_depth=2, rg_repetitions=2, rg_batch=0, rg_sum=2)
    with pytest.raises(ValueError):
        GaussianRatSpn(8, rg_depth=2, rg_repetitions=2, rg_batch=2, rg_sum=0)
    with pytest.raises(ValueError):
        GaussianRatSpn(8, in_dropout=1.0, **kw)
    with pytest.raises(ValueError):
        GaussianRatSpn(8, sum_dropout=0.0, **kw)
    with pytest.raises(ValueError):
        GaussianRatSpn(8, rg_depth=4, rg_repetitions=1)
    with pytest.raises(ValueError):
        GaussianRatSpn(8, rg_depth=2, rg_repetitions=0)


def test_no_cpu_fallback():
    m = BernoulliRatSpn(16, rg_depth=2, rg_repetitions=1, rg_batch=2, rg_sum=2, random_state=42).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(4, 16))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.log_prob(torch.zeros(4, 16))


def test_loss_and_constraints():
    m = GaussianRatSpn(8, rg_depth=2, rg_repetitions=2, rg_batch=2, rg_sum=2, optimize_scale=True, random_state=1)
    assert m.base_layer.scale.requires_grad
    with torch.no_grad():
        m.base_layer.scale.fill_(-1.0)
    m.apply_constraints()
    assert torch.all(m.base_layer.scale > 0) and float(m.base_layer.scale.max()) < 1.1e-5
    ll = torch.tensor([[-1.0], [-3.0]])
    assert float(m.loss(ll)) == 2.0
    mc = GaussianRatSpn(8, out_classes=3, rg_depth=2, rg_repetitions=2, rg_batch=2, rg_sum=2, random_state=1)
    out = torch.randn(5, 3)
    y = torch.tensor([0, 1, 2, 1, 0])
    assert torch.allclose(mc.loss(out, y), torch.nn.functional.cross_entropy(out, y))


def test_initialisers():
    # mirrors deeprob-kit tests/test_torch.py:29-42
    t = torch.empty(4, 5, 6)
    dirichlet_(t, log_space=False)
    assert torch.allclose(t.sum(-1), torch.ones(4, 5))
    dirichlet_(t, log_space=True)
    assert torch.allclose(t.exp().sum(-1), torch.ones(4, 5))
    dirichlet_(t, log_space=False, dim=1)
    assert torch.allclose(t.sum(1), torch.ones(4, 6))
    with pytest.raises(ValueError):
        dirichlet_(torch.tensor(0.0))
    with pytest.raises(IndexError):
        dirichlet_(t, dim=3)
    with pytest.raises(ValueError):
        ScaleClipper(eps=0.0)
    layer = GaussianLayer(8, 3, [(0, 1, 2, 3), (4, 5, 6, 7)], rg_depth=1, uniform_loc=(-2.0, 2.0))
    assert float(layer.loc.min()) >= -2.0 and float(layer.loc.max()) <= 2.0
