"""Host-side logic of the flow fast paths (no GPU): cached derived tensors, the eval batch-norm affine, the plan that
decides whether a conditioner may run on the tensor-core GEMM, layer pairing, and the loud failure on CPU tensors."""
import pytest
import torch
from torch import nn

from deeprob_kit_b200.flows import _engine
from deeprob_kit_b200.flows.layers.coupling import CouplingLayer1d
from deeprob_kit_b200.flows.models import MAF, RealNVP1d
from deeprob_kit_b200.flows.utils import BatchNormLayer1d
from deeprob_kit_b200.torch.utils import MaskedLinear


def test_derived_tensor_is_rebuilt_in_place_when_a_source_changes():
    w = nn.Parameter(torch.randn(6, 4))
    mask = torch.tensor([1.0, 0.0, 1.0, 0.0])
    cache = {}
    calls = []

    def make():
        calls.append(1)
        return w.detach() * mask

    d0 = _engine._derived(cache, "mw", (w, mask), make)
    v0 = d0._version
    assert _engine._derived(cache, "mw", (w, mask), make) is d0 and len(calls) == 1 and d0._version == v0
    with torch.no_grad():
        w.mul_(2.0)                                    # optimiser-style in-place update
    d1 = _engine._derived(cache, "mw", (w, mask), make)
    assert d1 is d0 and len(calls) == 2                # same storage: pointer-keyed consumers see a version bump
    assert d1._version > v0
    assert torch.equal(d1, w.detach() * mask)


def test_eval_batch_norm_affine_matches_the_reference_formula_and_is_cached():
    torch.manual_seed(0)
    bn = BatchNormLayer1d(5).eval()
    with torch.no_grad():
        bn.weight.normal_(0.0, 0.3)
        bn.bias.normal_()
        bn.running_mean.normal_()
        bn.running_var.uniform_(0.5, 2.0)
    x = torch.randn(7, 5)
    a, c, ldj = _engine.eval_batch_norm_affine(bn)
    # deeprob/flows/utils.py:118-139 in eval mode
    ref_u = (x - bn.running_mean) / torch.sqrt(bn.running_var + bn.eps) * torch.exp(bn.weight) + bn.bias
    ref_ldj = torch.sum(bn.weight - 0.5 * torch.log(bn.running_var + bn.eps))
    assert torch.allclose(x * a + c, ref_u, atol=1e-6)
    assert abs(ldj - float(ref_ldj)) < 1e-6
    assert _engine.eval_batch_norm_affine(bn)[0] is a             # cached
    with torch.no_grad():
        bn.running_mean.add_(1.0)                                 # a training step moved the statistics
    a2, c2, _ = _engine.eval_batch_norm_affine(bn)
    ref_u2 = (x - bn.running_mean) / torch.sqrt(bn.running_var + bn.eps) * torch.exp(bn.weight) + bn.bias
    assert torch.allclose(x * a2 + c2, ref_u2, atol=1e-6)


def test_mlp_plan_only_applies_to_large_cuda_inference_batches():
    net = nn.Sequential(nn.Linear(8, 16), nn.ReLU(inplace=True), nn.Linear(16, 16), nn.ReLU(inplace=True), nn.Linear(16, 16))
    x = torch.randn(4096, 8)
    with torch.no_grad():
        assert _engine._mlp_plan(net, x) is None                  # CPU tensor: stock modules
    assert _engine._mlp_plan(net, x) is None                      # gradients may be needed: stock modules


def test_layer_weight_applies_the_made_mask_once():
    import numpy as np
    m = MaskedLinear(4, 3, np.tril(np.ones((3, 4))))
    cache = {}
    w = _engine._layer_weight(m, cache)
    assert torch.equal(w, m.mask * m.weight.detach())
    assert _engine._layer_weight(m, cache) is w
    plain = nn.Linear(4, 3)
    assert _engine._layer_weight(plain, {}) is plain.weight


def test_flow_structure_pairs_couplings_with_batch_norm():
    model = RealNVP1d(8, n_flows=3, depth=1, units=16, batch_norm=True)
    kinds = [type(l) for l in model.layers]
    assert kinds == [CouplingLayer1d, BatchNormLayer1d] * 3
    # alternating masks: what one coupling transforms is what the next one conditions on (the chained side output)
    c0, c1 = model.layers[0], model.layers[2]
    assert torch.equal(c0.inv_mask, c1.mask) and torch.equal(c0.mask, c1.inv_mask)


@pytest.mark.parametrize("make", [lambda: RealNVP1d(8, n_flows=2, depth=1, units=16),
                                  lambda: MAF(8, n_flows=2, depth=1, units=16)])
def test_flows_fail_loudly_on_cpu_tensors(make):
    model = make().eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        model(torch.rand(4, 8))
