"""GPU parity of dpk_linear_forward (the conditioner MLP of the 1-D couplings on tcgen05, csrc/ratspn_leaf_mma.cu)
against float64 matmul, and of RealNVP1d evaluated through it against the cuBLAS path.
Tolerance: the 22-bit hi/lo operands are exact to 2e-7, but the tensor core adds into its fp32 accumulator with
truncation, once per 16-element K step and pass, so the error grows ~ K/16 * 3 * 2^-24 relative to the running
sum (3e-5 at K = 3072); still far inside the 1e-4 of the north star."""
import pytest
import torch

from conftest import rel_err
from deeprob_kit_b200.flows import _engine

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("batch,k,n,relu", [(300, 64, 40, True), (1000, 3072, 512, True), (513, 512, 6144, False),
                                              (128, 36, 7, False), (2049, 100, 300, True)])
def test_linear_matches_float64(batch, k, n, relu):
    g = torch.Generator(device=DEV).manual_seed(batch + k)
    x = torch.randn(batch, k, device=DEV, generator=g) * 3.0
    w = torch.randn(n, k, device=DEV, generator=g) / k ** 0.5
    b = torch.randn(n, device=DEV, generator=g)
    cache = {}
    out = _engine.linear(x, w, b, relu, cache)
    ref = x.double() @ w.double().t() + b.double()
    if relu:
        ref = ref.clamp_min(0.0)
    scale = float(ref.abs().max())
    tol = 2e-6 + 3.0 * (k / 16) * 2.0 ** -24
    assert float((out.double() - ref).abs().max()) / scale < tol
    # weight images are cached on (pointer, version): same result, and an in-place update is picked up
    assert torch.equal(_engine.linear(x, w, b, relu, cache), out)
    w.mul_(2.0)
    out2 = _engine.linear(x, w, None, relu, cache)
    ref2 = x.double() @ w.double().t()
    if relu:
        ref2 = ref2.clamp_min(0.0)
    assert float((out2.double() - ref2).abs().max()) / float(ref2.abs().max()) < tol


def test_linear_out_of_range_rows_are_exact():
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(400, 128, device=DEV, generator=g)
    x[7, 3] = float("nan")
    x[40, :] *= 1.0e6
    x[399, 100] = float("inf")
    w = torch.randn(96, 128, device=DEV, generator=g)
    b = torch.randn(96, device=DEV, generator=g)
    out = _engine.linear(x, w, b, True, {})
    ref = torch.relu(torch.nn.functional.linear(x, w, b))
    assert torch.isnan(out[7]).all() and torch.isnan(ref[7]).all()
    ok = torch.ones(400, dtype=torch.bool, device=DEV)
    ok[7] = False
    ok[399] = False
    ok[40] = False      # huge row: exact fp32 path, compared relative to the row's magnitude (summation order differs)
    assert rel_err(out[ok], ref[ok]) < 5e-5
    assert float((out[40] - ref[40]).abs().max() / ref[40].abs().max()) < 1e-5
    assert torch.equal(torch.isnan(out[399]), torch.isnan(ref[399]))
    fin = ~torch.isnan(ref[399])
    assert torch.equal(out[399][fin], ref[399][fin])


def test_realnvp1d_through_the_tensor_core_mlp(monkeypatch):
    from deeprob_kit_b200.flows.models import RealNVP1d
    torch.manual_seed(3)
    model = RealNVP1d(256, n_flows=4, depth=2, units=128, batch_norm=True, affine=True).to(DEV).eval()
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn_like(p))
    x = torch.rand(4096, 256, device=DEV)
    monkeypatch.setenv("DPK_LINEAR_MMA", "1")
    ll1 = model(x)
    monkeypatch.setenv("DPK_LINEAR_MMA", "0")
    ll0 = model(x)
    assert rel_err(ll1, ll0) < 1e-5
