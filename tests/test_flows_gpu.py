"""GPU parity of the flow models (fused bijector kernels of csrc/flows.cu) against the CPU oracle and the
reference golden vectors; tolerance 1e-4 relative fp32 on log-likelihoods."""
import numpy as np
import pytest
import torch

import param_gen as pg
from conftest import load_golden, norm_err, rel_err
from helpers import flow_reference_state
from oracle.flows_oracle import flow1d_log_prob

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"


@pytest.mark.parametrize("name", sorted(pg.FLOW_CASES))
def test_log_prob_and_gradients(name):
    cfg = pg.FLOW_CASES[name]
    gold = load_golden("flows_" + name)
    model, state = flow_reference_state(cfg, name)
    model = model.to(DEV).eval()
    x, g = pg.flow_inputs(cfg)
    with torch.enable_grad():
        xd = x.to(DEV).requires_grad_(True)
        out = model(xd)
        (out * g.to(DEV)).sum().backward()
    assert out.shape == (cfg["batch"],)
    assert rel_err(out.detach().cpu(), gold["ll"]) < TOL
    gtol = TOL + 4e-7 * float(out.abs().max())
    assert norm_err(xd.grad.cpu()[:4], gold["grad.x"]) < 2 * gtol
    grads = dict(model.named_parameters())
    for k, ref in gold.items():
        if k.startswith("grad.") and k != "grad.x":
            assert norm_err(grads[k[5:]].grad, ref) < 2 * gtol, k
    if cfg["model"] != "RealNVP2d":
        with torch.enable_grad():
            xx = x.double().requires_grad_(True)
            st64 = {k: (v.double() if v.is_floating_point() else v) for k, v in state.items()}
            ll, _ = flow1d_log_prob(xx, st64, cfg["model"], cfg["kw"])
            (ll * g.double()).sum().backward()
        assert rel_err(out.detach().cpu(), ll.detach()) < TOL
        assert norm_err(xd.grad.cpu(), xx.grad) < gtol


@pytest.mark.parametrize("name", sorted(pg.FLOW_CASES))
def test_invertibility(name):
    """deeprob-kit tests/test_flows.py:22-26: apply_forward(apply_backward(x)) == x and ildj == -ldj (atol 5e-7
    in the reference's float32 CPU run; the MADE sampling loop and the conv stacks get 2e-5 here)."""
    cfg = pg.FLOW_CASES[name]
    model, _ = flow_reference_state(cfg, name)
    model = model.to(DEV).eval()
    x, _ = pg.flow_inputs(cfg)
    gold = load_golden("flows_" + name)
    with torch.no_grad():
        z, _ = model.preprocess(x.to(DEV))
        u, ildj = model.apply_backward(z)
        back, ldj = model.apply_forward(u)
    assert norm_err(u.cpu()[:4], gold["u"]) < TOL
    ildj_t = ildj if isinstance(ildj, torch.Tensor) else torch.full((x.shape[0],), float(ildj))
    assert rel_err(ildj_t.cpu(), gold["ildj"]) < TOL
    assert float((back - z).abs().max()) < 2e-5 * max(1.0, float(z.abs().max()))
    if isinstance(ildj, torch.Tensor):
        assert float((ildj + ldj).abs().max()) < 1e-4 * max(1.0, float(ildj.abs().max()))


@pytest.mark.parametrize("name", ["nvp1d_small", "nvp1d_cifar", "maf_seq"])
def test_training_mode_batch_statistics(name):
    cfg = pg.FLOW_CASES[name]
    gold = load_golden("flows_" + name)
    model, _ = flow_reference_state(cfg)
    model = model.to(DEV).train()
    x, g = pg.flow_inputs(cfg)
    with torch.enable_grad():
        out = model(x.to(DEV))
        (out * g.to(DEV)).sum().backward()
    assert rel_err(out.detach().cpu(), gold["train.ll"]) < TOL
    gtol = TOL + 4e-7 * float(out.abs().max())
    sd = model.state_dict()
    for k, ref in gold.items():
        if k.startswith("train.state."):
            assert norm_err(sd[k[len("train.state."):]], ref) < 1e-5, k
    grads = dict(model.named_parameters())
    for k, ref in gold.items():
        if k.startswith("train.grad."):
            assert norm_err(grads[k[len("train.grad."):]].grad, ref) < 3 * gtol, k


def test_uniform_base_known_answer():
    """deeprob-kit tests/test_flows.py:134-156: with a Uniform(0, 10) base every sample has LL = D*log(1/10)
    once the preprocessing log-dets are removed."""
    from deeprob_kit_b200.flows.models import RealNVP1d
    torch.manual_seed(42)
    d = 16
    base = torch.distributions.Uniform(torch.full((d,), -5.0, device=DEV), torch.full((d,), 5.0, device=DEV))
    model = RealNVP1d(d, dequantize=True, logit=0.05, in_base=base, n_flows=1, depth=1, units=8, batch_norm=False).to(DEV).eval()
    with torch.no_grad():
        for p in model.layers.parameters():
            p.zero_()                          # identity coupling: u = x, ildj = 0
    x = torch.rand(64, d, device=DEV)
    torch.manual_seed(7)
    ll = model(x)
    torch.manual_seed(7)
    z, ildj = model.preprocess(x)
    assert torch.allclose(ll - ildj, torch.full((64,), d * np.log(0.1), device=DEV), atol=1e-4)
    assert bool(((z > -5) & (z < 5)).all())


def test_rsample_backward_and_spn_base():
    from deeprob_kit_b200.flows.models import RealNVP1d
    from deeprob_kit_b200.spn.models import GaussianRatSpn
    torch.manual_seed(0)
    model = RealNVP1d(10, n_flows=2, depth=1, units=16).to(DEV).train()
    with torch.enable_grad():
        model.rsample(64).mean().backward()    # tests/test_flows.py:29-33
    assert all(p.grad is not None for p in model.layers.parameters() if p.requires_grad)
    # a RAT-SPN as the base density (examples/ratspn_nvp1d_mnist.py:35-54): needs dLL/dx of the SPN kernels
    spn = GaussianRatSpn(10, rg_depth=2, rg_repetitions=3, rg_batch=4, rg_sum=3, random_state=42)
    flow = RealNVP1d(10, in_base=spn, n_flows=2, depth=1, units=16).to(DEV).eval()
    with torch.no_grad():
        for name, p in flow.named_parameters():
            if "scale_act" in name:
                p.fill_(0.5)
    x = torch.randn(32, 10, device=DEV)
    with torch.enable_grad():
        ll = flow(x)
        ll.sum().backward()
    assert ll.shape == (32,) and bool(torch.isfinite(ll).all())
    g = flow.layers[0].network[0].weight.grad
    assert g is not None and bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0
    # finite-difference check of one conditioner weight through the SPN base
    w = flow.layers[0].network[0].weight
    eps = 1e-2
    with torch.no_grad():
        base = float(flow(x).sum())
        w[0, 0] += eps
        plus = float(flow(x).sum())
        w[0, 0] -= eps
    assert abs((plus - base) / eps - float(g[0, 0])) < 5e-2 * max(1.0, abs(float(g[0, 0])))
