"""Top-down passes of the RAT-SPN (RatSpn.mpe / RatSpn.sample, deeprob/spn/models/ratspn.py:124-182).
CPU: the oracle's MPE restatement reproduces the reference outputs stored in tests/golden/ratspn_mpe.npz bit for bit.
GPU: the one-thread-per-sample kernels (csrc/ratspn_topdown.cu) against the oracle element by element -- also with
padding, where the reference itself is no ground truth -- and the sampler against exact probabilities."""
import numpy as np
import pytest
import torch

import param_gen as pg
from conftest import load_golden
from helpers import oracle_for, product_model

DEV = "cuda:0"


@pytest.mark.parametrize("name", sorted(pg.MPE_CASES))
def test_oracle_mpe_reproduces_reference(name):
    cfg = pg.MPE_CASES[name]
    gold = load_golden("ratspn_mpe")
    orc, _ = oracle_for(cfg)
    x, _ = pg.ratspn_inputs(cfg)
    assert torch.equal(orc.mpe(x), torch.from_numpy(gold[name]))
    if cfg["out_classes"] > 1:
        y = torch.arange(x.shape[0]) % cfg["out_classes"]
        assert torch.equal(orc.mpe(x, y), torch.from_numpy(gold[name + ".y"]))


PADDED = {
    "mpe_pad15": dict(pg.RATSPN_CASES["bern15_nan"]),
    "mpe_pad37": dict(pg.RATSPN_CASES["gauss_cls"], nan_frac=0.5),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted({**pg.MPE_CASES, **PADDED}))
def test_mpe_kernel_matches_oracle_and_reference(name):
    cfg = {**pg.MPE_CASES, **PADDED}[name]
    model = product_model(cfg, DEV)
    orc, _ = oracle_for(cfg)
    x, _ = pg.ratspn_inputs(cfg)
    got = model.mpe(x.to(DEV)).cpu()
    obs = ~torch.isnan(x)
    assert torch.equal(got[obs], x[obs]) and not bool(torch.isnan(got).any())
    ref = orc.mpe(x)
    # the arg-max of two paths whose values agree to fp32 rounding may differ between the CUDA forward and the oracle:
    # allow a handful of rows to differ, but only if their completion is as likely as the oracle's
    diff_rows = (got != ref).any(dim=1)
    if bool(diff_rows.any()):
        assert int(diff_rows.sum()) <= max(1, x.shape[0] // 20), int(diff_rows.sum())
        # a differing row must be a near-tie: of the classes (the class is the arg-max of the evidence's own
        # log-likelihoods) or of two paths -- its completion is then as likely as the oracle's under the class it chose
        ll = orc.log_prob(x[diff_rows])
        top2 = torch.topk(ll, min(2, ll.shape[1]), dim=1).values
        class_tie = (top2[:, 0] - top2[:, -1]).abs() <= 1e-3 if ll.shape[1] > 1 else torch.zeros(ll.shape[0], dtype=torch.bool)
        a, b = orc.log_prob(got[diff_rows]).max(1).values, orc.log_prob(ref[diff_rows]).max(1).values
        assert bool((class_tie | ((a - b).abs() <= 1e-3 * b.abs().clamp_min(1.0))).all())
    if name in pg.MPE_CASES:
        gold = torch.from_numpy(load_golden("ratspn_mpe")[name])
        assert int((got != gold).any(dim=1).sum()) <= max(1, x.shape[0] // 20)
    if cfg["out_classes"] > 1:
        y = torch.arange(x.shape[0]) % cfg["out_classes"]
        goty = model.mpe(x.to(DEV), y.to(DEV)).cpu()
        assert int((goty != orc.mpe(x, y)).any(dim=1).sum()) <= max(1, x.shape[0] // 20)
    # the layer-by-layer index walk of the module API gives the same completion
    assert int((model.mpe_layerwise(x.to(DEV)).cpu() != got).any(dim=1).sum()) <= max(1, x.shape[0] // 20)


@pytest.mark.gpu
def test_sampler_reproduces_the_exact_distribution():
    """Bernoulli RAT-SPN over 6 binary variables: the empirical frequencies of 400k ancestral samples against
    exp(log_prob) of all 64 states (which the model normalises to 1)."""
    from deeprob_kit_b200.spn.models import BernoulliRatSpn
    torch.manual_seed(3)
    model = BernoulliRatSpn(6, rg_depth=2, rg_repetitions=3, rg_batch=3, rg_sum=2, random_state=5).to(DEV).eval()
    with torch.no_grad():
        model.base_layer.logits.mul_(2.0)
    states = torch.tensor([[(s >> i) & 1 for i in range(6)] for s in range(64)], dtype=torch.float32, device=DEV)
    p = model(states).exp().flatten().double().cpu()
    assert abs(float(p.sum()) - 1.0) < 1e-4
    n = 400_000
    torch.manual_seed(11)
    s = model.sample(n)
    assert s.shape == (n, 6) and bool(((s == 0) | (s == 1)).all())
    code = (s.long() * (2 ** torch.arange(6, device=DEV))).sum(1)
    freq = torch.bincount(code, minlength=64).double().cpu() / n
    sigma = torch.sqrt(p * (1 - p) / n)
    assert bool(((freq - p).abs() <= 5 * sigma + 1e-5).all()), float(((freq - p).abs() / sigma).max())
    # a second call draws a different stream
    assert not torch.equal(model.sample(1000), model.sample(1000))


@pytest.mark.gpu
def test_gaussian_sampler_moments_and_classes():
    """Gaussian leaves: per-class samples have the mean / second moment of the class mixture, estimated from the
    model's own top-down structure by importance-free Monte Carlo on the oracle side is not available -- instead the
    exact mixture moments are computed by brute force for a depth-1 model (root over products of two leaf regions)."""
    from deeprob_kit_b200.spn.models import GaussianRatSpn
    torch.manual_seed(5)
    model = GaussianRatSpn(4, out_classes=2, rg_depth=1, rg_repetitions=2, rg_batch=3, rg_sum=2, random_state=9,
                           optimize_scale=True).to(DEV).eval()
    w = torch.softmax(model.root_layer.weight.detach(), dim=1).cpu().double()      # (C, R*K^2)
    loc, scale = model.base_layer.loc.detach().cpu().double(), model.base_layer.scale.detach().cpu().double()
    regions = [list(r) for r in model.rg_layers[0]]
    k = 3
    for c in range(2):
        mean = torch.zeros(4, dtype=torch.float64)
        second = torch.zeros(4, dtype=torch.float64)
        for r in range(2):
            for i in range(k):
                for j in range(k):
                    pr = w[c, r * k * k + i * k + j]
                    for g, ch in ((2 * r, i), (2 * r + 1, j)):
                        for q, f in enumerate(regions[g]):
                            mean[f] += pr * loc[g, ch, q]
                            second[f] += pr * (loc[g, ch, q] ** 2 + scale[g, ch, q] ** 2)
        n = 300_000
        y = torch.full((n,), c, dtype=torch.long)
        s = model.sample(n, y).double().cpu()
        assert bool(torch.isfinite(s).all())
        assert float((s.mean(0) - mean).abs().max()) < 0.02
        assert float(((s * s).mean(0) - second).abs().max()) < 0.05
