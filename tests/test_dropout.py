"""Probabilistic dropout in training mode (deeprob/spn/layers/ratspn.py:98-100,370-372; layers/dgcspn.py:113-115,297-299).
CPU part: the generator exported by the library against its NumPy mirror and its statistics.  GPU part: the CUDA
training path against the oracle with the same draws injected (exact parity, values and gradients), all-dropped rows,
and the DGC-SPN layers."""
import ctypes

import numpy as np
import pytest
import torch

import dropout_ref as dr
import param_gen as pg
from conftest import norm_err, rel_err
from helpers import oracle_for, product_model

DEV = "cuda:0"
TOL = 1e-4


def test_generator_matches_numpy_mirror_and_is_uniform():
    from deeprob_kit_b200 import _lib
    lib = _lib.lib()
    rng = np.random.RandomState(0)
    for seed, stream in ((0, 0), (123456789123, 1), ((1 << 62) - 5, 3)):
        idx = rng.randint(0, 2 ** 40, size=200).astype(np.uint64)
        got = np.array([lib.dpk_dropout_draw(seed, stream, int(i)) for i in idx], dtype=np.uint32)
        assert np.array_equal(got, dr.draw(seed, stream, idx))
    u = dr.draw(42, 0, np.arange(2_000_000, dtype=np.uint64))
    assert u.max() < (1 << 24)
    for rate in (0.1, 0.2, 0.5):
        frac = float((u < dr.threshold(rate)).mean())
        assert abs(frac - rate) < 3e-3, (rate, frac)
    # neighbouring counters and neighbouring streams are uncorrelated
    a, b = u[:-1].astype(np.float64), u[1:].astype(np.float64)
    assert abs(np.corrcoef(a, b)[0, 1]) < 5e-3
    v = dr.draw(42, 1, np.arange(2_000_000, dtype=np.uint64)).astype(np.float64)
    assert abs(np.corrcoef(u.astype(np.float64), v)[0, 1]) < 5e-3


CASES = {
    "gauss": dict(kind="gaussian", in_features=21, rg_depth=2, rg_repetitions=3, rg_batch=4, rg_sum=3, out_classes=2,
                  batch=150, nan_frac=0.0, optimize_scale=True, in_dropout=0.2, sum_dropout=0.3),
    "bern_deep": dict(kind="bernoulli", in_features=19, rg_depth=3, rg_repetitions=2, rg_batch=3, rg_sum=2, out_classes=1,
                      batch=97, nan_frac=0.0, binary=True, in_dropout=0.35, sum_dropout=0.15),
    "gauss_d1_in_only": dict(kind="gaussian", in_features=9, rg_depth=1, rg_repetitions=3, rg_batch=5, rg_sum=3,
                             out_classes=1, batch=64, nan_frac=0.0, optimize_scale=False, in_dropout=0.5, sum_dropout=None),
    # nearly everything dropped in front of the sums: rows with no kept entry give -inf and no gradient
    "all_dropped": dict(kind="gaussian", in_features=16, rg_depth=2, rg_repetitions=2, rg_batch=2, rg_sum=2, out_classes=1,
                        batch=80, nan_frac=0.0, optimize_scale=True, in_dropout=None, sum_dropout=0.97),
}


def _train_model(cfg):
    from deeprob_kit_b200.spn.models import BernoulliRatSpn, GaussianRatSpn
    cls = GaussianRatSpn if cfg["kind"] == "gaussian" else BernoulliRatSpn
    kw = pg.ratspn_ctor_kwargs(cfg)
    model = cls(in_dropout=cfg["in_dropout"], sum_dropout=cfg["sum_dropout"], **kw)
    model.load_state_dict(pg.ratspn_fill_state(model.state_dict(), cfg, 0))
    return model.to(DEV).train()


def _masks(orc, cfg, seed, batch):
    g0, k, dim = len(orc.leaf_regions), orc.K, orc.dim
    leaf = torch.from_numpy(dr.dropped(seed, 0, (batch, g0, k, dim), cfg["in_dropout"] or 0.0))
    sums, groups, nodes = [], g0, k
    for e in range(cfg["rg_depth"] - 1):
        groups, kin2 = groups // 2, nodes * nodes
        sums.append(torch.from_numpy(dr.dropped(seed, 1 + e, (batch, groups, kin2), cfg["sum_dropout"] or 0.0)))
        nodes = orc.O
    return leaf, sums


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_training_dropout_matches_oracle_with_injected_draws(name):
    cfg = CASES[name]
    model = _train_model(cfg)
    orc, _ = oracle_for(cfg)
    x, g = pg.ratspn_inputs(cfg)
    torch.manual_seed(1234)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())     # the draw _RatSpnLogProbDropout.forward makes
    torch.manual_seed(1234)
    with torch.enable_grad():
        xd = x.to(DEV).requires_grad_(True)
        out = model(xd)
        finite = torch.isfinite(out)
        (torch.where(finite, out, torch.zeros_like(out)) * g.to(DEV)).sum().backward()
    orc.leaf_drop, orc.sum_drops = _masks(orc, cfg, seed, x.shape[0])
    orc.double()
    ref = orc.log_prob(x.double())
    assert bool((torch.isfinite(ref) == finite.cpu()).all())
    assert rel_err(out.detach().cpu()[finite.cpu()], ref[finite.cpu()]) < TOL
    if name == "all_dropped":
        assert not bool(finite.all()), "the case is meant to contain all-dropped rows"
    gm = torch.where(finite.cpu(), g.double(), torch.zeros_like(g, dtype=torch.float64))
    gr = orc.grads(x.double(), gm, clean_nan=True)
    tol = TOL + 4e-7 * float(out[finite].abs().max())
    p0 = model.base_layer.loc if cfg["kind"] == "gaussian" else model.base_layer.logits
    assert norm_err(p0.grad, gr["loc" if cfg["kind"] == "gaussian" else "logits"]) < tol
    if cfg["kind"] == "gaussian" and cfg.get("optimize_scale"):
        assert norm_err(model.base_layer.scale.grad, gr["scale"]) < tol
    assert norm_err(xd.grad, gr["x"]) < tol
    assert norm_err(model.root_layer.weight.grad, gr["root"]) < tol
    sums = [l for l in model.layers if hasattr(l, "weight")]
    for layer, ref_g in zip(sums, gr["sums"]):
        assert norm_err(layer.weight.grad, ref_g) < tol
    for p in model.parameters():
        if p.grad is not None:
            assert bool(torch.isfinite(p.grad).all())
    # a second forward draws new masks; eval mode ignores dropout and equals the oracle without masks
    out2 = model(x.to(DEV))
    assert not torch.equal(out2, out.detach())
    orc.leaf_drop = orc.sum_drops = None
    assert rel_err(model.eval()(x.to(DEV)).cpu(), orc.log_prob(x.double())) < TOL


@pytest.mark.gpu
def test_reference_example_configuration_trains():
    """examples/ratspn_mnist.py:29-38 shape (depth 3, 8 repetitions, K = 16, in/sum dropout 0.2) on synthetic data:
    a few Adam steps through the reference-style loop (loss, backward, step, apply_constraints) lower the loss."""
    from deeprob_kit_b200.spn.models import GaussianRatSpn
    torch.manual_seed(0)
    model = GaussianRatSpn(784, rg_depth=3, rg_repetitions=8, rg_batch=16, rg_sum=16, in_dropout=0.2, sum_dropout=0.2,
                           optimize_scale=True, random_state=42).to(DEV).train()
    gen = torch.Generator().manual_seed(1)
    centers = torch.randn(4, 784, generator=gen)
    x = (centers[torch.randint(0, 4, (100,), generator=gen)] + 0.5 * torch.randn(100, 784, generator=gen)).to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=5e-2)
    losses = []
    with torch.enable_grad():
        for _ in range(12):
            opt.zero_grad()
            loss = model.loss(model(x))
            loss.backward()
            opt.step()
            model.apply_constraints()
            losses.append(float(loss.detach()))
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < losses[0] - 20.0, losses


@pytest.mark.gpu
def test_dgcspn_training_dropout_matches_oracle_with_injected_draws():
    from deeprob_kit_b200.spn.models import DgcSpn
    from helpers import dgc_oracle_for
    name = "full8"
    cfg = dict(pg.DGCSPN_CASES[name])
    kw = pg.dgcspn_ctor_kwargs(cfg)
    model = DgcSpn(in_dropout=0.25, sum_dropout=0.2, **kw)
    names = [k for k, _ in model.named_parameters()]
    model.load_state_dict(pg.dgcspn_fill_state(model.state_dict(), names, cfg, 0))
    model = model.to(DEV).train()
    orc, _ = dgc_oracle_for(cfg)
    x, g = pg.dgcspn_inputs(cfg)
    x = torch.nan_to_num(x)
    # the layers draw with torch.rand_like on the device, in layer order: replay the same stream
    torch.manual_seed(99)
    with torch.enable_grad():
        xd = x.to(DEV).requires_grad_(True)
        out = model(xd)
        fin = torch.isfinite(out)
        (torch.where(fin, out, torch.zeros_like(out)) * g.to(DEV)).sum().backward()
    torch.manual_seed(99)
    b, (c, h, w), k = x.shape[0], cfg["in_features"], cfg["n_batch"]
    leaf = torch.stack([torch.rand(b, k, h, w, device=DEV) < 0.25 for _ in range(c)], dim=2).cpu()
    sums = [(torch.rand(b, *s[1:], device=DEV) < 0.2).cpu()
            for s in [(None, p["out_shape"][0], p["out_shape"][1], p["out_shape"][2]) for p in orc.products[:len(orc.sum_shapes)]]]
    orc.leaf_drop, orc.sum_drops = leaf, sums
    orc.double()
    ref = orc.log_prob(x.double())
    assert bool((torch.isfinite(ref) == fin.cpu()).all())
    assert rel_err(out.detach().cpu()[fin.cpu()], ref[fin.cpu()]) < TOL
    gr = orc.grads(x.double(), torch.where(fin.cpu(), g.double(), torch.zeros_like(g, dtype=torch.float64)))
    tol = TOL + 4e-7 * float(out[fin].abs().max())
    assert norm_err(model.base_layer.loc.grad, gr["loc"]) < 3 * tol
    assert norm_err(model.root_layer.weight.grad, gr["root"]) < 3 * tol
    assert bool(torch.isfinite(xd.grad).all())
