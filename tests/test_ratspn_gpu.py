"""GPU parity of the CUDA RAT-SPN path (through the C ABI) against the CPU oracle and the golden vectors.
Tolerance: 1e-4 relative fp32 (BASELINE.json north_star); observed errors are ~1e-6."""
import numpy as np
import pytest
import torch

import param_gen as pg
from conftest import load_golden, norm_err, rel_err
from helpers import oracle_for, product_model, subsample_like

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"


@pytest.mark.parametrize("name", sorted(pg.RATSPN_CASES))
def test_log_prob_matches_oracle_and_golden(name):
    cfg = pg.RATSPN_CASES[name]
    model = product_model(cfg, DEV)
    orc, _ = oracle_for(cfg)
    x, _ = pg.ratspn_inputs(cfg)
    out = model(x.to(DEV)).cpu()
    assert out.shape == (cfg["batch"], cfg["out_classes"])
    assert rel_err(out, orc.log_prob(x)) < TOL
    assert rel_err(out, load_golden("ratspn_" + name)["ll"]) < TOL
    # log_prob is the same call
    assert torch.equal(model.log_prob(x.to(DEV)).cpu(), out)
    if cfg["kind"] == "gaussian" and not cfg.get("optimize_scale", False):
        # frozen scale == 1 selects the unit-scale kernels: same values
        unit = product_model(cfg, DEV, scale_grad=False)
        assert unit.base_layer.unit_scale()
        out_u = unit(x.to(DEV)).cpu()
        assert rel_err(out_u, orc.log_prob(x)) < TOL
        with torch.enable_grad():
            xd = x.to(DEV).requires_grad_(True)
            unit(xd).sum().backward()
        ref = orc.grads(x, torch.ones_like(out), clean_nan=True)
        tol = TOL + 4e-7 * float(out.abs().max())
        assert norm_err(unit.base_layer.loc.grad, ref["loc"]) < 2 * tol
        assert norm_err(torch.nan_to_num(xd.grad.cpu()), torch.nan_to_num(ref["x"])) < 2 * tol


@pytest.mark.parametrize("name", sorted(pg.RATSPN_CASES))
def test_gradients_match_oracle_and_golden(name):
    cfg = pg.RATSPN_CASES[name]
    model = product_model(cfg, DEV)
    orc, _ = oracle_for(cfg)
    x, g = pg.ratspn_inputs(cfg)
    with torch.enable_grad():
        xd = x.to(DEV).requires_grad_(True)
        out = model(xd)
        (out * g.to(DEV)).sum().backward()
    ref = orc.grads(x, g, clean_nan=True)
    gold = load_golden("ratspn_" + name)
    assert rel_err(out.detach(), ref["out"]) < TOL
    # Gradients: every posterior is exp(difference of log-values of magnitude |LL|), so fp32 carries an
    # inherent relative error of ~|LL| * 2^-24 (the reference's own autograd has the same); the float64
    # oracle is the ground truth and the tolerance is widened accordingly.
    gtol = TOL + 4e-7 * float(ref["out"].abs().max())
    truth = oracle_for(cfg)[0].double().grads(x.double(), g.double(), clean_nan=True)
    TOLG = gtol
    ref = {k: ([t.float() for t in v] if isinstance(v, list) else v.float()) for k, v in truth.items()}
    gx = torch.nan_to_num(xd.grad.cpu())
    assert norm_err(gx, torch.nan_to_num(ref["x"])) < TOLG
    base = model.base_layer
    mine = {"root": model.root_layer.weight.grad, "sums": [l.weight.grad for l in model._sum_layers()]}
    if cfg["kind"] == "gaussian":
        mine["loc"], mine["scale"] = base.loc.grad, base.scale.grad
    else:
        mine["logits"] = base.logits.grad
    for key in ("loc", "scale", "logits", "root"):
        if key in mine:
            assert norm_err(mine[key], ref[key]) < TOLG, key
    for a, b in zip(mine["sums"], ref["sums"]):
        assert norm_err(a, b) < TOLG
    # and against the reference's own autograd (NaN entries of the reference are unpinned)
    names = {"loc": "grad.base_layer.loc", "scale": "grad.base_layer.scale", "logits": "grad.base_layer.logits",
             "root": "grad.root_layer.weight"}
    for key, gk in names.items():
        if key in mine and gk in gold:
            assert norm_err(subsample_like(mine[key].cpu(), gold[gk].size), gold[gk].reshape(-1)) < 2 * TOLG, gk
    if cfg["nan_frac"] == 0:
        assert norm_err(gx[: gold["grad.x"].shape[0]], gold["grad.x"]) < 2 * TOLG


def test_normalisation_over_all_binary_states():
    """deeprob-kit tests/test_ratspn.py:46-48: sum_x exp(ll(x)) == 1 over the 2^15 complete assignments."""
    from deeprob_kit_b200.spn.models import BernoulliRatSpn
    torch.manual_seed(42)
    model = BernoulliRatSpn(15, rg_depth=3, rg_repetitions=4, rg_batch=4, rg_sum=2, random_state=42).eval().to(DEV)
    n = 15
    data = ((torch.arange(2 ** n).unsqueeze(1) >> torch.arange(n - 1, -1, -1)) & 1).float()
    ll = model(data.to(DEV)).double()
    assert np.isclose(float(ll.exp().sum()), 1.0, rtol=1e-5)
    # marginalising a variable must equal summing it out
    half = data[: 2 ** (n - 1)].clone()
    half[:, 0] = float("nan")
    ll_m = model(half.to(DEV)).double().cpu()
    both = torch.logsumexp(torch.stack([ll[: 2 ** (n - 1)].cpu(), ll[2 ** (n - 1):].cpu()]), 0)
    assert rel_err(ll_m, both) < 1e-5


def test_extreme_weights_use_exact_path():
    """Mixture weights spanning > e^87: the linear-domain sum underflows, the log-domain fallback must kick in."""
    cfg = dict(pg.RATSPN_CASES["gauss_cls"])
    model = product_model(cfg, DEV)
    orc, state = oracle_for(cfg)
    rng = np.random.RandomState(5)
    for k in [k for k in state if k.endswith(".weight")]:
        state[k] = state[k] * 1.0 + torch.from_numpy(rng.choice([0.0, -150.0, -60.0], size=state[k].shape)).float()
    orc.load_reference_state(state)
    model.load_state_dict({**model.state_dict(), **{k: v for k, v in state.items() if k.endswith(".weight")}})
    x, _ = pg.ratspn_inputs(cfg)
    assert rel_err(model(x.to(DEV)).cpu(), orc.log_prob(x)) < TOL


def test_infinite_inputs_follow_nan_to_num():
    cfg = pg.RATSPN_CASES["gauss_d1"]
    model = product_model(cfg, DEV)
    orc, _ = oracle_for(cfg)
    x, _ = pg.ratspn_inputs(cfg)
    x[3, 2] = float("inf")
    x[5, 0] = float("-inf")
    x[7, :] = float("inf")
    out, ref = model(x.to(DEV)).cpu(), orc.log_prob(x)
    finite = torch.isfinite(ref) & (ref > -1e30)
    assert rel_err(out[finite], ref[finite]) < TOL
    assert bool(((out < -1e30) == (ref < -1e30)).all())


def test_batch_sizes_and_empty():
    cfg = pg.RATSPN_CASES["gauss_cls"]
    model = product_model(cfg, DEV)
    orc, _ = oracle_for(cfg)
    rng = np.random.RandomState(0)
    for b in (0, 1, 31, 33, 129, 700):
        x = torch.from_numpy(rng.standard_normal((b, cfg["in_features"])).astype(np.float32))
        out = model(x.to(DEV)).cpu()
        assert out.shape == (b, cfg["out_classes"])
        if b:
            assert rel_err(out, orc.log_prob(x)) < TOL


def test_north_star_shape_large_batch_properties():
    """BASELINE config 2 at a batch the oracle cannot finish quickly: size-independent properties."""
    cfg = dict(pg.RATSPN_CASES["gauss784"])
    model = product_model(cfg, DEV)
    orc, _ = oracle_for(cfg)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(16384 + 37, 784, generator=g)
    out = model(x.to(DEV))
    # (1) batch independence: any slice evaluated alone gives the same values (the small slice runs the CUDA-core
    # leaf kernel, the big batch the tensor-core GEMM: equal to fp32 rounding, not bitwise)
    part = model(x[5000:5300].to(DEV))
    assert rel_err(part, out[5000:5300]) < 1e-5
    # (2) permutation equivariance over the batch
    perm = torch.randperm(x.shape[0], generator=g)
    assert torch.equal(model(x[perm].to(DEV)), out[perm.to(DEV)])
    # (3) oracle on a strided subset
    idx = torch.arange(0, x.shape[0], 97)
    assert rel_err(out[idx.to(DEV)].cpu(), orc.log_prob_chunked(x[idx], 64)) < TOL
    # (4) a fully marginalised row integrates to 1
    xm = torch.full((4, 784), float("nan"))
    assert float(model(xm.to(DEV)).abs().max()) < 1e-3


def test_em_statistics_match_oracle():
    for name in ("gauss784", "bern15_nan", "gauss_d1"):
        cfg = pg.RATSPN_CASES[name]
        model = product_model(cfg, DEV)
        orc, _ = oracle_for(cfg)
        x, _ = pg.ratspn_inputs(cfg)
        st = model.em_statistics(x.to(DEV))
        ref = orc.double().em_statistics(x.double())     # float64 ground truth, see the gradient test
        tol = TOL + 4e-7 * float(st["ll"].abs().max())
        assert rel_err(st["ll"].sum().cpu(), ref["ll_sum"]) < TOL
        assert norm_err(st["root_counts"], ref["root_counts"]) < tol
        for a, b in zip(st["sum_counts"], ref["sum_counts"]):
            assert norm_err(a, b) < tol
        assert norm_err(st["s0"], ref["s0"]) < tol
        assert norm_err(st["s1"], ref["s1"]) < tol
        if cfg["kind"] == "gaussian":
            assert norm_err(st["s2"], ref["s2"]) < tol
        # posterior counts of every sum node add up to the number of samples reaching it
        assert abs(float(st["root_counts"].sum()) - x.shape[0]) < 1e-2 * x.shape[0]


def test_standalone_layers_and_mpe():
    from oracle.ratspn_oracle import product_layer, root_layer, sum_layer
    cfg = pg.RATSPN_CASES["gauss_cls"]
    model = product_model(cfg, DEV)
    orc, _ = oracle_for(cfg)
    x, _ = pg.ratspn_inputs(cfg)
    h_ref = orc.leaf(x)
    h = model.base_layer(x.to(DEV))
    assert rel_err(h.cpu(), h_ref) < TOL
    si = 0
    for layer in model.layers:
        h = layer(h)
        if layer.__class__.__name__ == "ProductLayer":
            h_ref = product_layer(h_ref)
        else:
            h_ref = sum_layer(h_ref, orc.sum_weights[si])
            si += 1
        assert h.shape == h_ref.shape
        assert rel_err(h.cpu(), h_ref) < TOL
    assert rel_err(model.root_layer(h).cpu(), root_layer(h_ref, orc.root_weight)) < TOL
    # mpe: observed entries are kept, missing ones are filled with finite values, and the completion
    # is at least as likely as the same rows completed with the leaf means of a random path
    filled = model.mpe(x.to(DEV))
    obs = ~torch.isnan(x)
    assert torch.equal(filled.cpu()[obs], x[obs])
    assert bool(torch.isfinite(filled).all())
    samples = model.sample(16)
    assert samples.shape == (16, cfg["in_features"]) and bool(torch.isfinite(samples).all())


def test_em_steps_increase_likelihood():
    """Batch EM (extension, config 5 on one rank): E-step kernels + M-step raise the data likelihood."""
    from deeprob_kit_b200.spn import em
    from deeprob_kit_b200.spn.models import BernoulliRatSpn, GaussianRatSpn
    gen = torch.Generator().manual_seed(1)
    centers = torch.randn(3, 16, generator=gen) * 2.0
    x = (centers[torch.randint(0, 3, (4096,), generator=gen)] + 0.3 * torch.randn(4096, 16, generator=gen)).to(DEV)
    torch.manual_seed(0)
    model = GaussianRatSpn(16, rg_depth=2, rg_repetitions=4, rg_batch=4, rg_sum=3, random_state=42, optimize_scale=True).to(DEV)
    hist = [float(model(x).mean())]
    for _ in range(6):
        reported = em.em_step(model, x, step_size=0.5)
        assert abs(reported - hist[-1]) < 1e-3 * max(1.0, abs(hist[-1]))     # em_step returns the pre-update mean LL
        hist.append(float(model(x).mean()))
    assert all(b > a - 1e-3 for a, b in zip(hist[:-1], hist[1:])), hist
    assert hist[-1] > hist[0] + 5.0, hist
    xb = (torch.rand(2048, 12, generator=gen) < torch.rand(12, generator=gen)).float().to(DEV)
    bern = BernoulliRatSpn(12, rg_depth=2, rg_repetitions=3, rg_batch=3, rg_sum=2, random_state=1).to(DEV)
    l0 = float(bern(xb).mean())
    for _ in range(5):
        em.em_step(bern, xb, step_size=0.5)
    assert float(bern(xb).mean()) > l0 + 0.1


def test_em_step_with_frozen_unit_scale():
    """the default GaussianRatSpn (optimize_scale=False) has no second moment (s2 is None) and a frozen
    scale: the M-step re-estimates the means only, the scale stays 1, the likelihood still rises."""
    from deeprob_kit_b200.spn import em
    from deeprob_kit_b200.spn.models import GaussianRatSpn
    gen = torch.Generator().manual_seed(3)
    centers = torch.randn(3, 16, generator=gen) * 1.5
    x = (centers[torch.randint(0, 3, (2048,), generator=gen)] + torch.randn(2048, 16, generator=gen)).to(DEV)
    torch.manual_seed(0)
    model = GaussianRatSpn(16, rg_depth=2, rg_repetitions=4, rg_batch=4, rg_sum=3, random_state=42).to(DEV)
    assert model.base_layer.unit_scale()
    st = model.em_statistics(x)
    assert st["s2"] is None
    l0 = float(model(x).mean())
    for _ in range(4):
        em.em_step(model, x, step_size=0.5)
    assert float(model(x).mean()) > l0 + 1.0
    assert bool((model.base_layer.scale == 1).all())


def test_log_prob_host_result_is_complete_on_return():
    """the pinned result of log_prob_host can be read as soon as the call returns."""
    from deeprob_kit_b200.spn.streaming import log_prob_host
    cfg = pg.RATSPN_CASES["gauss_cls"]
    model = product_model(cfg, DEV)
    rng = np.random.RandomState(11)
    x = torch.from_numpy(rng.standard_normal((5000, cfg["in_features"])).astype(np.float32)).pin_memory()
    ref = model(x.to(DEV)).cpu()
    for _ in range(3):
        out = log_prob_host(model, x, chunk=1024)
        got = out.clone()                       # read immediately, no synchronize in between
        assert torch.equal(got, ref)
        out.fill_(float("nan"))                 # next round must overwrite every element again
    oh = torch.full((5000, cfg["out_classes"]), float("nan")).pin_memory()
    assert torch.equal(log_prob_host(model, x, chunk=777, out_host=oh).clone(), ref)
    empty = log_prob_host(model, x[:0])
    assert empty.shape == (0, cfg["out_classes"])


def test_cache_invalidation_and_copies():
    """a write through .data is invisible to the version-keyed table cache until invalidate_caches();
    deepcopy / pickling carry no workspaces."""
    import copy
    cfg = pg.RATSPN_CASES["gauss_d1"]
    model = product_model(cfg, DEV)
    x, _ = pg.ratspn_inputs(cfg)
    xd = x.to(DEV)
    a = model(xd).clone()
    model(xd)                                   # second call: tables cached
    model.base_layer.loc.data.add_(0.25)
    model.invalidate_caches()
    b = model(xd)
    assert not torch.allclose(a, b)
    twin = copy.deepcopy(model)
    assert twin._ws_cache == {} and torch.equal(twin(xd), b)


def test_em_statistics_config5_slice_4096():
    """BASELINE config 5 model (= config 2 structure) on a 4096-row slice: E-step statistics of the
    CUDA path (tensor-core leaf moments are selected by the batch size the bench uses; forced here) against the
    float64 oracle, accumulated over 512-row chunks (the statistics are sums over the batch)."""
    import os
    cfg = dict(pg.RATSPN_CASES["gauss784"], batch=4096)
    model = product_model(cfg, DEV, scale_grad=False)
    orc = oracle_for(cfg)[0].double()
    x = torch.randn(4096, 784, generator=torch.Generator().manual_seed(5))
    prev = os.environ.get("DPK_STATS_MMA")
    os.environ["DPK_STATS_MMA"] = "1"
    try:
        st = model.em_statistics(x.to(DEV))
    finally:
        if prev is None:
            del os.environ["DPK_STATS_MMA"]
        else:
            os.environ["DPK_STATS_MMA"] = prev
    ref = None
    for i in range(0, 4096, 512):
        part = orc.em_statistics(x[i:i + 512].double())
        if ref is None:
            ref = part
        else:
            for k in ("ll_sum", "root_counts", "s0", "s1", "s2"):
                ref[k] = ref[k] + part[k]
            ref["sum_counts"] = [a + b for a, b in zip(ref["sum_counts"], part["sum_counts"])]
    tol = TOL + 4e-7 * float(st["ll"].abs().max())
    assert rel_err(st["ll"].sum().cpu(), ref["ll_sum"]) < TOL
    assert norm_err(st["root_counts"], ref["root_counts"]) < tol
    for a, b in zip(st["sum_counts"], ref["sum_counts"]):
        assert norm_err(a, b) < tol
    assert norm_err(st["s0"], ref["s0"]) < tol
    assert norm_err(st["s1"], ref["s1"]) < tol
    assert st["s2"] is None          # frozen unit scale: no second moment (the M-step re-estimates the means only)
