"""GPU parity of the tensor-core (tcgen05) leaf path, csrc/ratspn_leaf_mma.cu, against the float64 oracle.

The path is selected automatically for batches >= 8192; DPK_LEAF_MMA=1 forces it for the small batches
the oracle finishes in seconds, DPK_LEAF_MMA=0 gives the CUDA-core kernel to compare with.  Leaf values
are held to 2e-6 relative (the hi/lo fp16 split carries 22 bits), whole-model log-likelihoods to the
north-star 1e-4."""
import numpy as np
import pytest
import torch

import param_gen as pg
from conftest import norm_err, rel_err
from helpers import oracle_for, product_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CASES = {
    # north-star structure, batch not a multiple of the 256-sample tile
    "gauss784": dict(kind="gaussian", in_features=784, rg_depth=3, rg_repetitions=16, rg_batch=10, rg_sum=10,
                     out_classes=1, batch=700, nan_frac=0.0, optimize_scale=False),
    # padded regions (36 / 8 = 4.5), K not a multiple of anything, > 256 columns, 2 K blocks (last one short)
    "gauss36": dict(kind="gaussian", in_features=36, rg_depth=3, rg_repetitions=9, rg_batch=7, rg_sum=3,
                    out_classes=2, batch=300, nan_frac=0.0, optimize_scale=False),
    "gauss_deep": dict(pg.RATSPN_CASES["gauss_deep"], batch=513),
    # > 256 regions: two x^2 tiles
    "gauss_wide": dict(kind="gaussian", in_features=128, rg_depth=5, rg_repetitions=10, rg_batch=4, rg_sum=3,
                       out_classes=1, batch=260, nan_frac=0.0, optimize_scale=False),
    "bern784": dict(pg.RATSPN_CASES["bern784"], batch=333),
    "bern16": dict(pg.RATSPN_CASES["bern16"], batch=257),
    # marginalised inputs: flagged 32-sample groups are redone by the exact kernel
    "gauss784_nan": dict(kind="gaussian", in_features=784, rg_depth=3, rg_repetitions=16, rg_batch=10, rg_sum=10,
                         out_classes=1, batch=400, nan_frac=0.001, optimize_scale=False),
    # general (learnable) scale: x and x^2 images against mu/sigma^2 and -1/(2 sigma^2) images
    "gauss784_scale": dict(kind="gaussian", in_features=784, rg_depth=3, rg_repetitions=16, rg_batch=10, rg_sum=10,
                           out_classes=1, batch=300, nan_frac=0.0, optimize_scale=True),
    "gauss36_scale_nan": dict(kind="gaussian", in_features=36, rg_depth=3, rg_repetitions=9, rg_batch=7, rg_sum=3,
                              out_classes=2, batch=260, nan_frac=0.02, optimize_scale=True),
    "bern36_nan": dict(kind="bernoulli", in_features=36, rg_depth=3, rg_repetitions=9, rg_batch=7, rg_sum=3,
                       out_classes=2, batch=300, nan_frac=0.01, binary=True),
}


def _run(cfg, monkeypatch, mma):
    monkeypatch.setenv("DPK_LEAF_MMA", "1" if mma else "0")
    model = product_model(cfg, DEV, scale_grad=False)
    x, g = pg.ratspn_inputs(cfg)
    xd = x.to(DEV)
    leaf = model.base_layer(xd).cpu()
    out = model(xd).cpu()
    return model, x, g, leaf, out


@pytest.mark.parametrize("name", sorted(CASES))
def test_mma_leaf_matches_float64_oracle(name, monkeypatch):
    cfg = CASES[name]
    orc = oracle_for(cfg)[0].double()
    _, x, _, leaf, out = _run(cfg, monkeypatch, True)
    ref_leaf = orc.leaf(x.double())
    ref_out = orc.log_prob(x.double())
    assert leaf.shape == ref_leaf.shape
    # general scale: twice the accumulation steps, terms x^2/(2 sigma^2) and x mu/sigma^2 up to 8x larger and cancelling
    tol = 1e-5 if cfg.get("optimize_scale", False) else 2e-6
    assert rel_err(leaf, ref_leaf) < tol, "leaf"
    assert rel_err(out, ref_out) < 1e-4, "log_prob"
    # the CUDA-core kernel on the same inputs: both paths agree to fp32 rounding
    _, _, _, leaf0, out0 = _run(cfg, monkeypatch, False)
    assert rel_err(leaf, leaf0) < tol
    assert rel_err(out, out0) < tol


def test_mma_leaf_out_of_range_inputs_take_the_exact_path(monkeypatch):
    cfg = CASES["gauss784"]
    monkeypatch.setenv("DPK_LEAF_MMA", "1")
    model = product_model(cfg, DEV, scale_grad=False)
    orc = oracle_for(cfg)[0]  # fp32 oracle: nan_to_num(-inf) is -FLT_MAX there, like the reference
    x, _ = pg.ratspn_inputs(cfg)
    x[3, 5] = 1.0e4          # x^2 leaves the fp16 range
    x[40, 700] = float("inf")
    x[699, 0] = float("nan")
    x[300:310, :] *= 300.0
    out = model(x.to(DEV)).cpu()
    ref = orc.log_prob(x)
    assert rel_err(out, ref) < 1e-4


def test_mma_leaf_huge_parameters_take_the_exact_path(monkeypatch):
    cfg = CASES["gauss36"]
    monkeypatch.setenv("DPK_LEAF_MMA", "1")
    model = product_model(cfg, DEV, scale_grad=False)
    orc, state = oracle_for(cfg)
    with torch.no_grad():
        model.base_layer.loc[2, 1, 0] = 1.0e5
    state = dict(state)
    state["base_layer.loc"] = model.base_layer.loc.detach().cpu().clone()
    orc = orc.load_reference_state(state).double()
    x, _ = pg.ratspn_inputs(cfg)
    assert rel_err(model.base_layer(x.to(DEV)).cpu(), orc.leaf(x.double())) < 1e-5


@pytest.mark.parametrize("name", ["gauss784", "bern784", "gauss36"])
def test_mma_gradients_match_float64_oracle(name, monkeypatch):
    cfg = CASES[name]
    monkeypatch.setenv("DPK_LEAF_MMA", "1")
    model = product_model(cfg, DEV, scale_grad=False)
    x, g = pg.ratspn_inputs(cfg)
    with torch.enable_grad():
        xd = x.to(DEV).requires_grad_(True)
        out = model(xd)
        (out * g.to(DEV)).sum().backward()
    truth = oracle_for(cfg)[0].double().grads(x.double(), g.double(), clean_nan=True)
    tol = 1e-4 + 4e-7 * float(truth["out"].abs().max())
    key = "loc" if cfg["kind"] == "gaussian" else "logits"
    mine = model.base_layer.loc.grad if cfg["kind"] == "gaussian" else model.base_layer.logits.grad
    assert norm_err(mine, truth[key].float()) < tol
    assert norm_err(xd.grad.cpu(), truth["x"].float()) < tol
    assert norm_err(model.root_layer.weight.grad, truth["root"].float()) < tol


def test_mma_full_batch_properties(monkeypatch):
    """BASELINE config 2 at its full size (65536 x 784): size-independent checks -- the automatic path choice
    agrees with the CUDA-core kernel, is invariant to a permutation of the batch, and a row of the big batch
    equals the same row evaluated in a small batch."""
    cfg = dict(CASES["gauss784"], batch=65536)
    monkeypatch.delenv("DPK_LEAF_MMA", raising=False)
    model = product_model(cfg, DEV, scale_grad=False)
    g = torch.Generator(device=DEV).manual_seed(7)
    x = torch.randn(65536, 784, device=DEV, generator=g)
    out = model(x)
    monkeypatch.setenv("DPK_LEAF_MMA", "0")
    out0 = model(x)
    assert rel_err(out, out0) < 2e-6
    monkeypatch.delenv("DPK_LEAF_MMA", raising=False)
    perm = torch.randperm(65536, device=DEV, generator=g)
    assert rel_err(model(x[perm]), out[perm]) < 1e-6
    monkeypatch.setenv("DPK_LEAF_MMA", "0")
    assert rel_err(model(x[1000:1064]), out[1000:1064]) < 2e-6


def test_table_cache_follows_parameter_updates(monkeypatch):
    """DPK_F_TABLES_VALID: the derived tables are reused only while the parameters' version counters stand still."""
    cfg = CASES["gauss36"]
    monkeypatch.setenv("DPK_LEAF_MMA", "1")
    model = product_model(cfg, DEV, scale_grad=False)
    x, _ = pg.ratspn_inputs(cfg)
    xd = x.to(DEV)
    out0 = model(xd)
    assert torch.equal(model(xd), out0)                       # second call: cached tables, same bits
    with torch.no_grad():
        model.base_layer.loc.add_(0.25)                        # in-place update bumps the version -> rebuild
        model.root_layer.weight.mul_(0.5)
    out1 = model(xd)
    orc, state = oracle_for(cfg)
    state = dict(state)
    state["base_layer.loc"] = model.base_layer.loc.detach().cpu().clone()
    state["root_layer.weight"] = model.root_layer.weight.detach().cpu().clone()
    ref = orc.load_reference_state(state).double().log_prob(x.double())
    assert rel_err(out1, ref) < 1e-4
    assert rel_err(out0, ref) > 1e-3
    model.cache_tables = False
    assert torch.equal(model(xd), out1)
    # a different batch size re-plans the workspace: tables rebuilt, same values row by row
    assert rel_err(model(xd[:100]), out1[:100]) < 1e-6


# ---- product+sum levels with the contraction on the tensor cores (csrc/ratspn_einsum_mma.cu) -----------------
EINSUM_CASES = {
    "k10o10": CASES["gauss784"],
    "k4o3": CASES["gauss_wide"],
    "k8o8": dict(kind="gaussian", in_features=64, rg_depth=3, rg_repetitions=3, rg_batch=8, rg_sum=8,
                 out_classes=2, batch=384, nan_frac=0.0, optimize_scale=True),
    "k16o8": dict(kind="bernoulli", in_features=32, rg_depth=3, rg_repetitions=2, rg_batch=16, rg_sum=8,
                  out_classes=1, batch=130, nan_frac=0.0, binary=True),
    "k2o2_nan": dict(kind="gaussian", in_features=40, rg_depth=3, rg_repetitions=4, rg_batch=2, rg_sum=2,
                     out_classes=3, batch=200, nan_frac=0.3, optimize_scale=True),
}


@pytest.mark.parametrize("name", sorted(EINSUM_CASES))
def test_mma_einsum_matches_float64_oracle(name, monkeypatch):
    cfg = EINSUM_CASES[name]
    orc = oracle_for(cfg)[0].double()
    x, g = pg.ratspn_inputs(cfg)
    ref = orc.log_prob(x.double())
    outs = {}
    for knob in ("1", "0"):
        monkeypatch.setenv("DPK_EINSUM_MMA", knob)
        model = product_model(cfg, DEV, scale_grad=False)
        outs[knob] = model(x.to(DEV)).cpu()
        assert rel_err(outs[knob], ref) < 1e-4
    # both contractions agree far inside the north-star tolerance
    assert rel_err(outs["1"], outs["0"]) < 2e-5


def test_mma_einsum_peaked_weights_take_the_exact_path(monkeypatch):
    """Nearly one-hot mixture weights make the linear-domain sum underflow for most (i, j): exact fallback."""
    cfg = EINSUM_CASES["k8o8"]
    monkeypatch.setenv("DPK_EINSUM_MMA", "1")
    model = product_model(cfg, DEV, scale_grad=False)
    orc, state = oracle_for(cfg)
    with torch.no_grad():
        for layer in model._sum_layers():
            layer.weight.mul_(40.0)
    state = dict(state)
    for k, v in model.state_dict().items():
        if k.endswith(".weight") and k in state:
            state[k] = v.detach().cpu().clone()
    orc = orc.load_reference_state(state).double()
    x, _ = pg.ratspn_inputs(cfg)
    assert rel_err(model(x.to(DEV)).cpu(), orc.log_prob(x.double())) < 1e-4


def test_mma_einsum_gradients(monkeypatch):
    cfg = EINSUM_CASES["k10o10"]
    monkeypatch.setenv("DPK_EINSUM_MMA", "1")
    monkeypatch.setenv("DPK_LEAF_MMA", "1")
    model = product_model(cfg, DEV, scale_grad=False)
    x, g = pg.ratspn_inputs(cfg)
    with torch.enable_grad():
        xd = x.to(DEV).requires_grad_(True)
        out = model(xd)
        (out * g.to(DEV)).sum().backward()
    truth = oracle_for(cfg)[0].double().grads(x.double(), g.double(), clean_nan=True)
    tol = 1e-4 + 4e-7 * float(truth["out"].abs().max())
    assert norm_err(model.base_layer.loc.grad, truth["loc"].float()) < tol
    assert norm_err(model.root_layer.weight.grad, truth["root"].float()) < tol
    for a, b in zip([l.weight.grad for l in model._sum_layers()], truth["sums"]):
        assert norm_err(a, b.float()) < tol


# ---- leaf moments of the backward / E-step as a GEMM over the batch (ratspn_run_leaf_stats_mma) ----------------
@pytest.mark.parametrize("name", ["gauss784", "gauss36", "bern784", "gauss_scale"])
def test_mma_leaf_statistics_match_float64_oracle(name, monkeypatch):
    cfg = dict(CASES["gauss784"], optimize_scale=True, batch=520) if name == "gauss_scale" else dict(CASES[name])
    if cfg["batch"] % 4:
        cfg["batch"] += 4 - cfg["batch"] % 4
    orc = oracle_for(cfg)[0].double()
    x, g = pg.ratspn_inputs(cfg)
    ref = orc.em_statistics(x.double()) if cfg["out_classes"] == 1 else None
    res = {}
    for knob in ("1", "0"):
        monkeypatch.setenv("DPK_STATS_MMA", knob)
        model = product_model(cfg, DEV, scale_grad=(name == "gauss_scale"))
        if ref is not None:
            res[knob] = model.em_statistics(x.to(DEV))
        else:       # classes > 1: exercise the same kernels through the gradient path
            with torch.enable_grad():
                out = model(x.to(DEV))
                (out * g.to(DEV)).sum().backward()
            p = model.base_layer.loc if cfg["kind"] == "gaussian" else model.base_layer.logits
            res[knob] = {"grad": p.grad.clone()}
    if ref is not None:
        tol = 1e-4 + 4e-7 * float(res["1"]["ll"].abs().max())
        for key in ("s0", "s1", "s2"):
            if res["1"].get(key) is None:      # Bernoulli / frozen unit scale: no second moment
                continue
            assert norm_err(res["1"][key], ref[key]) < tol, key
            assert norm_err(res["1"][key], res["0"][key]) < tol, key
    else:
        assert norm_err(res["1"]["grad"], res["0"]["grad"]) < 2e-4


def test_mma_leaf_statistics_nan_inputs_fall_back(monkeypatch):
    cfg = dict(CASES["gauss784_nan"])
    monkeypatch.setenv("DPK_STATS_MMA", "1")
    model = product_model(cfg, DEV, scale_grad=False)
    orc = oracle_for(cfg)[0].double()
    x, _ = pg.ratspn_inputs(cfg)
    st = model.em_statistics(x.to(DEV))
    ref = orc.em_statistics(x.double())
    tol = 1e-4 + 4e-7 * float(st["ll"].abs().max())
    for key in ("s0", "s1", "s2"):
        if st.get(key) is not None:
            assert norm_err(st[key], ref[key]) < tol, key


# ---- d LL / d x of the leaf level as the transposed GEMM (ratspn_run_leaf_bwd_x_mma) -----------------------------
@pytest.mark.parametrize("name", ["gauss784", "gauss_scale", "bern784", "gauss784_nan"])
def test_mma_input_gradient_matches_float64_oracle(name, monkeypatch):
    """The gradient a flow with a RatSpn base needs (deeprob/flows/models/base.py:139): tensor-core path (selected
    automatically for batches >= 8192, forced here) against the float64 oracle and the CUDA-core kernel."""
    cfg = dict(CASES["gauss784"], optimize_scale=True, batch=520) if name == "gauss_scale" else dict(CASES[name])
    if cfg["batch"] % 4:
        cfg["batch"] += 4 - cfg["batch"] % 4
    orc = oracle_for(cfg)[0].double()
    x, g = pg.ratspn_inputs(cfg)
    ref = orc.grads(x.double(), g.double(), clean_nan=True)["x"]
    got = {}
    for knob in ("1", "0"):
        monkeypatch.setenv("DPK_STATS_MMA", knob)
        model = product_model(cfg, DEV, scale_grad=(name == "gauss_scale"))
        with torch.enable_grad():
            xd = x.to(DEV).requires_grad_(True)
            out = model(xd)
            (out * g.to(DEV)).sum().backward()
        got[knob] = torch.nan_to_num(xd.grad.cpu())
    tol = 1e-4 + 4e-7 * float(out.abs().max())
    assert norm_err(got["1"], torch.nan_to_num(ref)) < 2 * tol
    assert norm_err(got["1"], got["0"]) < 2 * tol
