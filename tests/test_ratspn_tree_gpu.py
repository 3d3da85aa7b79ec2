"""GPU parity of the fused upper-level kernel (csrc/ratspn_tree_mma.cu: every product + sum level and the root in
one launch, mixtures on tcgen05 kind::tf32 with hi/lo operands) against the CPU oracle, the golden vectors and the
layer-wise CUDA path.  The kernel is selected automatically for inference batches >= 8192; DPK_TREE_MMA=1 forces it
for the small batches the oracle can check exhaustively."""
import os

import numpy as np
import pytest
import torch

import param_gen as pg
from conftest import load_golden, rel_err
from helpers import oracle_for, product_model

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"

EXTRA = {
    # classes > 1, depth 2
    "t_cls": dict(kind="gaussian", in_features=32, rg_depth=2, rg_repetitions=3, rg_batch=4, rg_sum=4, out_classes=3,
                  batch=300, nan_frac=0.1, optimize_scale=True),
    # the deepest tree the kernel walks (3 slots), classes 2
    "t_deep": dict(kind="gaussian", in_features=64, rg_depth=4, rg_repetitions=2, rg_batch=2, rg_sum=2, out_classes=2,
                   batch=200, nan_frac=0.0, optimize_scale=False),
    # depth 1: the root product sits directly on the leaves
    "t_d1": dict(kind="bernoulli", in_features=12, rg_depth=1, rg_repetitions=5, rg_batch=4, rg_sum=4, out_classes=2,
                 batch=130, nan_frac=0.0, binary=True),
    "t_k8": dict(kind="gaussian", in_features=40, rg_depth=3, rg_repetitions=3, rg_batch=8, rg_sum=8, out_classes=1,
                 batch=257, nan_frac=0.0, optimize_scale=False),
    # two K steps and 256 accumulator columns (two warpgroups per CTA)
    "t_k16": dict(kind="gaussian", in_features=24, rg_depth=2, rg_repetitions=2, rg_batch=16, rg_sum=16, out_classes=4,
                  batch=140, nan_frac=0.0, optimize_scale=True),
}
CASES = {**{k: pg.RATSPN_CASES[k] for k in ("gauss784", "bern16", "bern15", "bern15_nan")}, **EXTRA}


@pytest.fixture
def tree_on():
    prev = os.environ.get("DPK_TREE_MMA")
    os.environ["DPK_TREE_MMA"] = "1"
    yield
    if prev is None:
        del os.environ["DPK_TREE_MMA"]
    else:
        os.environ["DPK_TREE_MMA"] = prev


def layerwise(model, x):
    prev = os.environ.get("DPK_TREE_MMA")
    os.environ["DPK_TREE_MMA"] = "0"
    try:
        return model(x)
    finally:
        if prev is None:
            del os.environ["DPK_TREE_MMA"]
        else:
            os.environ["DPK_TREE_MMA"] = prev


@pytest.mark.parametrize("name", sorted(CASES))
def test_tree_matches_oracle(name, tree_on):
    cfg = CASES[name]
    model = product_model(cfg, DEV, scale_grad=False)
    orc, _ = oracle_for(cfg)
    x, _ = pg.ratspn_inputs(cfg)
    out = model(x.to(DEV))
    ref = orc.log_prob(x)
    assert out.shape == ref.shape
    assert rel_err(out.cpu(), ref) < TOL
    if name in pg.RATSPN_CASES:
        assert rel_err(out.cpu(), load_golden("ratspn_" + name)["ll"]) < TOL
    # the layer-wise CUDA path computes the same values (fp32 rounding apart)
    assert rel_err(out, layerwise(model, x.to(DEV))) < 2e-5
    # other batch sizes: partial tiles, one sample, several tiles per warpgroup
    rng = np.random.RandomState(3)
    for b in (1, 127, 129, 1500):
        xb = x[rng.randint(0, x.shape[0], size=b)]
        assert rel_err(model(xb.to(DEV)).cpu(), orc.log_prob(xb)) < TOL


def test_tree_extreme_weights_take_the_exact_path(tree_on):
    """Mixture weights down to e^-150: the linear-domain sums leave the range the fast path trusts and the samples
    are recomputed in the log domain (torch.logsumexp semantics)."""
    cfg = dict(EXTRA["t_cls"])
    model = product_model(cfg, DEV, scale_grad=False)
    orc, state = oracle_for(cfg)
    rng = np.random.RandomState(5)
    for k in [k for k in state if k.endswith(".weight")]:
        state[k] = state[k] + torch.from_numpy(rng.choice([0.0, -150.0, -60.0], size=state[k].shape)).float()
    orc.load_reference_state(state)
    model.load_state_dict({**model.state_dict(), **{k: v for k, v in state.items() if k.endswith(".weight")}})
    x, _ = pg.ratspn_inputs(cfg)
    assert rel_err(model(x.to(DEV)).cpu(), orc.log_prob(x)) < TOL


def test_tree_infinite_inputs(tree_on):
    cfg = dict(EXTRA["t_cls"])
    cfg["nan_frac"] = 0.0
    model = product_model(cfg, DEV, scale_grad=False)
    orc, _ = oracle_for(cfg)
    x, _ = pg.ratspn_inputs(cfg)
    x[3, 2] = float("inf")
    x[5, 0] = float("-inf")
    x[7, :] = float("inf")
    out, ref = model(x.to(DEV)).cpu(), orc.log_prob(x)
    finite = torch.isfinite(ref) & (ref > -1e30)
    assert rel_err(out[finite], ref[finite]) < TOL
    assert bool(((out < -1e30) == (ref < -1e30)).all())


def test_bench_batch_against_oracle():
    """The exact inputs of bench.py (65536 x 784, seed 1234, the auto-selected tcgen05 leaf + tree kernels):
    an oracle-checked strided subset, the full batch against the layer-wise CUDA path, and the size-independent
    properties (batch independence, permutation equivariance)."""
    from deeprob_kit_b200.spn.models import GaussianRatSpn
    torch.manual_seed(0)
    model = GaussianRatSpn(784, rg_depth=3, rg_repetitions=16, rg_batch=10, rg_sum=10, random_state=42).eval().to(DEV)
    x = torch.randn(65536, 784, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1234))
    out = model(x)
    from oracle.ratspn_oracle import RatSpnOracle
    orc = RatSpnOracle(784, "gaussian", 3, 16, 10, 10, 1, 42)
    orc.load_reference_state({k: v.detach().cpu() for k, v in model.state_dict().items()})
    idx = torch.arange(0, 65536, 331)
    assert rel_err(out[idx.to(DEV)].cpu(), orc.log_prob_chunked(x[idx.to(DEV)].cpu(), 64)) < TOL
    assert rel_err(out, layerwise(model, x)) < 2e-5
    part = model(x[20000:28192 + 20000])
    assert rel_err(part, out[20000:28192 + 20000]) < 1e-6
    perm = torch.randperm(65536, device=DEV)
    assert rel_err(model(x[perm]), out[perm]) < 1e-6


@pytest.mark.parametrize("name", ["t_k8", "gauss784"])
def test_unaligned_input_takes_the_exact_leaf_path(name, tree_on, monkeypatch):
    """x whose storage starts 4 bytes off a 16-byte boundary cannot feed the tensor-core leaf kernels (16-byte loads,
    TMA): the exact kernel computes the whole leaf level -- quadratic term included, so the root must not add the
    per-sample -x^2/2 again -- and the fused tree kernel runs behind it.  Same values as for the aligned copy."""
    monkeypatch.setenv("DPK_LEAF_MMA", "1")
    cfg = CASES[name]
    model = product_model(cfg, DEV, scale_grad=False)
    x, _ = pg.ratspn_inputs(cfg)
    flat = torch.empty(x.numel() + 1, device=DEV)
    xu = flat[1:].view_as(x)
    xu.copy_(x)
    assert xu.data_ptr() % 16 == 4 and xu.is_contiguous()
    out_u = model(xu)
    out_a = model(x.to(DEV))
    assert rel_err(out_u, out_a) < 2e-6
    for leaf_stream in ("1",):
        monkeypatch.setenv("DPK_LEAF_STREAM", leaf_stream)
        if cfg["rg_repetitions"] * cfg["rg_batch"] * (1 << cfg["rg_depth"]) <= 256:
            assert rel_err(model(xu), out_a) < 2e-6
