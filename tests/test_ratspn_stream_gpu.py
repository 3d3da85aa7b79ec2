"""GPU parity of the streaming leaf kernel for narrow models (csrc/ratspn_leaf_stream.cu: x read once through TMA,
hi/lo fp16 split in shared memory, tcgen05 kind::f16) against the CPU oracle and the exact CUDA-core path.  Selected
automatically for inference batches >= 8192 with R*K < 24 and at most 256 leaf columns; DPK_LEAF_STREAM=1 forces it
for the small batches the oracle checks."""
import os

import numpy as np
import pytest
import torch

import param_gen as pg
from conftest import rel_err
from helpers import oracle_for, product_model

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"

CASES = {
    # D = 40: two K blocks, the second one 8 features wide (TMA zero fill); 24 * 8 = 192 columns -> 256-column tile
    "s_k8": dict(kind="gaussian", in_features=40, rg_depth=3, rg_repetitions=3, rg_batch=8, rg_sum=8, out_classes=1,
                 batch=257, nan_frac=0.0, optimize_scale=False),
    # one K block, 16 columns, classes
    "s_small": dict(kind="bernoulli", in_features=12, rg_depth=1, rg_repetitions=2, rg_batch=4, rg_sum=4, out_classes=2,
                    batch=130, nan_frac=0.0, binary=True),
    # odd number of K blocks (5): the two converter groups swap parity every tile
    "s_odd": dict(kind="gaussian", in_features=132, rg_depth=2, rg_repetitions=2, rg_batch=4, rg_sum=4, out_classes=3,
                  batch=300, nan_frac=0.0, optimize_scale=False),
    # NaN evidence: flagged groups are redone by the exact kernel
    "s_nan": dict(kind="gaussian", in_features=64, rg_depth=2, rg_repetitions=4, rg_batch=8, rg_sum=8, out_classes=1,
                  batch=400, nan_frac=0.05, optimize_scale=False),
}


@pytest.fixture
def stream_on():
    keys = {"DPK_LEAF_STREAM": "1", "DPK_TREE_MMA": "1"}
    prev = {k: os.environ.get(k) for k in keys}
    os.environ.update(keys)
    yield
    for k, v in prev.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def exact(model, x):
    prev = {k: os.environ.get(k) for k in ("DPK_LEAF_STREAM", "DPK_TREE_MMA", "DPK_LEAF_MMA")}
    os.environ.update({"DPK_LEAF_STREAM": "0", "DPK_TREE_MMA": "0", "DPK_LEAF_MMA": "0"})
    try:
        return model(x)
    finally:
        for k, v in prev.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("mt,grid", [("64", "148"), ("128", "148"), ("112", "2"), ("16", "3")])
@pytest.mark.parametrize("name", sorted(CASES))
def test_stream_matches_oracle(name, mt, grid, stream_on):
    """Tile heights 64 / 128 / 112 / 16 rows; grids of 2 and 3 CTAs give every CTA many tiles (both accumulator
    slots, the stage ring wrapping inside and across tiles)."""
    os.environ["DPK_STREAM_MT"] = mt
    os.environ["DPK_STREAM_GRID"] = grid
    try:
        cfg = CASES[name]
        model = product_model(cfg, DEV, scale_grad=False)
        orc, _ = oracle_for(cfg)
        x, _ = pg.ratspn_inputs(cfg)
        out = model(x.to(DEV))
        ref = orc.log_prob(x)
        assert out.shape == ref.shape
        assert rel_err(out.cpu(), ref) < TOL
        assert rel_err(out, exact(model, x.to(DEV))) < 2e-5
        rng = np.random.RandomState(3)
        for b in (1, 63, 65, 129, 1500):   # partial tiles, several tiles per CTA pipeline slot
            xb = x[rng.randint(0, x.shape[0], size=b)]
            assert rel_err(model(xb.to(DEV)).cpu(), orc.log_prob(xb)) < TOL
    finally:
        del os.environ["DPK_STREAM_MT"]
        del os.environ["DPK_STREAM_GRID"]


def test_stream_out_of_range_inputs(stream_on):
    """|x| beyond the fp16 hi/lo range and infinities: the sample groups are flagged and redone exactly."""
    cfg = dict(CASES["s_k8"])
    model = product_model(cfg, DEV, scale_grad=False)
    orc, _ = oracle_for(cfg)
    x, _ = pg.ratspn_inputs(cfg)
    x[3, 2] = 1e6
    x[70, 39] = -4e4
    x[200, 0] = float("inf")
    out, ref = model(x.to(DEV)).cpu(), orc.log_prob(x)
    finite = torch.isfinite(ref) & (ref > -1e30)
    assert rel_err(out[finite], ref[finite]) < TOL
    assert bool(((out < -1e30) == (ref < -1e30)).all())


def test_stream_auto_selected_full_batch():
    """bench.py's hbm_bound_check workload (65536 x 784, R=1, K=8): the kernel is picked without any knob; an
    oracle-checked strided subset, the whole batch against the exact CUDA path, and many tiles per CTA."""
    from deeprob_kit_b200.spn.models import GaussianRatSpn
    from oracle.ratspn_oracle import RatSpnOracle
    torch.manual_seed(0)
    model = GaussianRatSpn(784, rg_depth=3, rg_repetitions=1, rg_batch=8, rg_sum=8, random_state=42).eval().to(DEV)
    x = torch.randn(65536, 784, device=DEV, generator=torch.Generator(device=DEV).manual_seed(7))
    out = model(x)
    orc = RatSpnOracle(784, "gaussian", 3, 1, 8, 8, 1, 42)
    orc.load_reference_state({k: v.detach().cpu() for k, v in model.state_dict().items()})
    idx = torch.arange(0, 65536, 257)
    assert rel_err(out[idx.to(DEV)].cpu(), orc.log_prob_chunked(x[idx.to(DEV)].cpu(), 64)) < TOL
    assert rel_err(out, exact(model, x)) < 2e-5
    perm = torch.randperm(65536, device=DEV)
    assert rel_err(model(x[perm]), out[perm]) < 1e-6
