"""CPU: the oracle restatement is pinned against vectors produced by the live reference."""
import numpy as np
import pytest
import torch

import param_gen as pg
from conftest import load_golden, norm_err, rel_err
from helpers import oracle_for, subsample_like
from oracle import ref_env
from oracle.ratspn_oracle import region_graph_layers

TOL = 1e-5


@pytest.mark.parametrize("name", sorted(pg.RATSPN_CASES))
def test_ratspn_oracle_matches_reference_golden(name):
    cfg = pg.RATSPN_CASES[name]
    gold = load_golden("ratspn_" + name)
    orc, _ = oracle_for(cfg)
    assert np.array_equal(orc.mask.numpy().astype(np.int32), gold["mask"])   # structure is bit-exact
    x, g = pg.ratspn_inputs(cfg)
    res = orc.grads(x, g)
    assert rel_err(res["out"], gold["ll"]) < TOL
    key_map = {"grad.base_layer.loc": "loc", "grad.base_layer.scale": "scale", "grad.base_layer.logits": "logits",
               "grad.root_layer.weight": "root"}
    sum_keys = sorted((k for k in gold if k.startswith("grad.layers.")), key=lambda k: int(k.split(".")[2]))
    for k, ref in gold.items():
        if not k.startswith("grad."):
            continue
        if k == "grad.x":
            mine = torch.nan_to_num(res["x"])[: ref.shape[0]]
        elif k in key_map:
            mine = res[key_map[k]]
        else:
            mine = res["sums"][sum_keys.index(k)]
        mine = subsample_like(mine, ref.size) if k != "grad.x" else mine
        assert norm_err(mine.reshape(-1), ref.reshape(-1)) < 1e-4, k


def test_region_graph_oracle_matches_reference_tables():
    gold = load_golden("region_graph")
    for key, tab in gold.items():
        _, d, depth, reps, seed = key.split("_")
        leaf = region_graph_layers(int(d), int(depth), int(reps), int(seed))[-1]
        assert len(leaf) == tab.shape[0]
        for reg, row in zip(leaf, tab):
            assert tuple(row[row >= 0]) == reg


@pytest.mark.skipif(not ref_env.available(), reason="reference tree not mounted")
def test_oracle_against_live_reference():
    """Where /root/reference is present, A/B the oracle against the reference itself (random init)."""
    ref_env.enable()
    from deeprob.spn.models.ratspn import GaussianRatSpn
    torch.manual_seed(3)
    cfg = dict(kind="gaussian", in_features=50, rg_depth=3, rg_repetitions=3, rg_batch=4, rg_sum=5, out_classes=3)
    ref = GaussianRatSpn(50, out_classes=3, rg_depth=3, rg_repetitions=3, rg_batch=4, rg_sum=5, random_state=42,
                         optimize_scale=True).eval()
    from oracle.ratspn_oracle import RatSpnOracle
    orc = RatSpnOracle(50, "gaussian", 3, 3, 4, 5, 3, 42).load_reference_state(ref.state_dict())
    x = torch.randn(40, 50)
    x[::3, ::4] = float("nan")
    assert rel_err(orc.log_prob(x), ref(x)) < 1e-6
    del cfg


def test_oracle_is_a_normalised_distribution_over_all_binary_states():
    """The reference's own known-answer property for this path (tests/test_ratspn.py:46-48): a Bernoulli RAT-SPN
    over 15 binary variables sums to one over the 2^15 complete states; with marginalised variables (NaN,
    tests/utils.py random_marginalize_data) the remaining ones still do."""
    cfg = pg.RATSPN_CASES["bern15"]
    orc, _ = oracle_for(cfg)
    states = ((torch.arange(2 ** 15).unsqueeze(1) >> torch.arange(14, -1, -1)) & 1).float()
    ll = orc.log_prob(states).double().reshape(-1)
    assert abs(float(torch.logsumexp(ll, 0))) < 1e-4
    # marginalise the last 5 variables: 2^10 states of the first 10, each listed once
    sub = states[:: 2 ** 5].clone()
    sub[:, 10:] = float("nan")
    llm = orc.log_prob(sub).double().reshape(-1)
    assert abs(float(torch.logsumexp(llm, 0))) < 1e-4
