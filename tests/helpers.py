"""Builders shared by CPU and GPU tests (oracle side and product side of every RAT-SPN case)."""
import numpy as np
import torch

import param_gen as pg
from oracle.ratspn_oracle import RatSpnOracle


def subsample_like(t, ref_size):
    """make_golden.py stores big gradient tensors strided; reproduce the same subsampling."""
    t = t.reshape(-1)
    if t.numel() == ref_size:
        return t
    return t[:: max(1, t.numel() // 4096)]


def oracle_for(cfg, seed=0):
    orc = RatSpnOracle(cfg["in_features"], cfg["kind"], cfg["rg_depth"], cfg["rg_repetitions"], cfg["rg_batch"],
                       cfg["rg_sum"], cfg["out_classes"], pg.RATSPN_SEED)
    state = oracle_state(orc, cfg, seed)
    return orc.load_reference_state(state), state


def oracle_state(orc, cfg, seed=0):
    """A reference-keyed state dict (shapes derived from the oracle's own structure)."""
    g0, k, dim = len(orc.leaf_regions), orc.K, orc.dim
    state = {"base_layer.mask": orc.mask}
    if cfg["kind"] == "gaussian":
        state["base_layer.loc"] = torch.zeros(g0, k, dim)
        state["base_layer.scale"] = torch.ones(g0, k, dim)
    else:
        state["base_layer.logits"] = torch.zeros(g0, k, dim)
    groups, nodes, idx = g0, k, 0
    for lvl in range(cfg["rg_depth"]):
        groups, nodes = groups // 2, nodes * nodes
        idx += 1  # product layer index
        if lvl < cfg["rg_depth"] - 1:
            state["layers.%d.weight" % idx] = torch.zeros(groups, orc.O, nodes)
            nodes = orc.O
            idx += 1
    state["root_layer.weight"] = torch.zeros(orc.C, groups * nodes)
    return pg.ratspn_fill_state(state, cfg, seed)


def product_model(cfg, device, seed=0, scale_grad=True):
    from deeprob_kit_b200.spn.models import BernoulliRatSpn, GaussianRatSpn
    cls = GaussianRatSpn if cfg["kind"] == "gaussian" else BernoulliRatSpn
    model = cls(**pg.ratspn_ctor_kwargs(cfg)).eval()
    state = pg.ratspn_fill_state(model.state_dict(), cfg, seed)
    model.load_state_dict(state)
    if cfg["kind"] == "gaussian" and scale_grad:
        model.base_layer.scale.requires_grad_(True)     # also exercises d/dscale; disables the unit-scale kernels
    return model.to(device)


# ---- DGC-SPN -------------------------------------------------------------------------------------
def dgc_oracle_for(cfg, seed=0):
    from oracle.dgcspn_oracle import DgcSpnOracle
    kw = pg.dgcspn_ctor_kwargs(cfg)
    kw.pop("optimize_scale")
    orc = DgcSpnOracle(**kw)
    k, (c, h, w) = cfg["n_batch"], cfg["in_features"]
    state = {"base_layer.loc": torch.zeros(k, c, h, w), "base_layer.scale": torch.ones(k, c, h, w)}
    for i, shp in enumerate(orc.sum_shapes):      # (C_out, C_in, H, W)
        state["layers.%d.weight" % (2 * i + 1)] = torch.zeros(*shp)
    state["root_layer.weight"] = torch.zeros(cfg["out_classes"], int(np.prod(orc.root_in)))
    state = pg.dgcspn_fill_state(state, list(state.keys()), cfg, seed)
    return orc.load_reference_state(state), state


def dgc_product_model(cfg, device, seed=0):
    from deeprob_kit_b200.spn.models import DgcSpn
    model = DgcSpn(**pg.dgcspn_ctor_kwargs(cfg)).eval()
    names = [k for k, _ in model.named_parameters()]
    model.load_state_dict(pg.dgcspn_fill_state(model.state_dict(), names, cfg, seed))
    model.base_layer.scale.requires_grad_(True)
    return model.to(device)


# ---- flows ------------------------------------------------------------------------------------------
def flow_reference_state(cfg, name=None):
    """Reference-keyed parameters of a flow case: 1D models are rebuilt from the product constructors (same
    keys/shapes/masks as the reference, checked in test_flows_host.py) + the numpy-seeded fill; 2D models
    read the state stored in their golden fixture (conv conditioners are initialised by torch's RNG)."""
    from conftest import load_golden
    from deeprob_kit_b200.flows import models as fm
    import warnings
    warnings.simplefilter("ignore")
    model = getattr(fm, cfg["model"])(**cfg["kw"])
    if cfg.get("fill_all", False):
        state = pg.flow_fill_state(model.state_dict(), fill_all=True)
    elif cfg["model"] == "RealNVP2d":
        gold = load_golden("flows_" + name)
        state = {k[len("state."):]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("state.")}
    else:
        state = pg.flow_fill_state(model.state_dict())
    model.load_state_dict(state)
    return model, state
