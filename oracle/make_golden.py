"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run here (the CPU container):   python oracle/make_golden.py [ratspn|dgcspn|flows|all]
The fixtures hold only reference outputs (+ tiny index tables); parameters and inputs are
re-created from tests/param_gen.py by seed.  Test infrastructure only.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

from oracle import ref_env  # noqa: E402

ref_env.enable()
import param_gen as pg  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def make_ratspn():
    from deeprob.spn.models.ratspn import GaussianRatSpn, BernoulliRatSpn
    from deeprob.utils.region import RegionGraph

    # region-graph structure goldens (tests/test_ratspn.py:24-43 uses (15, 2, seed 42, 2 reps))
    rg = {}
    for (d, depth, reps, seed) in [(15, 2, 2, 42), (16, 2, 1, 42), (784, 3, 4, 42), (37, 3, 5, 7), (64, 5, 2, 0)]:
        layers = RegionGraph(d, depth, seed).make_layers(reps)
        leaf = layers[-1]
        width = max(len(r) for r in leaf)
        tab = -np.ones((len(leaf), width), np.int64)
        for i, r in enumerate(leaf):
            tab[i, :len(r)] = r
        rg["leaf_%d_%d_%d_%d" % (d, depth, reps, seed)] = tab
    np.savez_compressed(os.path.join(GOLDEN, "region_graph.npz"), **rg)

    for name, cfg in pg.RATSPN_CASES.items():
        cls = GaussianRatSpn if cfg["kind"] == "gaussian" else BernoulliRatSpn
        torch.manual_seed(0)
        model = cls(**pg.ratspn_ctor_kwargs(cfg)).eval()
        model.load_state_dict(pg.ratspn_fill_state(model.state_dict(), cfg))
        if cfg["kind"] == "gaussian":
            model.base_layer.scale.requires_grad_(True)     # golden d/dscale even when the ctor froze it
        x, g = pg.ratspn_inputs(cfg)
        xg = x.clone().requires_grad_(True)
        out = model(xg)
        (out * g).sum().backward()
        rec = {"ll": _np(out), "mask": _np(model.base_layer.mask).astype(np.int32)}
        small = cfg["in_features"] <= 64
        grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
        for k, v in grads.items():
            v = _np(v)
            rec["grad." + k] = v if (small or v.size <= 20000) else v.reshape(-1)[:: max(1, v.size // 4096)]
        gx = _np(torch.nan_to_num(xg.grad))
        rec["grad.x"] = gx if small else gx[:8]
        np.savez_compressed(os.path.join(GOLDEN, "ratspn_%s.npz" % name), **rec)
        print("ratspn", name, "ll[:3]=", rec["ll"].reshape(-1)[:3])


def make_mpe():
    """Reference RatSpn.mpe outputs (CPU; the reference builds its index tensors on the CPU)."""
    from deeprob.spn.models.ratspn import GaussianRatSpn, BernoulliRatSpn
    rec = {}
    for name, cfg in pg.MPE_CASES.items():
        cls = GaussianRatSpn if cfg["kind"] == "gaussian" else BernoulliRatSpn
        torch.manual_seed(0)
        model = cls(**pg.ratspn_ctor_kwargs(cfg)).eval()
        model.load_state_dict(pg.ratspn_fill_state(model.state_dict(), cfg))
        x, _ = pg.ratspn_inputs(cfg)
        with torch.no_grad():
            rec[name] = _np(model.mpe(x))
            if cfg["out_classes"] > 1:
                y = torch.arange(x.shape[0]) % cfg["out_classes"]
                rec[name + ".y"] = _np(model.mpe(x, y))
        print("mpe", name, rec[name].shape, "NaN left:", int(np.isnan(rec[name]).sum()))
    np.savez_compressed(os.path.join(GOLDEN, "ratspn_mpe.npz"), **rec)


def make_dgcspn():
    from deeprob.spn.models.dgcspn import DgcSpn
    for name, cfg in pg.DGCSPN_CASES.items():
        torch.manual_seed(0)
        model = DgcSpn(**pg.dgcspn_ctor_kwargs(cfg)).eval()
        names = [k for k, _ in model.named_parameters()]
        model.load_state_dict(pg.dgcspn_fill_state(model.state_dict(), names, cfg))
        model.base_layer.scale.requires_grad_(True)
        x, g = pg.dgcspn_inputs(cfg)
        xg = x.clone().requires_grad_(True)
        out = model(xg)
        (out * g).sum().backward()
        rec = {"ll": _np(out)}
        for k, p in model.named_parameters():
            if p.grad is not None:
                v = _np(p.grad)
                rec["grad." + k] = v if v.size <= 20000 else v.reshape(-1)[:: max(1, v.size // 4096)]
        rec["grad.x"] = _np(torch.nan_to_num(xg.grad))[:4]
        np.savez_compressed(os.path.join(GOLDEN, "dgcspn_%s.npz" % name), **rec)
        print("dgcspn", name, "ll[:3]=", rec["ll"].reshape(-1)[:3])


def make_flows(only=None):
    from deeprob.flows import models as ref_models
    import warnings
    warnings.simplefilter("ignore")
    for name, cfg in pg.FLOW_CASES.items():
        if only and name not in only:
            continue
        torch.manual_seed(0)
        model = getattr(ref_models, cfg["model"])(**cfg["kw"])
        model.load_state_dict(pg.flow_fill_state(model.state_dict(), fill_all=cfg.get("fill_all", False)))
        x, g = pg.flow_inputs(cfg)
        rec = {}
        if cfg["model"] == "RealNVP2d" and not cfg.get("fill_all", False):   # conv conditioners: keep the whole (small) state in the fixture
            for k, v in model.state_dict().items():
                rec["state." + k] = _np(v)
        model.eval()
        xg = x.clone().requires_grad_(True)
        out = model(xg)
        (out * g).sum().backward()
        rec["ll"] = _np(out)
        rec["grad.x"] = _np(xg.grad)[:4]
        for k, p in model.named_parameters():
            if p.grad is not None and ("scale_act" in k or p.numel() <= 4096):
                rec["grad." + k] = _np(p.grad)
        # invertibility vector (tests/test_flows.py:22-26): apply_forward(apply_backward(x)) == x
        with torch.no_grad():
            u, ildj = model.apply_backward(model.preprocess(x)[0])
            rec["u"] = _np(u)[:4]
            rec["ildj"] = _np(ildj if isinstance(ildj, torch.Tensor) else torch.full((x.shape[0],), float(ildj)))
        # training mode: batch statistics in the batch-norm bijectors (+ running-stat update)
        if any("running_var" in k for k in model.state_dict()) and cfg["model"] != "RealNVP2d":
            model.zero_grad()
            model.train()
            out_t = model(x)
            (out_t * g).sum().backward()
            rec["train.ll"] = _np(out_t)
            for k, v in model.state_dict().items():
                if k.endswith("running_mean") or k.endswith("running_var"):
                    rec["train.state." + k] = _np(v)
            for k, p in model.named_parameters():
                if p.grad is not None and ("scale_act" in k or p.numel() <= 4096):
                    rec["train.grad." + k] = _np(p.grad)
        np.savez_compressed(os.path.join(GOLDEN, "flows_%s.npz" % name), **rec)
        print("flows", name, "ll[:3]=", rec["ll"][:3])


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    os.makedirs(GOLDEN, exist_ok=True)
    if what in ("ratspn", "all"):
        make_ratspn()
    if what in ("mpe", "all"):
        make_mpe()
    if what in ("dgcspn", "all") and "make_dgcspn" in globals():
        globals()["make_dgcspn"]()
    if what in ("flows", "all") and "make_flows" in globals():
        globals()["make_flows"](sys.argv[2:] or None)
