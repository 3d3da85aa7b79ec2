"""Deterministic (numpy-seeded) parameters and inputs shared by the golden generator and the tests.

Keeping parameters a pure function of (config, seed) lets the golden fixtures store only the
reference OUTPUTS; the same state is loaded into the reference (oracle/make_golden.py), into the
CPU oracle and into the CUDA-backed modules through the reference's `state_dict` keys.
"""
import numpy as np
import torch

# name -> constructor config of a RAT-SPN case (keys follow deeprob/spn/models/ratspn.py:17-30,195-208)
RATSPN_CASES = {
    # BASELINE config 1 as written: Bernoulli D=16 depth 2 K=2 batch 128
    "bern16": dict(kind="bernoulli", in_features=16, rg_depth=2, rg_repetitions=1, rg_batch=2, rg_sum=2,
                   out_classes=1, batch=128, nan_frac=0.0, binary=True),
    # the shape tests/test_ratspn.py:18-21 really uses (pad>0: 15 % 8 != 0)
    "bern15": dict(kind="bernoulli", in_features=15, rg_depth=3, rg_repetitions=4, rg_batch=4, rg_sum=2,
                   out_classes=1, batch=256, nan_frac=0.0, binary=True),
    "bern15_nan": dict(kind="bernoulli", in_features=15, rg_depth=3, rg_repetitions=4, rg_batch=4, rg_sum=2,
                       out_classes=1, batch=256, nan_frac=0.5, binary=True),
    # depth 1: no inner sum layer at all (Product -> Root)
    "gauss_d1": dict(kind="gaussian", in_features=9, rg_depth=1, rg_repetitions=3, rg_batch=5, rg_sum=3,
                     out_classes=1, batch=64, nan_frac=0.0, optimize_scale=True),
    # classes > 1, odd sizes, padding, learnable scale, NaNs
    "gauss_cls": dict(kind="gaussian", in_features=37, rg_depth=3, rg_repetitions=5, rg_batch=6, rg_sum=7,
                      out_classes=4, batch=96, nan_frac=0.2, optimize_scale=True),
    "gauss_deep": dict(kind="gaussian", in_features=64, rg_depth=5, rg_repetitions=2, rg_batch=3, rg_sum=4,
                       out_classes=2, batch=80, nan_frac=0.0, optimize_scale=False),
    # north-star structure (BASELINE config 2) on a small batch
    "gauss784": dict(kind="gaussian", in_features=784, rg_depth=3, rg_repetitions=16, rg_batch=10, rg_sum=10,
                     out_classes=1, batch=192, nan_frac=0.0, optimize_scale=False),
    "gauss784_scale_nan": dict(kind="gaussian", in_features=784, rg_depth=3, rg_repetitions=16, rg_batch=10,
                               rg_sum=10, out_classes=1, batch=64, nan_frac=0.1, optimize_scale=True),
    "bern784": dict(kind="bernoulli", in_features=784, rg_depth=4, rg_repetitions=8, rg_batch=16, rg_sum=12,
                    out_classes=10, batch=64, nan_frac=0.0, binary=True),
}
RATSPN_SEED = 42          # region-graph random_state used by every case

# MPE completion cases (reference RatSpn.mpe, models/ratspn.py:124-160).  All have pad == 0 (D % 2^depth == 0):
# with padding the reference's unpad step keeps the padded entries (layers/ratspn.py:82-84) and is no ground truth.
MPE_CASES = {
    "mpe_gauss": dict(kind="gaussian", in_features=16, rg_depth=2, rg_repetitions=3, rg_batch=3, rg_sum=2,
                      out_classes=3, batch=40, nan_frac=0.4, optimize_scale=True),
    "mpe_bern": dict(kind="bernoulli", in_features=32, rg_depth=3, rg_repetitions=2, rg_batch=4, rg_sum=3,
                     out_classes=1, batch=33, nan_frac=0.5, binary=True),
    "mpe_d1": dict(kind="gaussian", in_features=8, rg_depth=1, rg_repetitions=4, rg_batch=3, rg_sum=2,
                   out_classes=2, batch=20, nan_frac=0.3, optimize_scale=False),
    "mpe_784": dict(kind="gaussian", in_features=784, rg_depth=3, rg_repetitions=16, rg_batch=10, rg_sum=10,
                    out_classes=1, batch=24, nan_frac=0.5, optimize_scale=False),
}


def ratspn_ctor_kwargs(cfg):
    kw = {k: cfg[k] for k in ("in_features", "rg_depth", "rg_repetitions", "rg_batch", "rg_sum", "out_classes")}
    kw["random_state"] = RATSPN_SEED
    if cfg["kind"] == "gaussian":
        kw["optimize_scale"] = cfg.get("optimize_scale", False)
    return kw


def _log_dirichlet(rng, shape):
    g = rng.gamma(1.0, 1.0, size=shape).astype(np.float64) + 1e-12
    return np.log(g / g.sum(-1, keepdims=True)).astype(np.float32)


def ratspn_fill_state(state, cfg, seed=0):
    """Overwrite every learnable tensor of a reference-keyed state_dict in place (sorted-key order)."""
    rng = np.random.RandomState(1000 + seed)
    out = {}
    for key in sorted(state.keys()):
        t = state[key]
        if key == "base_layer.loc" or key == "base_layer.logits":
            v = rng.standard_normal(t.shape).astype(np.float32)
        elif key == "base_layer.scale":
            if cfg.get("optimize_scale", False):
                v = (0.35 + 0.9 * rng.random_sample(t.shape)).astype(np.float32)
            else:
                v = np.ones(t.shape, np.float32)
        elif key.endswith(".weight"):
            v = _log_dirichlet(rng, tuple(t.shape)) + rng.standard_normal((*t.shape[:-1], 1)).astype(np.float32)
        else:
            out[key] = t
            continue
        out[key] = torch.from_numpy(v)
    return out


def ratspn_inputs(cfg, seed=0):
    rng = np.random.RandomState(2000 + seed)
    b, d = cfg["batch"], cfg["in_features"]
    if cfg.get("binary", False):
        x = (rng.random_sample((b, d)) < 0.5).astype(np.float32)
    else:
        x = rng.standard_normal((b, d)).astype(np.float32)
    if cfg.get("nan_frac", 0.0) > 0:
        x[rng.random_sample((b, d)) < cfg["nan_frac"]] = np.nan
        x[0, :] = np.nan              # one fully marginalised row: LL must be ~0
    g = rng.standard_normal((b, cfg["out_classes"])).astype(np.float32)
    return torch.from_numpy(x), torch.from_numpy(g)


# ------------------------------------------------------------------------------------------------
# DGC-SPN cases (constructor kwargs of deeprob/spn/models/dgcspn.py:16-29)
# ------------------------------------------------------------------------------------------------
DGCSPN_CASES = {
    "dw8": dict(in_features=(1, 8, 8), out_classes=1, n_batch=4, sum_channels=4, depthwise=True, n_pooling=0,
                optimize_scale=False, batch=24, nan_frac=0.0),
    "full8": dict(in_features=(2, 8, 8), out_classes=3, n_batch=2, sum_channels=2, depthwise=False, n_pooling=1,
                  optimize_scale=True, batch=20, nan_frac=0.25),
    "mixed16": dict(in_features=(3, 16, 16), out_classes=2, n_batch=4, sum_channels=5, depthwise=[False, True],
                    n_pooling=2, optimize_scale=True, batch=12, nan_frac=0.0),
    # BASELINE config 3 structure on a small batch
    "mnist": dict(in_features=(1, 28, 28), out_classes=1, n_batch=8, sum_channels=8, depthwise=True, n_pooling=0,
                  optimize_scale=False, batch=16, nan_frac=0.0),
    "mnist_pool": dict(in_features=(1, 28, 28), out_classes=10, n_batch=16, sum_channels=32, depthwise=True, n_pooling=2,
                       optimize_scale=True, batch=8, nan_frac=0.1),
}


def dgcspn_ctor_kwargs(cfg):
    kw = {k: cfg[k] for k in ("in_features", "out_classes", "n_batch", "sum_channels", "n_pooling", "optimize_scale")}
    dw = cfg["depthwise"]
    kw["depthwise"] = list(dw) if isinstance(dw, (list, tuple)) else dw
    return kw


def dgcspn_fill_state(state, param_names, cfg, seed=0):
    """Overwrite the learnable tensors (keys in `param_names`) of a reference-keyed state_dict."""
    rng = np.random.RandomState(3000 + seed)
    out = dict(state)
    for key in sorted(param_names):
        shape = tuple(state[key].shape)
        if key == "base_layer.loc":
            v = rng.standard_normal(shape).astype(np.float32)
        elif key == "base_layer.scale":
            v = (0.4 + 0.8 * rng.random_sample(shape)).astype(np.float32) if cfg["optimize_scale"] else np.ones(shape, np.float32)
        elif key == "root_layer.weight":
            v = _log_dirichlet(rng, shape) + np.float32(0.3)
        else:   # spatial sum layer (C_out, C_in, H, W): Dirichlet over dim 1 + a per-(o,h,w) shift
            g = rng.gamma(1.0, 1.0, size=shape) + 1e-12
            v = (np.log(g / g.sum(1, keepdims=True)) + rng.standard_normal((shape[0], 1, *shape[2:]))).astype(np.float32)
        out[key] = torch.from_numpy(v)
    return out


def dgcspn_inputs(cfg, seed=0):
    rng = np.random.RandomState(4000 + seed)
    shape = (cfg["batch"], *cfg["in_features"])
    x = rng.standard_normal(shape).astype(np.float32)
    if cfg["nan_frac"] > 0:
        x[rng.random_sample(shape) < cfg["nan_frac"]] = np.nan
        x[0] = np.nan
    g = rng.standard_normal((cfg["batch"], cfg["out_classes"])).astype(np.float32)
    return torch.from_numpy(x), torch.from_numpy(g)


# ------------------------------------------------------------------------------------------------
# Flow cases (constructor kwargs of deeprob/flows/models/realnvp.py:17-28,76-87 and maf.py:13-26)
# ------------------------------------------------------------------------------------------------
FLOW_CASES = {
    "nvp1d_small": dict(model="RealNVP1d", kw=dict(in_features=10, n_flows=3, depth=2, units=16), batch=48, unit_x=False),
    "nvp1d_nice": dict(model="RealNVP1d", kw=dict(in_features=9, n_flows=4, depth=1, units=12, batch_norm=False, affine=False),
                       batch=40, unit_x=False),
    "nvp1d_logit": dict(model="RealNVP1d", kw=dict(in_features=12, n_flows=2, depth=1, units=16, logit=0.05), batch=32,
                        unit_x=True),
    # BASELINE config 4 (primary): 8 affine couplings with MLP conditioners on flattened 32x32x3
    "nvp1d_cifar": dict(model="RealNVP1d", kw=dict(in_features=3072, n_flows=8, depth=2, units=512), batch=24, unit_x=True),
    "maf_seq": dict(model="MAF", kw=dict(in_features=10, n_flows=2, depth=2, units=16), batch=48, unit_x=False),
    "maf_rand": dict(model="MAF", kw=dict(in_features=7, n_flows=3, depth=1, units=20, activation="tanh", sequential=False,
                                          random_state=42, batch_norm=False), batch=32, unit_x=False),
    "nvp2d_res": dict(model="RealNVP2d", kw=dict(in_features=(3, 8, 8), n_flows=1, n_blocks=1, channels=4), batch=12,
                      unit_x=False),
    "nvp2d_dense": dict(model="RealNVP2d", kw=dict(in_features=(2, 8, 8), network="densenet", n_flows=1, n_blocks=1, channels=4,
                                                   logit=0.1), batch=10, unit_x=True),
    # BASELINE config 4 (secondary, examples/nvp2d_cifar10.py:30-38): 32x32x3, 10 couplings with residual conv
    # conditioners of 4 blocks x 64 channels.  10 M parameters: every tensor is a function of the NumPy seed
    # (fill_all) instead of being stored in the fixture.
    "nvp2d_cifar": dict(model="RealNVP2d", kw=dict(in_features=(3, 32, 32), n_flows=1, n_blocks=4, channels=64), batch=4,
                        unit_x=False, fill_all=True),
}


def flow_fill_state(state, seed=0, fill_all=False):
    """Perturb a freshly initialised flow state_dict so that every term is exercised: ScaledTanh weights are
    zero-initialised (log-det == 0 at init, deeprob/torch/utils.py:61) and the batch-norm statistics are trivial."""
    rng = np.random.RandomState(5000 + seed)
    out = {}
    for key in sorted(state.keys()):
        t = state[key]
        shape = tuple(t.shape)
        name = key.split(".")[-1]
        if "scale_act" in key:
            v = 0.2 + 0.8 * rng.random_sample(shape)
        elif name == "running_var":
            v = 0.5 + rng.random_sample(shape)
        elif name == "running_mean":
            v = 0.3 * rng.standard_normal(shape)
        elif name in ("weight", "bias") and t.dim() in (2, 4) and shape[0] == 1:      # batch-norm bijector (1,F[,1,1])
            v = 0.2 * rng.standard_normal(shape)
        elif name == "weight" and t.dim() == 2 and "network" in key:                   # Linear / MaskedLinear
            v = rng.standard_normal(shape) * (0.7 / np.sqrt(shape[1]))
        elif name == "bias" and "network" in key and t.dim() == 1 and "conv" not in key:
            v = 0.1 * rng.standard_normal(shape)
        elif fill_all and "perm_matrices" in key:                                      # fixed 0/1 multi-scale permutations
            out[key] = t
            continue
        elif fill_all and t.is_floating_point() and t.dim() == 4:                      # conv kernels (weight_v | weight)
            v = rng.standard_normal(shape) * ((0.5 if name == "weight_g" else 1.0) / np.sqrt(max(1, int(np.prod(shape[1:])))))
            if name == "weight_g":
                v = 0.15 + 0.15 * rng.random_sample(shape)
        elif fill_all and t.is_floating_point() and t.dim() == 1 and name == "weight":  # BatchNorm2d gain
            v = 0.5 + rng.random_sample(shape)
        elif fill_all and t.is_floating_point() and t.dim() == 1 and name == "bias":
            v = 0.1 * rng.standard_normal(shape)
        else:
            out[key] = t
            continue
        out[key] = torch.from_numpy(np.asarray(v, dtype=np.float32)).reshape(shape)
    return out


def flow_inputs(cfg, seed=0):
    rng = np.random.RandomState(6000 + seed)
    feats = cfg["kw"]["in_features"]
    shape = (cfg["batch"],) + (tuple(feats) if isinstance(feats, tuple) else (feats,))
    x = rng.random_sample(shape) if cfg["unit_x"] else rng.standard_normal(shape)
    g = rng.standard_normal((cfg["batch"],))
    return torch.from_numpy(x.astype(np.float32)), torch.from_numpy(g.astype(np.float32))
