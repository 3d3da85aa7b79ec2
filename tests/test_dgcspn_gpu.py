"""GPU parity of the CUDA DGC-SPN layers against the CPU oracle and the reference golden vectors."""
import pytest
import torch

import param_gen as pg
from conftest import load_golden, norm_err, rel_err
from helpers import dgc_oracle_for, dgc_product_model, subsample_like

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"


@pytest.mark.parametrize("name", sorted(pg.DGCSPN_CASES))
def test_log_prob_and_gradients(name):
    cfg = pg.DGCSPN_CASES[name]
    model = dgc_product_model(cfg, DEV)
    orc, _ = dgc_oracle_for(cfg)
    x, g = pg.dgcspn_inputs(cfg)
    gold = load_golden("dgcspn_" + name)
    with torch.enable_grad():
        xd = x.to(DEV).requires_grad_(True)
        out = model(xd)
        (out * g.to(DEV)).sum().backward()
    assert out.shape == (cfg["batch"], cfg["out_classes"])
    assert rel_err(out.detach().cpu(), orc.log_prob(x)) < TOL
    assert rel_err(out.detach().cpu(), gold["ll"]) < TOL
    # Gradients are posteriors = exp(differences of fp32 log-values of magnitude |LL|): the float64 oracle is the
    # ground truth and the bar is "within 1e-4 (+ the fp32 conditioning term) of it, or no worse than 3x the
    # error the reference's own fp32 op sequence (the fp32 oracle) makes against the same truth".
    truth = dgc_oracle_for(cfg)[0].double().grads(x.double(), g.double(), clean_nan=True)
    ref32 = orc.grads(x, g, clean_nan=True)
    gtol = TOL + 4e-7 * float(out.abs().max())
    sums = [l.weight.grad for l in model.layers if l.__class__.__name__ == "SpatialSumLayer"]

    def check(mine, key, idx=None):
        t = truth[key] if idx is None else truth[key][idx]
        r = ref32[key] if idx is None else ref32[key][idx]
        t, r = torch.nan_to_num(t), torch.nan_to_num(r)
        bar = max(gtol, 3.0 * norm_err(r, t))
        assert norm_err(torch.nan_to_num(mine), t) < bar, (key, idx, bar)

    check(model.base_layer.loc.grad, "loc")
    check(model.base_layer.scale.grad, "scale")
    check(model.root_layer.weight.grad, "root")
    for i, a in enumerate(sums):
        check(a, "sums", i)
    check(xd.grad.cpu(), "x")
    if cfg["nan_frac"] == 0:
        assert norm_err(subsample_like(model.base_layer.loc.grad.cpu(), gold["grad.base_layer.loc"].size),
                        gold["grad.base_layer.loc"].reshape(-1)) < 2 * gtol


def test_product_of_ones_and_shapes():
    """deeprob-kit tests/test_dgcspn.py:46-75: inside the valid region a 2x2 product of ones is 4."""
    from deeprob_kit_b200.spn.layers.dgcspn import SpatialProductLayer
    ones = torch.ones(8, 3, 32, 32, device=DEV)
    p = SpatialProductLayer((3, 32, 32), kernel_size=2, padding='full', stride=1, dilation=4, depthwise=True)
    assert torch.allclose(p(ones)[:, :, 4:-4, 4:-4], torch.tensor(4.0, device=DEV))
    p = SpatialProductLayer((3, 32, 32), kernel_size=2, padding='valid', stride=2, dilation=1, depthwise=True)
    out = p(ones)
    assert out.shape == (8, 3, 16, 16) and torch.allclose(out, torch.tensor(4.0, device=DEV))
    p = SpatialProductLayer((3, 32, 32), kernel_size=2, padding='full', stride=1, dilation=8, depthwise=False).to(DEV)
    out = p(ones)
    assert out.shape == (8, 81, 40, 40)
    assert torch.allclose(out[:, :, 8:-8, 8:-8], torch.tensor(4.0, device=DEV))


@pytest.mark.parametrize("n_pooling,depthwise", [(0, False), (2, False), (0, True), (2, True)])
def test_mpe_improves_log_prob(n_pooling, depthwise):
    """deeprob-kit tests/test_dgcspn.py:89-96."""
    from deeprob_kit_b200.spn.models import DgcSpn
    torch.manual_seed(42)
    data = torch.randn(8, 3, 32, 32)
    mar = data.clone()
    mar[torch.rand_like(mar) < 0.5] = float("nan")
    model = DgcSpn((3, 32, 32), n_batch=4, sum_channels=4, n_pooling=n_pooling, depthwise=depthwise).to(DEV)
    lls = model.log_prob(data.to(DEV))
    with torch.enable_grad():
        mpe = model.mpe(mar.to(DEV))
    assert torch.all(model.log_prob(mpe).squeeze() > lls.squeeze())


def test_config3_batch_properties():
    """BASELINE config 3 structure at a larger batch: slice/permutation invariance + marginalisation to 1."""
    cfg = dict(pg.DGCSPN_CASES["mnist"])
    model = dgc_product_model(cfg, DEV)
    orc, _ = dgc_oracle_for(cfg)
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(1030, 1, 28, 28, generator=gen)
    out = model(x.to(DEV))
    assert torch.equal(model(x[100:164].to(DEV)), out[100:164])
    idx = torch.arange(0, 1030, 41)
    assert rel_err(out[idx.to(DEV)].cpu(), orc.log_prob(x[idx])) < TOL
    assert float(model(torch.full((3, 1, 28, 28), float("nan"), device=DEV)).abs().max()) < 1e-3


@pytest.mark.parametrize("name", ["dw8", "mnist", "mnist_pool", "mixed16"])
def test_fused_product_sum_matches_layerwise(name, monkeypatch):
    """Inference fuses every depthwise product layer with the sum layer behind it (dpk_dgc_prodsum_forward):
    same values as the layer-by-layer path, NaN (marginalised) inputs included."""
    cfg = pg.DGCSPN_CASES[name]
    model = dgc_product_model(cfg, DEV)
    x = pg.dgcspn_inputs(cfg)[0].to(DEV)
    with torch.no_grad():      # the fused kernels carry no autograd node: they are used in inference only
        monkeypatch.setenv("DPK_DGC_FUSE", "1")
        fused = model(x)
        monkeypatch.setenv("DPK_DGC_FUSE", "0")
        plain = model(x)
    assert rel_err(fused, plain) < 2e-6
    assert rel_err(plain, model(x)) < 1e-7              # grad mode: layer by layer
    xn = x.clone()
    xn[::3, :, ::2, 1::3] = float("nan")
    with torch.no_grad():
        plain_n = model(xn)
        monkeypatch.setenv("DPK_DGC_FUSE", "1")
        fused_n = model(xn)
    assert rel_err(fused_n, plain_n) < 2e-6


def test_shape_mismatch_raises_instead_of_reading_out_of_bounds():
    """the kernels take C/H/W from the parameters; a different input shape must raise like the
    reference's broadcast / conv shape errors do."""
    cfg = pg.DGCSPN_CASES[sorted(pg.DGCSPN_CASES)[0]]
    model = dgc_product_model(cfg, DEV)
    c, h, w = cfg["in_features"]
    for shape in ((2, c + 1, h, w), (2, c, h + 1, w), (2, c, h, w - 1)):
        with pytest.raises(ValueError):
            model(torch.zeros(*shape, device=DEV))


@pytest.mark.parametrize("name", ["dw8", "mnist", "mixed16"])
def test_training_fusion_matches_layerwise(name, monkeypatch):
    """With gradients a depthwise product + sum pair is ONE autograd node (fused forward; the backward recomputes the
    product values from the taps, dpk_dgc_prodsum_backward, or through a temporary for more than 8 channels): values and
    every gradient equal the layer-by-layer autograd path (DPK_DGC_FUSE_TRAIN=0), NaN inputs included."""
    cfg = pg.DGCSPN_CASES[name]
    x, g = pg.dgcspn_inputs(cfg)
    x = x.clone()
    x[::4, :, 1::3, ::2] = float("nan")
    res = {}
    for mode in ("1", "0", "gather"):
        monkeypatch.setenv("DPK_DGC_FUSE_TRAIN", "0" if mode == "0" else "1")
        monkeypatch.setenv("DPK_DGC_BWD_GATHER", "1" if mode == "gather" else "0")
        model = dgc_product_model(cfg, DEV)
        with torch.enable_grad():
            xd = x.to(DEV).requires_grad_(True)
            out = model(xd)
            (out * g.to(DEV)).sum().backward()
        res[mode] = (out.detach(), torch.nan_to_num(xd.grad), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
    # two fp32 paths whose posteriors are exp(differences of log-values of magnitude |LL|): same bar as above
    gtol = TOL + 4e-7 * float(res["0"][0].abs().max())
    for mode in ("1", "gather"):
        assert rel_err(res[mode][0], res["0"][0]) < 2e-6
        assert norm_err(res[mode][1], res["0"][1]) < gtol
        assert res[mode][2].keys() == res["0"][2].keys()
        for n in res[mode][2]:
            assert norm_err(torch.nan_to_num(res[mode][2][n]), torch.nan_to_num(res["0"][2][n])) < gtol, (mode, n)
