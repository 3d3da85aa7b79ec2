"""GPU parity of dpk_linear_forward (the conditioner MLP of the 1-D couplings on tcgen05, csrc/ratspn_leaf_mma.cu)
against float64 matmul, and of RealNVP1d evaluated through it against the cuBLAS path.
Tolerance: the 22-bit hi/lo operands are exact to 2e-7, but the tensor core adds into its fp32 accumulator with
truncation, once per 16-element K step and pass, so the error grows ~ K/16 * 3 * 2^-24 relative to the running
sum (3e-5 at K = 3072); still far inside the 1e-4 of the north star."""
import pytest
import torch

import numpy as np

from conftest import norm_err, rel_err
from deeprob_kit_b200.flows import _engine

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("batch,k,n,relu", [(300, 64, 40, True), (1000, 3072, 512, True), (513, 512, 6144, False),
                                              (128, 36, 7, False), (2049, 100, 300, True)])
def test_linear_matches_float64(batch, k, n, relu):
    g = torch.Generator(device=DEV).manual_seed(batch + k)
    x = torch.randn(batch, k, device=DEV, generator=g) * 3.0
    w = torch.randn(n, k, device=DEV, generator=g) / k ** 0.5
    b = torch.randn(n, device=DEV, generator=g)
    cache = {}
    out = _engine.linear(x, w, b, relu, cache)
    ref = x.double() @ w.double().t() + b.double()
    if relu:
        ref = ref.clamp_min(0.0)
    scale = float(ref.abs().max())
    tol = 2e-6 + 3.0 * (k / 16) * 2.0 ** -24
    assert float((out.double() - ref).abs().max()) / scale < tol
    # weight images are cached on (pointer, version): same result, and an in-place update is picked up
    assert torch.equal(_engine.linear(x, w, b, relu, cache), out)
    w.mul_(2.0)
    out2 = _engine.linear(x, w, None, relu, cache)
    ref2 = x.double() @ w.double().t()
    if relu:
        ref2 = ref2.clamp_min(0.0)
    assert float((out2.double() - ref2).abs().max()) / float(ref2.abs().max()) < tol


def test_linear_out_of_range_rows_are_exact():
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(400, 128, device=DEV, generator=g)
    x[7, 3] = float("nan")
    x[40, :] *= 1.0e6
    x[399, 100] = float("inf")
    w = torch.randn(96, 128, device=DEV, generator=g)
    b = torch.randn(96, device=DEV, generator=g)
    out = _engine.linear(x, w, b, True, {})
    ref = torch.relu(torch.nn.functional.linear(x, w, b))
    assert torch.isnan(out[7]).all() and torch.isnan(ref[7]).all()
    ok = torch.ones(400, dtype=torch.bool, device=DEV)
    ok[7] = False
    ok[399] = False
    ok[40] = False      # huge row: exact fp32 path, compared relative to the row's magnitude (summation order differs)
    assert rel_err(out[ok], ref[ok]) < 5e-5
    assert float((out[40] - ref[40]).abs().max() / ref[40].abs().max()) < 1e-5
    assert torch.equal(torch.isnan(out[399]), torch.isnan(ref[399]))
    fin = ~torch.isnan(ref[399])
    assert torch.equal(out[399][fin], ref[399][fin])


def _perturbed_nvp(d, batch_norm, affine, seed=3):
    from deeprob_kit_b200.flows.models import RealNVP1d
    from deeprob_kit_b200.flows.utils import BatchNormLayer1d
    torch.manual_seed(seed)
    model = RealNVP1d(d, n_flows=4, depth=2, units=128, batch_norm=batch_norm, affine=affine).to(DEV).eval()
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn_like(p))
        for m in model.modules():
            if isinstance(m, BatchNormLayer1d):     # away from the identity, else the fused affine is untested
                m.running_mean.add_(0.3 * torch.randn_like(m.running_mean))
                m.running_var.mul_(torch.exp(0.3 * torch.randn_like(m.running_var)))
    return model


def test_realnvp1d_through_the_tensor_core_mlp(monkeypatch):
    model = _perturbed_nvp(256, True, True)
    x = torch.rand(4096, 256, device=DEV)
    monkeypatch.setenv("DPK_FLOW_COMPACT", "0")         # full-width conditioner with the mask folded into layer 1
    with torch.no_grad():
        monkeypatch.setenv("DPK_LINEAR_MMA", "1")
        ll1 = model(x)
        monkeypatch.setenv("DPK_LINEAR_MMA", "0")
        ll0 = model(x)
    assert rel_err(ll1, ll0) < 1e-5
    assert rel_err(ll0, model(x)) < 1e-6                # grad mode: stock modules


@pytest.mark.parametrize("d,batch_norm,affine", [(256, True, True), (24, True, True), (36, True, True),
                                                   (256, False, True), (64, True, False)])
def test_realnvp1d_compact_inference(monkeypatch, d, batch_norm, affine):
    """Inference fast path (live conditioner columns only, compact z, fused eval batch-norm) against the stock
    module path: d = 24 gathers 12 live input columns, d = 36 has 18 (not a multiple of 4: mask folded instead)."""
    model = _perturbed_nvp(d, batch_norm, affine, seed=d)
    x = torch.rand(3000, d, device=DEV)
    with torch.no_grad():
        ll = model(x)
        u, ildj = model.apply_backward(x)
        xr, ldj = model.apply_forward(u)
        monkeypatch.setenv("DPK_FLOW_SIDE", "0")            # gather the live columns instead of the chained side output
        assert torch.equal(model(x), ll)
        monkeypatch.delenv("DPK_FLOW_SIDE")
        monkeypatch.setenv("DPK_LINEAR_MMA", "0")
        ll_ref = model(x)
        u_ref, ildj_ref = model.apply_backward(x)
        monkeypatch.delenv("DPK_LINEAR_MMA")
    assert rel_err(ll, ll_ref) < 1e-5
    assert float((u - u_ref).abs().max()) < 2e-5 * max(1.0, float(u_ref.abs().max()))
    if affine:
        assert float((ildj - ildj_ref).abs().max()) < 1e-4 * max(1.0, float(ildj_ref.abs().max()))
        assert float((ildj + ldj).abs().max()) < 1e-3
    assert float((xr - x).abs().max()) < 1e-4
    # in-place parameter / running-statistic updates are picked up by the cached compact weights and affine
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(1.05)
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.add_(0.1)
        ll2 = model(x)
        monkeypatch.setenv("DPK_LINEAR_MMA", "0")
        ll2_ref = model(x)
    assert rel_err(ll2, ll2_ref) < 1e-5
    assert rel_err(ll2, ll) > 1e-4


def test_compact_coupling_entry_matches_plain_layout():
    """dpk_coupling_forward_compact against dpk_coupling_forward on the scattered z, both directions."""
    import ctypes
    from deeprob_kit_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(11)
    batch, n = 257, 40
    inv = (torch.arange(n, device=DEV) % 3 != 0).float()
    live = torch.nonzero(inv).reshape(-1)
    zc = torch.randn(batch, 2 * live.numel(), device=DEV, generator=g)
    z = torch.randn(batch, 2 * n, device=DEV, generator=g)          # dead columns hold garbage on purpose
    z[:, live] = zc[:, :live.numel()]
    z[:, n + live] = zc[:, live.numel():]
    zmap = torch.zeros(n, dtype=torch.int32, device=DEV)
    zmap[live] = torch.arange(live.numel(), dtype=torch.int32, device=DEV)
    x = torch.randn(batch, n, device=DEV, generator=g)
    w = torch.tensor([0.7], device=DEV)
    pa = torch.rand(n, device=DEV, generator=g) + 0.5
    pc = torch.randn(n, device=DEV, generator=g)
    for direction in (0, 1):
        ref, ref_ldj = _engine.coupling(x, z, w, inv, n, 0, True, direction, 1)
        out = torch.empty_like(x)
        side = torch.full((batch, live.numel()), float("nan"), device=DEV)
        ldj = torch.zeros(batch, device=DEV)
        d = _engine._coupling_desc(batch, n, True, direction, w, 1, n, zc.shape[1], inv)
        rc = _lib.lib().dpk_coupling_forward_compact(
            ctypes.byref(d), _engine._ptr(x), _engine._ptr(zc), _engine._ptr(zmap), live.numel(), _engine._ptr(pa),
            _engine._ptr(pc), ctypes.c_float(0.25), _engine._ptr(out), n, _engine._ptr(side), _engine._ptr(ldj),
            _engine._stream(x.device))
        _lib.check(rc, "dpk_coupling_forward_compact")
        assert float((out - (ref * pa + pc)).abs().max()) < 1e-5
        assert torch.equal(side, out[:, live])
        assert float((ldj - (ref_ldj + 0.25)).abs().max()) < 1e-5


@pytest.mark.parametrize("activation", ["relu", "tanh"])
def test_maf_through_the_tensor_core_made(monkeypatch, activation):
    """MADE conditioner (MaskedLinear = mask * weight, deeprob/torch/utils.py:73-96) on the tcgen05 GEMM in inference
    against the stock modules; in-place parameter updates are picked up by the cached masked weights."""
    from deeprob_kit_b200.flows.models import MAF
    torch.manual_seed(5)
    model = MAF(64, n_flows=3, depth=2, units=128, batch_norm=True, activation=activation).to(DEV).eval()
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn_like(p))
    x = torch.randn(2500, 64, device=DEV)
    for _ in range(2):
        with torch.no_grad():
            monkeypatch.setenv("DPK_LINEAR_MMA", "1")
            ll1 = model(x)
            monkeypatch.setenv("DPK_LINEAR_MMA", "0")
            ll0 = model(x)
        assert rel_err(ll1, ll0) < 1e-5
        assert rel_err(ll0, model(x)) < 1e-6            # grad mode: stock modules
        with torch.no_grad():
            for p in model.parameters():
                p.mul_(1.03)


def test_chained_side_output_is_dropped_after_an_in_place_write():
    """The compact live-column copy a coupling leaves for the next one is tied to the tensor version: writing to the
    output in place invalidates it (the next layer gathers again)."""
    model = _perturbed_nvp(64, False, True, seed=9)
    x = torch.rand(2500, 64, device=DEV)
    c0, c1 = model.layers[0], model.layers[1]
    with torch.no_grad():
        h, _ = c0.apply_backward(x)
        assert getattr(h, "_dpk_live", None) is not None
        ref, _ = c1.apply_backward(h.clone().mul_(2.0))
        h.mul_(2.0)
        out, _ = c1.apply_backward(h)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("batch,k,n,relu,gscale", [(300, 64, 48, True, 1.0), (2500, 512, 1024, False, 1e-7),
                                                   (4096, 1536, 512, True, 3e-4), (257, 36, 260, True, 1e3)])
def test_linear_backward_matches_float64(batch, k, n, relu, gscale):
    """dpk_linear_backward (dgrad, split-K wgrad, bias sums on the tcgen05 GEMM, power-of-two operand scaling) against
    the float64 autograd of F.linear (+ReLU); gradients of very different magnitudes."""
    from deeprob_kit_b200.flows._engine import linear_train
    rng = np.random.RandomState(batch + k)
    x = torch.from_numpy(rng.standard_normal((batch, k)).astype(np.float32)).cuda()
    w = torch.from_numpy((rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)).cuda()
    b = torch.from_numpy(rng.standard_normal(n).astype(np.float32) * 0.1).cuda()
    dy = torch.from_numpy((rng.standard_normal((batch, n)) * rng.lognormal(0, 2, (batch, 1)) * gscale).astype(np.float32)).cuda()
    with torch.enable_grad():
        xa, wa, ba = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        y = linear_train(xa, wa, ba, relu)
        y.backward(dy)
        xr, wr, br = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
        yr = torch.nn.functional.linear(xr, wr, br)
        if relu:
            yr = torch.relu(yr)
        # the ReLU mask of the kernel is the sign of ITS forward output: mask the reference identically
        mask = (y > 0).double() if relu else torch.ones_like(yr)
        (torch.nn.functional.linear(xr, wr, br) * mask * dy.double()).sum().backward()
    assert rel_err(y.detach(), yr.detach()) < 1e-4
    for got, ref in ((xa.grad, xr.grad), (wa.grad, wr.grad), (ba.grad, br.grad)):
        assert norm_err(got, ref) < 1e-4, (float((got.double() - ref).abs().max()), float(ref.abs().max()))


def test_realnvp1d_training_gradients_through_the_tensor_core_mlp(monkeypatch):
    """BASELINE config 4 structure in training mode with every conditioner GEMM (forward, dgrad, wgrad) on tcgen05:
    the reference's training-mode log-likelihoods, and the gradients of the float64 oracle.  A ReLU network's
    gradient jumps where a hidden pre-activation crosses zero, so the comparison weights out the samples that come
    within 5e-5 of a kink in any hidden layer (an evaluation that differs by 1e-5 may legitimately sit on the other
    side; the fp32 reference has the same property, only with a narrower band)."""
    import oracle.flows_oracle as fo
    import param_gen as pg
    from conftest import load_golden
    from helpers import flow_reference_state
    monkeypatch.setenv("DPK_LINEAR_MMA_TRAIN", "1")
    cfg = pg.FLOW_CASES["nvp1d_cifar"]
    gold = load_golden("flows_nvp1d_cifar")
    model, state = flow_reference_state(cfg)
    # eval-mode batch norm keeps the samples independent (with batch statistics a kink of one sample leaks into
    # the gradient of all others); the gradients still flow through every conditioner GEMM
    model = model.cuda().eval()
    x, g = pg.flow_inputs(cfg)
    st64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() else v) for k, v in state.items()}
    fo.KINK_MARGINS = []
    try:
        with torch.enable_grad():
            ll64, _ = fo.flow1d_log_prob(x.double(), st64, "RealNVP1d", cfg["kw"], training=False)
        margin = torch.stack(fo.KINK_MARGINS).min(0).values
    finally:
        fo.KINK_MARGINS = None
    safe = (margin > 5e-5).double()
    assert float(safe.mean()) > 0.25, "too few samples away from the ReLU kinks to compare"
    ge = g.double() * safe
    with torch.enable_grad():
        (ll64 * ge).sum().backward()
    from deeprob_kit_b200 import _lib
    _lib.profile_read()
    with torch.enable_grad():
        out = model(x.cuda())
        (out * ge.float().cuda()).sum().backward()
    _, launches = _lib.profile_read()
    assert launches["gemm"] >= 8 * 3 * 2, launches          # 8 couplings x 3 layers x (forward + backward)
    assert rel_err(out.detach().cpu(), gold["ll"]) < 1e-4
    assert rel_err(out.detach().cpu(), ll64.detach()) < 1e-4
    gtol = 1e-4 + 4e-7 * float(out.abs().max())
    checked = 0
    for k, p in model.named_parameters():
        if p.grad is None or k not in st64 or st64[k].grad is None:
            continue
        assert norm_err(p.grad, st64[k].grad) < 3 * gtol, k
        checked += 1
    assert checked >= 8 * 6
