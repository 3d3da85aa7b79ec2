"""ctypes glue + autograd Functions of the flow bijectors (csrc/flows.cu)."""
import ctypes

import torch

from .. import _lib
from ..spn._engine import _PTR, _f32c, _ptr


def _stream(dev):
    return _PTR(_lib.stream_ptr(dev))


def _coupling_desc(batch, n, affine, direction, w, w_inner, x_stride, z_stride, inv_mask):
    d = _lib.CouplingDesc()
    d.batch, d.features, d.affine, d.direction = batch, n, 1 if affine else 0, direction
    d.w_count = w.numel() if w is not None else 1
    d.w_inner = w_inner
    d.x_stride, d.z_stride = x_stride, z_stride
    d.inv_mask = inv_mask.data_ptr() if inv_mask is not None else None
    d.scale_weight = w.data_ptr() if w is not None else None
    return d


class _Coupling(torch.autograd.Function):
    """(x, z, w) -> (out, log_det).  `offset`/`n` select the transformed slice of every sample row; the rest of
    the row passes through unchanged (channel-wise couplings)."""

    @staticmethod
    def forward(ctx, x, z, w, inv_mask, n, offset, affine, direction, w_inner):
        _lib.require_cuda(x, "coupling layer")
        x, z = _f32c(x), _f32c(z)
        ctx.w_shape = w.shape if w is not None else None
        w = _f32c(w).reshape(-1) if w is not None else None
        batch = x.shape[0]
        row = x[0].numel() if batch else n
        out = torch.empty_like(x) if n == row else x.clone()
        ldj = torch.zeros(batch, dtype=torch.float32, device=x.device)
        d = _coupling_desc(batch, n, affine, direction, w, w_inner, row, z[0].numel() if batch else 2 * n, inv_mask)
        with torch.cuda.device(x.device):
            rc = _lib.lib().dpk_coupling_forward(ctypes.byref(d), _PTR(x.data_ptr() + 4 * offset), _ptr(z),
                                                 _PTR(out.data_ptr() + 4 * offset), row, _ptr(ldj), _stream(x.device))
        _lib.check(rc, "dpk_coupling_forward")
        ctx.save_for_backward(x, z, w if w is not None else x.new_empty(0), inv_mask if inv_mask is not None else x.new_empty(0))
        ctx.cfg = (n, offset, affine, direction, w_inner, row)
        return out, ldj

    @staticmethod
    def backward(ctx, gout, gldj):
        x, z, w, inv_mask = ctx.saved_tensors
        n, offset, affine, direction, w_inner, row = ctx.cfg
        w = w if w.numel() else None
        inv_mask = inv_mask if inv_mask.numel() else None
        gout = _f32c(gout)
        batch = x.shape[0]
        gx = torch.empty_like(x) if n == row else gout.clone()
        gz = torch.empty_like(z)
        gw = torch.zeros_like(w) if (w is not None and ctx.needs_input_grad[2]) else None
        gl = _f32c(gldj) if gldj is not None else None
        d = _coupling_desc(batch, n, affine, direction, w, w_inner, row, z[0].numel() if batch else 2 * n, inv_mask)
        with torch.cuda.device(x.device):
            rc = _lib.lib().dpk_coupling_backward(
                ctypes.byref(d), _PTR(x.data_ptr() + 4 * offset), _ptr(z), _PTR(gout.data_ptr() + 4 * offset), row,
                _ptr(gl), _PTR(gx.data_ptr() + 4 * offset), row, _ptr(gz), z[0].numel() if batch else 2 * n, _ptr(gw),
                _stream(x.device))
        _lib.check(rc, "dpk_coupling_backward")
        if gw is not None:
            gw = gw.reshape(ctx.w_shape)
        return gx, gz, gw, None, None, None, None, None, None


def coupling(x, z, w, inv_mask, n, offset, affine, direction, w_inner):
    return _Coupling.apply(x, z, w, inv_mask, n, offset, affine, direction, w_inner)


# ------------------------------------------------------------------------------------------------
# batch-norm bijector
# ------------------------------------------------------------------------------------------------
def _feature_reduce(x, center, other, features, inner, mode):
    s = torch.zeros(features, dtype=torch.float32, device=x.device)
    d = torch.zeros(features, dtype=torch.float32, device=x.device) if mode == 2 else None
    batch = x.shape[0]
    with torch.cuda.device(x.device):
        rc = _lib.lib().dpk_feature_reduce(_ptr(x), _ptr(center), _ptr(other), _ptr(s), _ptr(d), batch, features, inner,
                                           mode, _stream(x.device))
    _lib.check(rc, "dpk_feature_reduce")
    return s, d


def _feature_affine(x, a, c, features, inner, y=None, k=None, mu=None):
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.lib().dpk_feature_affine(_ptr(x), _ptr(a), _ptr(c), _ptr(y), _ptr(k), _ptr(mu), _ptr(out), x.shape[0],
                                           features, inner, _stream(x.device))
    _lib.check(rc, "dpk_feature_affine")
    return out


class _BatchNorm(torch.autograd.Function):
    """u = (x - mean) / sqrt(var + eps) * exp(weight) + bias, log_det = sum(weight - log(var+eps)/2) * inner.
    direction 1 is the inverse map (always with the running statistics)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training, momentum, eps, unbiased, direction):
        _lib.require_cuda(x, "batch-norm bijector")
        x = _f32c(x)
        batch, features = x.shape[0], x.shape[1]
        inner = x[0, 0].numel() if x.dim() > 2 else 1
        n = batch * inner
        w, b = weight.detach().reshape(-1).float(), bias.detach().reshape(-1).float()
        use_batch = training and direction == 0
        if use_batch:
            s, _ = _feature_reduce(x, None, None, features, inner, 0)
            mean = s / n
            m2, _ = _feature_reduce(x, mean, None, features, inner, 1)
            var = m2 / (n - 1 if unbiased else n)
            with torch.no_grad():
                running_var.mul_(momentum).add_(var.view_as(running_var) * (1.0 - momentum))
                running_mean.mul_(momentum).add_(mean.view_as(running_mean) * (1.0 - momentum))
        else:
            mean, var = running_mean.reshape(-1).float(), running_var.reshape(-1).float()
        veps = var + eps
        if direction == 0:
            a = torch.rsqrt(veps) * torch.exp(w)
            c = b - mean * a
            ldj = torch.sum(w - 0.5 * torch.log(veps)) * inner
        else:
            a = torch.sqrt(veps) * torch.exp(-w)
            c = mean - b * a
            ldj = torch.sum(0.5 * torch.log(veps) - w) * inner
        out = _feature_affine(x, a.contiguous(), c.contiguous(), features, inner)
        ctx.save_for_backward(x, w, b, mean, veps, a)
        ctx.cfg = (use_batch, unbiased, direction, inner, features, n, weight.shape)
        return out, ldj.expand(batch)

    @staticmethod
    def backward(ctx, gout, gldj):
        x, w, b, mean, veps, a = ctx.saved_tensors
        use_batch, unbiased, direction, inner, features, n, wshape = ctx.cfg
        gout = _f32c(gout)
        gi = gldj.sum() if gldj is not None else gout.new_zeros(())
        if direction == 0:
            s1, s2 = _feature_reduce(x, mean.contiguous(), gout, features, inner, 2)   # sum g, sum g*(x-mean)
            gw = s2 * a + gi * inner
            gb = s1
            if use_batch:
                gv = s2 * torch.exp(w) * (-0.5) * veps.pow(-1.5) + gi * inner * (-0.5) / veps
                gmu = -a * s1
                denom = (n - 1) if unbiased else n
                gx = _feature_affine(gout, a.contiguous(), (gmu / n).contiguous(), features, inner, y=x,
                                     k=(2.0 * gv / denom).contiguous(), mu=mean.contiguous())
            else:
                gx = _feature_affine(gout, a.contiguous(), torch.zeros_like(a), features, inner)
        else:   # x_out = (u - b) * a + mean,  a = sqrt(veps) * exp(-w)
            s1, s2 = _feature_reduce(x, b.contiguous(), gout, features, inner, 2)       # sum g, sum g*(u-b)
            gw = -s2 * a - gi * inner
            gb = -s1 * a
            gx = _feature_affine(gout, a.contiguous(), torch.zeros_like(a), features, inner)
        return gx, gw.reshape(wshape), gb.reshape(wshape), None, None, None, None, None, None, None


def batch_norm(x, weight, bias, running_mean, running_var, training, momentum, eps, unbiased, direction=0):
    return _BatchNorm.apply(x, weight, bias, running_mean, running_var, training, momentum, eps, unbiased, direction)


# ------------------------------------------------------------------------------------------------
# preprocessing and prior
# ------------------------------------------------------------------------------------------------
class _Preprocess(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, noise, bins, alpha):
        _lib.require_cuda(x, "flow preprocessing")
        x = _f32c(x)
        batch = x.shape[0]
        n = x[0].numel() if batch else 1
        out = torch.empty_like(x)
        ildj = torch.zeros(batch, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().dpk_flow_preprocess_forward(_ptr(x), _ptr(noise), float(bins), float(alpha), _ptr(out),
                                                        _ptr(ildj), batch, n, _stream(x.device))
        _lib.check(rc, "dpk_flow_preprocess_forward")
        ctx.save_for_backward(x, noise if noise is not None else x.new_empty(0))
        ctx.cfg = (bins, alpha, n)
        return out, ildj

    @staticmethod
    def backward(ctx, gout, gildj):
        x, noise = ctx.saved_tensors
        bins, alpha, n = ctx.cfg
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            rc = _lib.lib().dpk_flow_preprocess_backward(_ptr(x), _ptr(noise if noise.numel() else None), float(bins),
                                                         float(alpha), _ptr(_f32c(gout)),
                                                         _ptr(_f32c(gildj) if gildj is not None else None), _ptr(gx),
                                                         x.shape[0], n, _stream(x.device))
        _lib.check(rc, "dpk_flow_preprocess_backward")
        return gx, None, None, None


def preprocess(x, noise, bins, alpha):
    return _Preprocess.apply(x, noise, bins, alpha)


class _NormalPrior(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, ildj, loc, scale):
        _lib.require_cuda(z, "flow prior")
        z = _f32c(z)
        batch = z.shape[0]
        n = z[0].numel() if batch else 1
        out = torch.empty(batch, dtype=torch.float32, device=z.device)
        il = _f32c(ildj) if isinstance(ildj, torch.Tensor) else None
        with torch.cuda.device(z.device):
            rc = _lib.lib().dpk_normal_prior_forward(_ptr(z), _ptr(loc), _ptr(scale), _ptr(il), _ptr(out), batch, n,
                                                     _stream(z.device))
        _lib.check(rc, "dpk_normal_prior_forward")
        ctx.save_for_backward(z, loc if loc is not None else z.new_empty(0), scale if scale is not None else z.new_empty(0))
        ctx.n = n
        return out

    @staticmethod
    def backward(ctx, gout):
        z, loc, scale = ctx.saved_tensors
        gout = _f32c(gout)
        gz = torch.empty_like(z)
        with torch.cuda.device(z.device):
            rc = _lib.lib().dpk_normal_prior_backward(_ptr(z), _ptr(loc if loc.numel() else None),
                                                      _ptr(scale if scale.numel() else None), _ptr(gout), _ptr(gz),
                                                      z.shape[0], ctx.n, _stream(z.device))
        _lib.check(rc, "dpk_normal_prior_backward")
        return gz, (gout if ctx.needs_input_grad[1] else None), None, None


def normal_prior(z, ildj, loc, scale):
    return _NormalPrior.apply(z, ildj, loc, scale)


# ---- conditioner MLP on the tensor cores (inference) ---------------------------------------------------------
MLP_MIN_BATCH = 2048


def linear(x, weight, bias, relu, cache=None):
    """act(x @ weight.T + bias) through dpk_linear_forward (tcgen05 GEMM, fp32-accurate 3-pass fp16 operands).
    `cache`: dict owned by the calling layer -- workspace + weight signature, so the weight images are rebuilt only
    when the weight changed (data pointer / version counter, like RatSpn._workspace)."""
    import ctypes
    import os
    from .. import _lib
    x = x.contiguous()
    batch, k = x.shape
    n = weight.shape[0]
    out = torch.empty(batch, n, dtype=torch.float32, device=x.device)
    nbytes = _lib.lib().dpk_linear_workspace_bytes(batch, k, n)
    cache = cache if cache is not None else {}
    key = (str(x.device), torch.cuda.current_stream(x.device).cuda_stream)
    ws = cache.get(key)
    if ws is None or ws[0].numel() < nbytes:
        ws = [torch.empty(nbytes, dtype=torch.uint8, device=x.device), None]
        cache[key] = ws
    sig = (batch, weight.data_ptr(), weight._version)
    flags = _lib.F_TABLES_VALID if (ws[1] == sig and os.environ.get("DPK_TABLE_CACHE", "1") != "0") else 0
    ws[1] = sig
    w = weight.detach()
    b = bias.detach() if bias is not None else None
    with torch.cuda.device(x.device):
        rc = _lib.lib().dpk_linear_forward(
            ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(w.data_ptr()),
            ctypes.c_void_p(b.data_ptr() if b is not None else 0), batch, k, n, 1 if relu else 0,
            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(ws[0].data_ptr()), ws[0].numel(), flags,
            ctypes.c_void_p(_lib.stream_ptr(x.device)))
    _lib.check(rc, "dpk_linear_forward")
    return out


class _LinearMMA(torch.autograd.Function):
    """act(x @ weight.T + bias) with forward AND backward on the tcgen05 GEMM (dpk_linear_forward / dpk_linear_backward):
    the training path of the conditioner MLPs (deeprob/flows/layers/coupling.py:45-56, autoregressive.py:72-79)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        x = x.contiguous()
        w = weight.contiguous()
        out = linear(x, w, bias, relu, None)
        ctx.relu = relu
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, w, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, dy):
        import ctypes
        from .. import _lib
        x, w, y = ctx.saved_tensors
        dy = dy.contiguous()
        batch, k = x.shape
        n = w.shape[0]
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        dx = torch.empty_like(x) if need_x else None
        dw = torch.empty_like(w) if need_w else None
        db = torch.empty(n, dtype=torch.float32, device=x.device) if need_b else None
        nbytes = _lib.lib().dpk_linear_backward_workspace_bytes(batch, k, n)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().dpk_linear_backward(_ptr(x), _ptr(w), _ptr(y), _ptr(dy), batch, k, n, 1 if ctx.relu else 0,
                                                _ptr(dx), _ptr(dw), _ptr(db), _ptr(ws), ws.numel(), _stream(x.device))
        _lib.check(rc, "dpk_linear_backward")
        return dx, dw, db, None


def linear_train(x, weight, bias, relu):
    return _LinearMMA.apply(x, weight, bias, relu)


def _mlp_train_plan(network, x):
    """Like _mlp_plan, for a call that needs gradients (DPK_LINEAR_MMA_TRAIN=0 keeps the stock modules / cuBLAS;
    =1 forces the tensor-core path for small batches, as the tests do)."""
    import os
    from torch import nn
    from ..torch.utils import MaskedLinear
    knob = os.environ.get("DPK_LINEAR_MMA_TRAIN", "")
    if (not torch.is_grad_enabled() or not x.is_cuda or x.dim() != 2 or x.dtype != torch.float32 or knob == "0"
            or (x.shape[0] < MLP_MIN_BATCH and knob != "1")):
        return None
    mods = list(network)
    plan = []
    i = 0
    while i < len(mods):
        m = mods[i]
        if type(m) not in (nn.Linear, MaskedLinear) or m.in_features % 4 or m.weight.dtype != torch.float32:
            return None
        act = mods[i + 1] if i + 1 < len(mods) and not isinstance(mods[i + 1], nn.Linear) else None
        relu = isinstance(act, nn.ReLU)
        plan.append((m, relu, None if relu else act))
        i += 1 if act is None else 2
    return plan or None


def _mlp_plan(network, x):
    """[(Linear, fused_relu, activation_module_or_None)] when `network` is a stack of nn.Linear / MaskedLinear layers
    with activations that the tcgen05 GEMM can evaluate for this call (inference on a large fp32 CUDA batch), else
    None.  A ReLU is fused into the GEMM epilogue; any other activation module runs as its own elementwise op."""
    import os
    from torch import nn
    from ..torch.utils import MaskedLinear
    if (torch.is_grad_enabled() or not x.is_cuda or x.dim() != 2 or x.dtype != torch.float32
            or x.shape[0] < MLP_MIN_BATCH or os.environ.get("DPK_LINEAR_MMA", "1") == "0"):
        return None
    mods = list(network)
    plan = []
    i = 0
    while i < len(mods):
        m = mods[i]
        if (type(m) not in (nn.Linear, MaskedLinear) or m.in_features % 4 or m.weight.dtype != torch.float32
                or not m.weight.is_contiguous()):
            return None
        act = mods[i + 1] if i + 1 < len(mods) and not isinstance(mods[i + 1], nn.Linear) else None
        relu = isinstance(act, nn.ReLU)
        plan.append((m, relu, None if relu else act))
        i += 1 if act is None else 2
    return plan or None


def _layer_weight(m, cache):
    """Effective weight of a conditioner layer: MaskedLinear multiplies by its fixed 0/1 mask
    (deeprob/torch/utils.py:73-96); cached until the parameter changes."""
    mask = getattr(m, "mask", None)
    if mask is None:
        return m.weight
    return _derived(cache, "made_w", (m.weight, mask), lambda: mask * m.weight.detach())


def _derived(cache, name, sources, make):
    """Tensor derived from parameters, rebuilt IN PLACE (so that its own version counter moves and `linear` refreshes
    its operand images) only when a source tensor changed."""
    sig = tuple((t.data_ptr(), t._version) for t in sources)
    if cache.get(name + "_sig") != sig:
        new = make()
        old = cache.get(name)
        if old is not None and old.shape == new.shape and old.device == new.device:
            old.copy_(new)
        else:
            cache[name] = new.contiguous()
        cache[name + "_sig"] = sig
    return cache[name]


def mlp(network, x, in_mask=None):
    """Evaluate an nn.Sequential of Linear / ReLU layers on `in_mask * x` (the conditioner of CouplingLayer1d,
    deeprob/flows/layers/coupling.py:45-56,72-75).  Inference on a large CUDA batch goes through `linear`, with the
    0/1 input mask folded into the first layer's weight columns ((m*x) W^T = x (W*m)^T: no masked copy of x is made);
    training (any gradient needed), small batches and unusual layer stacks use the stock modules (cuBLAS)."""
    plan = _mlp_plan(network, x)
    if plan is None:
        tplan = _mlp_train_plan(network, x)
        if tplan is None:
            return network(x if in_mask is None else in_mask * x)
        # gradients needed: every layer is one autograd node whose forward and backward are tcgen05 GEMMs; the 0/1
        # masks (conditioner input, MADE connectivity) are folded into the weights by differentiable tensor ops
        for li, (m, relu, act) in enumerate(tplan):
            weight = m.weight if getattr(m, "mask", None) is None else m.mask * m.weight
            if li == 0 and in_mask is not None:
                weight = weight * in_mask.reshape(1, -1)
            x = linear_train(x, weight, m.bias, relu)
            if act is not None:
                x = act(x)
        return x
    for li, (m, relu, act) in enumerate(plan):
        cache = m.__dict__.setdefault("_dpk_linear_cache", {})
        weight = _layer_weight(m, cache)
        if li == 0 and in_mask is not None:
            src = weight
            weight = _derived(cache, "masked_w", (src, in_mask), lambda: src.detach() * in_mask.reshape(1, -1))
        x = linear(x, weight, m.bias, relu, cache)
        if act is not None:
            x = act(x)
    return x


def eval_batch_norm_affine(bn):
    """(a, c, log_det) of an eval-mode batch-norm bijector in the density direction, u = a*x + c
    (deeprob/flows/utils.py:118-139 with the running statistics); cached on the layer until a tensor changes
    (one host read of the log-det constant per change)."""
    ts = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
    sig = tuple((t.data_ptr(), t._version) for t in ts)
    c = bn.__dict__.get("_dpk_eval_cache")
    if c is None or c[0] != sig:
        with torch.no_grad():
            veps = bn.running_var.reshape(-1).float() + bn.eps
            w = bn.weight.reshape(-1).float()
            a = (torch.rsqrt(veps) * torch.exp(w)).contiguous()
            shift = (bn.bias.reshape(-1).float() - bn.running_mean.reshape(-1).float() * a).contiguous()
            ldj = float(torch.sum(w - 0.5 * torch.log(veps)))
        c = (sig, a, shift, ldj)
        bn.__dict__["_dpk_eval_cache"] = c
    return c[1], c[2], c[3]


def coupling1d_infer(layer, x, direction, bn=None):
    """Inference fast path of CouplingLayer1d (deeprob/flows/layers/coupling.py:72-104): returns (out, log_det) or
    None when it does not apply (gradients needed, small batch, unusual conditioner).

    The reference evaluates the conditioner on mask*x and multiplies t and s by inv_mask, so (a) the input columns
    with mask == 0 and (b) the output columns with inv_mask == 0 never reach the result.  Here the first Linear runs
    on the gathered live input columns (K halves), the last Linear only produces the live t|s columns (N halves),
    and the coupling kernel reads that compact z; `bn` (an eval-mode BatchNormLayer1d that follows in the density
    direction) is applied in the same pass."""
    import os
    plan = _mlp_plan(layer.network, x)
    if plan is None or len(plan) < 2 or os.environ.get("DPK_FLOW_COMPACT", "1") == "0":
        return None
    x = x.contiguous()
    batch, n = x.shape
    cache = layer.__dict__.setdefault("_dpk_cache", {})
    key = (str(x.device), layer.mask._version, layer.inv_mask._version)
    if cache.get("key") != key:
        mask, inv = layer.mask.reshape(-1), layer.inv_mask.reshape(-1)
        live_in = torch.nonzero(mask != 0).reshape(-1)
        live_out = torch.nonzero(inv != 0).reshape(-1)
        zmap = torch.zeros(n, dtype=torch.int32, device=x.device)
        zmap[live_out] = torch.arange(live_out.numel(), dtype=torch.int32, device=x.device)
        rows = torch.cat([live_out, live_out + n]) if layer.affine else live_out
        binary = bool(((mask == 0) | (mask == 1)).all())
        cache.clear()
        cache.update(key=key, live_in=live_in, live_out=live_out, zmap=zmap, rows=rows, binary=binary)
    live_in, live_out, rows = cache["live_in"], cache["live_out"], cache["rows"]
    if live_out.numel() == 0:
        return None
    if any(type(m) is not torch.nn.Linear or act is not None for m, _, act in plan):
        return None
    first, last = plan[0][0], plan[-1][0]
    # first layer: live input columns when the mask is 0/1 and their count keeps 16-byte rows, else the fold.  The
    # previous coupling of the chain leaves exactly these columns next to its output (live_out of the C entry)
    # when its transformed set is this layer's conditioner input (alternating masks); otherwise they are gathered.
    if cache["binary"] and live_in.numel() and live_in.numel() % 4 == 0:
        h = None
        hint = getattr(x, "_dpk_live", None)
        if hint is not None and hint[2] == x._version and hint[1].shape[0] == batch:
            same = cache.setdefault("hint_ok", {})      # index tensor (kept alive: its address stays unique) -> verdict
            known = same.get(hint[0].data_ptr())
            if known is None or known[0] is not hint[0]:
                known = (hint[0], hint[0].shape == live_in.shape and bool(torch.equal(hint[0], live_in)))
                same[hint[0].data_ptr()] = known
            if known[1]:
                h = hint[1]
        if h is None:
            h = x.index_select(1, live_in)
        w0 = _derived(cache, "w_first", (first.weight,), lambda: first.weight.detach().index_select(1, live_in))
    else:
        h = x
        w0 = _derived(cache, "w_first", (first.weight, layer.mask),
                      lambda: first.weight.detach() * layer.mask.reshape(1, -1))
    h = linear(h, w0, first.bias, plan[0][1], first.__dict__.setdefault("_dpk_linear_cache", {}))
    for m, relu, _ in plan[1:-1]:
        h = linear(h, m.weight, m.bias, relu, m.__dict__.setdefault("_dpk_linear_cache", {}))
    w_l = _derived(cache, "w_last", (last.weight,), lambda: last.weight.detach().index_select(0, rows))
    b_l = None
    if last.bias is not None:
        b_l = _derived(cache, "b_last", (last.bias,), lambda: last.bias.detach().index_select(0, rows))
    z = linear(h, w_l, b_l, plan[-1][1], last.__dict__.setdefault("_dpk_linear_cache", {}))
    w = _f32c(layer.scale_act.weight.detach()).reshape(-1) if layer.affine else None
    post_a = post_c = None
    post_ldj = 0.0
    if bn is not None:
        post_a, post_c, post_ldj = eval_batch_norm_affine(bn)
    out = torch.empty_like(x)
    ldj = torch.zeros(batch, dtype=torch.float32, device=x.device)
    side = None
    if direction == 0 and os.environ.get("DPK_FLOW_SIDE", "1") != "0":
        side = torch.empty(batch, live_out.numel(), dtype=torch.float32, device=x.device)
    d = _coupling_desc(batch, n, layer.affine, direction, w, 1, n, z.shape[1], layer.inv_mask)
    with torch.cuda.device(x.device):
        rc = _lib.lib().dpk_coupling_forward_compact(
            ctypes.byref(d), _ptr(x), _ptr(z), _ptr(cache["zmap"]), live_out.numel(), _ptr(post_a), _ptr(post_c),
            ctypes.c_float(post_ldj), _ptr(out), n, _ptr(side), _ptr(ldj), _stream(x.device))
    _lib.check(rc, "dpk_coupling_forward_compact")
    if side is not None:
        out._dpk_live = (live_out, side, out._version)      # valid for this tensor object until it is written to
    return out, ldj
