"""ctypes glue between the SPN nn.Modules and libdeeprob_b200.so (RAT-SPN part).

Everything here is plumbing: build the POD descriptor from the module's parameters, allocate the
workspace as a torch tensor, pass raw device pointers + the current CUDA stream to the C ABI.
"""
import ctypes
from typing import List, Optional

import torch

from .. import _lib

_PTR = ctypes.c_void_p


def _f32c(t: torch.Tensor) -> torch.Tensor:
    """Contiguous fp32 view/copy of a tensor (parameters already are: no copy in the common case)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return _PTR(t.data_ptr()) if t is not None else _PTR(0)


class RatSpnCall:
    """One descriptor + the tensors it borrows (kept alive for the duration of the call / autograd node)."""

    def __init__(self, base_layer, sum_weights: List[torch.Tensor], root_weight: Optional[torch.Tensor],
                 out_classes: int, sum_nodes: int, repetitions: int, device, force_scale: bool = False):
        p0, p1 = base_layer.leaf_parameters()
        self.n_leaf = 1 if p1 is None else 2          # leaf tensors among the autograd inputs
        if p1 is not None and not force_scale and getattr(base_layer, "unit_scale", lambda: False)():
            p1 = None                                  # frozen scale == 1: NULL selects the unit-scale kernels
        self.keep = [_f32c(p0), _f32c(p1) if p1 is not None else None]
        self.sums = [_f32c(w) for w in sum_weights]
        self.root = _f32c(root_weight) if root_weight is not None else None
        for t in [*self.keep, *self.sums, self.root, base_layer._mask_i32, base_layer._region_len]:
            if t is not None and t.device != device:
                raise RuntimeError("model parameters live on %s but the input is on %s" % (t.device, device))
        if len(self.sums) > _lib.MAX_LEVELS:
            raise ValueError("region graph too deep for the kernels (max %d levels)" % _lib.MAX_LEVELS)
        d = _lib.RatSpnDesc()
        d.leaf_kind = base_layer.leaf_kind
        d.in_features = base_layer.in_features
        d.depth = base_layer.rg_depth
        d.repetitions = repetitions
        d.leaf_channels = base_layer.out_channels
        d.sum_nodes = sum_nodes
        d.out_classes = out_classes
        d.dimension = base_layer.dimension
        d.mask = base_layer._mask_i32.data_ptr()
        d.region_len = base_layer._region_len.data_ptr()
        d.leaf_p0 = self.keep[0].data_ptr()
        d.leaf_p1 = self.keep[1].data_ptr() if self.keep[1] is not None else None
        for i, w in enumerate(self.sums):
            d.sum_weight[i] = w.data_ptr()
        d.root_weight = self.root.data_ptr() if self.root is not None else None
        self.desc = d

    def workspace_bytes(self, batch: int, flags: int) -> int:
        n = _lib.lib().dpk_ratspn_workspace_bytes(ctypes.byref(self.desc), batch, flags)
        if n == 0:
            _lib.check(-1, "dpk_ratspn_workspace_bytes")
        return n

    def workspace(self, batch: int, flags: int, device) -> torch.Tensor:
        return torch.empty(self.workspace_bytes(batch, flags), dtype=torch.uint8, device=device)


def _check_input(x: torch.Tensor, features: int, what: str) -> torch.Tensor:
    _lib.require_cuda(x, what)
    if x.dim() != 2 or x.shape[1] != features:
        raise ValueError("%s: expected a (batch, %d) tensor, got %s" % (what, features, tuple(x.shape)))
    return _f32c(x)


def ratspn_leaf_forward(layer, x: torch.Tensor) -> torch.Tensor:
    """RegionGraphLayer.forward through dpk_ratspn_leaf_forward: (B, D) -> (B, G0, K)."""
    x = _check_input(x, layer.in_features, "RegionGraphLayer.forward")
    reps = layer.in_regions >> layer.rg_depth
    call = RatSpnCall(layer, [], None, 1, 1, reps, x.device)
    batch = x.shape[0]
    out = torch.empty(batch, layer.in_regions, layer.out_channels, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        ws = call.workspace(batch, 0, x.device)
        rc = _lib.lib().dpk_ratspn_leaf_forward(ctypes.byref(call.desc), _ptr(x), batch, _ptr(out), _ptr(ws),
                                                ws.numel(), _PTR(_lib.stream_ptr(x.device)))
    _lib.check(rc, "dpk_ratspn_leaf_forward")
    return out


def outer_sum_forward(x: torch.Tensor, partitions: int, nodes: int) -> torch.Tensor:
    """ProductLayer.forward through dpk_outer_sum_forward: (B, 2P, K) -> (B, P, K*K)."""
    _lib.require_cuda(x, "ProductLayer.forward")
    x = _f32c(x)
    batch = x.shape[0]
    out = torch.empty(batch, partitions, nodes * nodes, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().dpk_outer_sum_forward(_ptr(x), batch, partitions, nodes, _ptr(out),
                                              _PTR(_lib.stream_ptr(x.device)))
    _lib.check(rc, "dpk_outer_sum_forward")
    return out


def mixture_forward(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """SumLayer/RootLayer.forward through dpk_mixture_forward: (B, P, Kin) x (P, O, Kin) -> (B, P, O)."""
    _lib.require_cuda(x, "SumLayer.forward")
    x, w = _f32c(x), _f32c(weight.detach())
    batch, parts, kin = x.shape
    outs = w.shape[1]
    out = torch.empty(batch, parts, outs, dtype=torch.float32, device=x.device)
    scratch = torch.empty(parts * outs, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().dpk_mixture_forward(_ptr(x), _ptr(w), batch, parts, kin, outs, _ptr(out), _ptr(scratch),
                                            _PTR(_lib.stream_ptr(x.device)))
    _lib.check(rc, "dpk_mixture_forward")
    return out


class _RatSpnLogProb(torch.autograd.Function):
    """log_prob of the whole RAT-SPN: dpk_ratspn_forward / dpk_ratspn_backward."""

    @staticmethod
    def forward(ctx, model_and_mode, x, *params):
        # needs_input_grad reflects requires_grad of the inputs even when the caller is in no_grad mode (and grad
        # mode is always off inside forward): the caller's mode is passed in explicitly
        model, grad_mode = model_and_mode
        need_grad = grad_mode and any(ctx.needs_input_grad)
        call = model._make_call(x.device)
        batch = x.shape[0]
        flags = _lib.F_SAVE_ACTIVATIONS if need_grad else 0
        out = torch.empty(batch, model.out_classes, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            ws, extra = model._workspace(call, batch, flags, x.device, private=need_grad)
            rc = _lib.lib().dpk_ratspn_forward(ctypes.byref(call.desc), _ptr(x), batch, _ptr(out), _ptr(ws),
                                               ws.numel(), flags | extra, _PTR(_lib.stream_ptr(x.device)))
        _lib.check(rc, "dpk_ratspn_forward")
        if need_grad:
            ctx.call, ctx.ws, ctx.model = call, ws, model
            ctx.save_for_backward(x, out)
            ctx.param_shapes = [p.shape for p in params]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, out = ctx.saved_tensors
        call, ws, model = ctx.call, ctx.ws, ctx.model
        need = ctx.needs_input_grad        # (model, x, *params)
        batch = x.shape[0]
        g = _lib.RatSpnGrads()
        gx = torch.zeros_like(x) if need[1] else None
        g.grad_x = gx.data_ptr() if gx is not None else None
        n_leaf = call.n_leaf
        grads = []
        for i, shape in enumerate(ctx.param_shapes):
            grads.append(torch.zeros(shape, dtype=torch.float32, device=x.device) if need[2 + i] else None)
        g.leaf_p0 = grads[0].data_ptr() if grads[0] is not None else None
        if n_leaf == 2:
            g.leaf_p1 = grads[1].data_ptr() if grads[1] is not None else None
        for i in range(len(call.sums)):
            t = grads[n_leaf + i]
            g.sum_weight[i] = t.data_ptr() if t is not None else None
        g.root_weight = grads[-1].data_ptr() if grads[-1] is not None else None
        grad_out = _f32c(grad_out)
        with torch.cuda.device(x.device):
            rc = _lib.lib().dpk_ratspn_backward(ctypes.byref(call.desc), _ptr(x), batch, _ptr(out), _ptr(grad_out),
                                                ctypes.byref(g), _ptr(ws), ws.numel(),
                                                _PTR(_lib.stream_ptr(x.device)))
        _lib.check(rc, "dpk_ratspn_backward")
        return (None, gx, *grads)


class _RatSpnLogProbDropout(torch.autograd.Function):
    """Training-mode log_prob with probabilistic dropout: dpk_ratspn_forward_dropout / dpk_ratspn_backward_dropout.
    The masks are regenerated in the backward from the seed drawn here (one value per forward, taken from torch's
    CPU generator so that torch.manual_seed makes a run reproducible)."""

    @staticmethod
    def forward(ctx, model, x, *params):
        call = model._make_call(x.device, force_scale=True)
        batch = x.shape[0]
        drop = _lib.RatSpnDropout()
        drop.in_rate = float(model.in_dropout or 0.0)
        drop.sum_rate = float(model.sum_dropout or 0.0)
        drop.seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        out = torch.empty(batch, model.out_classes, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            nbytes = _lib.lib().dpk_ratspn_dropout_workspace_bytes(ctypes.byref(call.desc), batch)
            if nbytes == 0:
                _lib.check(-1, "dpk_ratspn_dropout_workspace_bytes")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            rc = _lib.lib().dpk_ratspn_forward_dropout(ctypes.byref(call.desc), _ptr(x), batch, ctypes.byref(drop), _ptr(out),
                                                       _ptr(ws), ws.numel(), _PTR(_lib.stream_ptr(x.device)))
        _lib.check(rc, "dpk_ratspn_forward_dropout")
        ctx.call, ctx.ws, ctx.drop = call, ws, drop
        ctx.save_for_backward(x, out)
        ctx.param_shapes = [p.shape for p in params]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, out = ctx.saved_tensors
        call, ws = ctx.call, ctx.ws
        need = ctx.needs_input_grad        # (model, x, *params)
        batch = x.shape[0]
        g = _lib.RatSpnGrads()
        gx = torch.zeros_like(x) if need[1] else None
        g.grad_x = gx.data_ptr() if gx is not None else None
        grads = [torch.zeros(shape, dtype=torch.float32, device=x.device) if need[2 + i] else None
                 for i, shape in enumerate(ctx.param_shapes)]
        n_leaf = call.n_leaf
        g.leaf_p0 = grads[0].data_ptr() if grads[0] is not None else None
        if n_leaf == 2:
            g.leaf_p1 = grads[1].data_ptr() if grads[1] is not None else None
        for i in range(len(call.sums)):
            t = grads[n_leaf + i]
            g.sum_weight[i] = t.data_ptr() if t is not None else None
        g.root_weight = grads[-1].data_ptr() if grads[-1] is not None else None
        grad_out = _f32c(grad_out)
        with torch.cuda.device(x.device):
            rc = _lib.lib().dpk_ratspn_backward_dropout(ctypes.byref(call.desc), _ptr(x), batch, ctypes.byref(ctx.drop),
                                                        _ptr(out), _ptr(grad_out), ctypes.byref(g), _ptr(ws), ws.numel(),
                                                        _PTR(_lib.stream_ptr(x.device)))
        _lib.check(rc, "dpk_ratspn_backward_dropout")
        return (None, gx, *grads)


def ratspn_log_prob(model, x: torch.Tensor) -> torch.Tensor:
    x = _check_input(x, model.in_features, "RatSpn.forward")
    if model.training and (model.in_dropout is not None or model.sum_dropout is not None):
        return _RatSpnLogProbDropout.apply(model, x, *model._kernel_parameters())
    return _RatSpnLogProb.apply((model, torch.is_grad_enabled()), x, *model._kernel_parameters())


class _NllLoss(torch.autograd.Function):
    """loss(ll, y) of the SPN models as one kernel (dpk_nll_loss); the gradient w.r.t. ll is produced by the same launch."""

    @staticmethod
    def forward(ctx, ll, y):
        ll = _f32c(ll)
        batch, classes = ll.shape
        loss = torch.empty((), dtype=torch.float32, device=ll.device)
        grad = torch.empty_like(ll)
        yi = y.to(device=ll.device, dtype=torch.int64).contiguous() if y is not None else None
        with torch.cuda.device(ll.device):
            rc = _lib.lib().dpk_nll_loss(_ptr(ll), _ptr(yi), batch, classes, _ptr(loss), _ptr(grad),
                                         _PTR(_lib.stream_ptr(ll.device)))
        _lib.check(rc, "dpk_nll_loss")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def nll_loss(ll: torch.Tensor, y: Optional[torch.Tensor]) -> torch.Tensor:
    """-mean(ll) for one output class, else the mean cross entropy of log_softmax(ll) against the labels y."""
    if ll.is_cuda and ll.dim() == 2 and ll.shape[0] > 0 and (ll.shape[1] == 1 or y is not None):
        return _NllLoss.apply(ll, y)
    if ll.shape[-1] == 1 or y is None:            # host tensors (e.g. logged outputs): plain tensor ops
        return -torch.mean(ll)
    return torch.nn.functional.nll_loss(torch.log_softmax(ll, dim=1), y)


def ratspn_mpe(model, x: torch.Tensor, y: Optional[torch.Tensor]) -> torch.Tensor:
    """RatSpn.mpe: forward keeping every level's log-likelihoods, then the top-down kernel (dpk_ratspn_mpe)."""
    x = _check_input(x, model.in_features, "RatSpn.mpe")
    call = model._make_call(x.device)
    batch, dev = x.shape[0], x.device
    out = torch.empty(batch, model.out_classes, dtype=torch.float32, device=dev)
    filled = torch.empty_like(x)
    yi = y.to(device=dev, dtype=torch.int32).contiguous() if y is not None else None
    with torch.cuda.device(dev):
        ws = call.workspace(batch, _lib.F_SAVE_ACTIVATIONS, dev)
        sp = _PTR(_lib.stream_ptr(dev))
        rc = _lib.lib().dpk_ratspn_forward(ctypes.byref(call.desc), _ptr(x), batch, _ptr(out), _ptr(ws), ws.numel(),
                                           _lib.F_SAVE_ACTIVATIONS, sp)
        _lib.check(rc, "dpk_ratspn_forward")
        rc = _lib.lib().dpk_ratspn_mpe(ctypes.byref(call.desc), _ptr(x), batch, _ptr(out), _ptr(yi), _ptr(filled), _ptr(ws),
                                       ws.numel(), sp)
    _lib.check(rc, "dpk_ratspn_mpe")
    return filled


def ratspn_sample(model, n_samples: int, y: Optional[torch.Tensor], device) -> torch.Tensor:
    """RatSpn.sample: ancestral sampling kernel (dpk_ratspn_sample); the seed comes from torch's CPU generator."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("RatSpn.sample: deeprob_kit_b200 runs on CUDA devices only; there is no CPU path")
    call = model._make_call(device, force_scale=True)
    out = torch.empty(n_samples, model.in_features, dtype=torch.float32, device=device)
    yi = y.to(device=device, dtype=torch.int32).contiguous() if y is not None else None
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    with torch.cuda.device(device):
        ws = call.workspace(n_samples, _lib.F_SAVE_ACTIVATIONS, device)
        rc = _lib.lib().dpk_ratspn_sample(ctypes.byref(call.desc), n_samples, _ptr(yi), seed, _ptr(out), _ptr(ws), ws.numel(),
                                          _PTR(_lib.stream_ptr(device)))
    _lib.check(rc, "dpk_ratspn_sample")
    return out


def ratspn_em_statistics(model, x: torch.Tensor):
    """E-step sufficient statistics of a batch (dict of tensors; see include/deeprob_b200.h)."""
    x = _check_input(x, model.in_features, "RatSpn.em_statistics")
    call = model._make_call(x.device)
    batch = x.shape[0]
    dev = x.device
    out = torch.empty(batch, model.out_classes, dtype=torch.float32, device=dev)
    base = model.base_layer
    shape = (base.in_regions, base.out_channels, base.dimension)
    stats = {
        "sum_counts": [torch.zeros_like(w) for w in call.sums],
        "root_counts": torch.zeros_like(call.root),
        "s0": torch.zeros(shape, dtype=torch.float32, device=dev),
        "s1": torch.zeros(shape, dtype=torch.float32, device=dev),
        "s2": torch.zeros(shape, dtype=torch.float32, device=dev) if call.keep[1] is not None else None,
    }
    st = _lib.RatSpnEmStats()
    for i, t in enumerate(stats["sum_counts"]):
        st.sum_counts[i] = t.data_ptr()
    st.root_counts = stats["root_counts"].data_ptr()
    st.s0, st.s1 = stats["s0"].data_ptr(), stats["s1"].data_ptr()
    st.s2 = stats["s2"].data_ptr() if stats["s2"] is not None else None
    with torch.cuda.device(dev):
        ws, extra = model._workspace(call, batch, _lib.F_SAVE_ACTIVATIONS, dev, private=False)
        sp = _PTR(_lib.stream_ptr(dev))
        rc = _lib.lib().dpk_ratspn_forward(ctypes.byref(call.desc), _ptr(x), batch, _ptr(out), _ptr(ws), ws.numel(),
                                           _lib.F_SAVE_ACTIVATIONS | extra, sp)
        _lib.check(rc, "dpk_ratspn_forward")
        rc = _lib.lib().dpk_ratspn_em_statistics(ctypes.byref(call.desc), _ptr(x), batch, _ptr(out), ctypes.byref(st),
                                                 _ptr(ws), ws.numel(), sp)
        _lib.check(rc, "dpk_ratspn_em_statistics")
    stats["ll"] = out
    return stats
