"""ctypes glue + autograd Functions of the DGC-SPN layers (csrc/dgcspn.cu).  Each layer is one
autograd node (the reference's `mpe` differentiates w.r.t. the base layer output, so the layers have
to stay separately differentiable: deeprob/spn/models/dgcspn.py:153-184)."""
import ctypes
import os

import torch

from .. import _lib
from ._engine import _PTR, _f32c, _ptr


def _stream(dev):
    return _PTR(_lib.stream_ptr(dev))


def _check4d(x, what):
    _lib.require_cuda(x, what)
    if x.dim() != 4:
        raise ValueError("%s: expected a (B, C, H, W) tensor, got %s" % (what, tuple(x.shape)))
    return _f32c(x)


class _Leaf(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, loc, scale):
        x, loc, scale = _f32c(x), _f32c(loc), _f32c(scale)
        b, cin, h, w = x.shape
        k = loc.shape[0]
        out = torch.empty(b, k, h, w, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().dpk_dgc_leaf_forward(_ptr(x), _ptr(loc), _ptr(scale), b, cin, k, h * w, _ptr(out), _stream(x.device))
        _lib.check(rc, "dpk_dgc_leaf_forward")
        ctx.save_for_backward(x, loc, scale)
        return out

    @staticmethod
    def backward(ctx, g):
        x, loc, scale = ctx.saved_tensors
        g = _f32c(g)
        b, cin, h, w = x.shape
        k = loc.shape[0]
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gl = torch.zeros_like(loc) if ctx.needs_input_grad[1] else None
        gs = torch.zeros_like(scale) if ctx.needs_input_grad[2] else None
        with torch.cuda.device(x.device):
            rc = _lib.lib().dpk_dgc_leaf_backward(_ptr(x), _ptr(loc), _ptr(scale), _ptr(g), b, cin, k, h * w, _ptr(gx),
                                                  _ptr(gl), _ptr(gs), _stream(x.device))
        _lib.check(rc, "dpk_dgc_leaf_backward")
        return gx, gl, gs


class _Product(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, desc):
        x = _f32c(x)
        b = x.shape[0]
        out = torch.empty(b, desc.out_channels, desc.out_height, desc.out_width, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().dpk_dgc_product_forward(ctypes.byref(desc), _ptr(x), b, _ptr(out), _stream(x.device))
        _lib.check(rc, "dpk_dgc_product_forward")
        ctx.desc, ctx.in_shape = desc, x.shape
        return out

    @staticmethod
    def backward(ctx, g):
        g = _f32c(g)
        gx = torch.empty(ctx.in_shape, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            rc = _lib.lib().dpk_dgc_product_backward(ctypes.byref(ctx.desc), _ptr(g), g.shape[0], _ptr(gx), _stream(g.device))
        _lib.check(rc, "dpk_dgc_product_backward")
        return gx, None


class _Sum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        x, weight = _f32c(x), _f32c(weight)
        b, cin, h, w = x.shape
        cout = weight.shape[0]
        out = torch.empty(b, cout, h, w, dtype=torch.float32, device=x.device)
        scratch = torch.empty(2 * weight.numel(), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().dpk_dgc_sum_forward(_ptr(x), _ptr(weight), b, cin, cout, h * w, _ptr(out), _ptr(scratch),
                                                _stream(x.device))
        _lib.check(rc, "dpk_dgc_sum_forward")
        ctx.save_for_backward(x, weight, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight, out = ctx.saved_tensors
        g = _f32c(g)
        b, cin, h, w = x.shape
        cout = weight.shape[0]
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gw = torch.zeros_like(weight) if ctx.needs_input_grad[1] else None
        scratch = torch.empty(3 * weight.numel(), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().dpk_dgc_sum_backward(_ptr(x), _ptr(weight), _ptr(out), _ptr(g), b, cin, cout, h * w, _ptr(gx),
                                                 _ptr(gw), _ptr(scratch), _stream(x.device))
        _lib.check(rc, "dpk_dgc_sum_backward")
        return gx, gw


class _Root(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        x, weight = _f32c(x), _f32c(weight)
        b = x.shape[0]
        flat = x.reshape(b, -1)
        c, q = weight.shape
        out = torch.empty(b, c, dtype=torch.float32, device=x.device)
        scratch = torch.empty(weight.numel(), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().dpk_dgc_root_forward(_ptr(flat), _ptr(weight), b, q, c, _ptr(out), _ptr(scratch), _stream(x.device))
        _lib.check(rc, "dpk_dgc_root_forward")
        ctx.save_for_backward(flat, weight, out)
        ctx.in_shape = x.shape
        return out

    @staticmethod
    def backward(ctx, g):
        flat, weight, out = ctx.saved_tensors
        g = _f32c(g)
        b = flat.shape[0]
        c, q = weight.shape
        gx = torch.empty_like(flat) if ctx.needs_input_grad[0] else None
        gw = torch.zeros_like(weight) if ctx.needs_input_grad[1] else None
        scratch = torch.empty(2 * weight.numel(), dtype=torch.float32, device=flat.device)
        with torch.cuda.device(flat.device):
            rc = _lib.lib().dpk_dgc_root_backward(_ptr(flat), _ptr(weight), _ptr(out), _ptr(g), b, q, c, _ptr(gx), _ptr(gw),
                                                  _ptr(scratch), _stream(flat.device))
        _lib.check(rc, "dpk_dgc_root_backward")
        return (gx.reshape(ctx.in_shape) if gx is not None else None), gw


def _check_shape(x, expect, what):
    """The kernels take C/H/W from the parameters / descriptor: a mismatching input would read out of bounds."""
    if tuple(x.shape[1:]) != tuple(expect):
        raise ValueError("%s: expected a (B, %s) tensor, got %s" % (what, ", ".join(str(int(e)) for e in expect), tuple(x.shape)))


def leaf(x, loc, scale):
    x = _check4d(x, "SpatialGaussianLayer.forward")
    _check_shape(x, loc.shape[1:], "SpatialGaussianLayer.forward")
    if tuple(scale.shape) != tuple(loc.shape):
        raise ValueError("SpatialGaussianLayer.forward: loc %s and scale %s differ in shape" % (tuple(loc.shape), tuple(scale.shape)))
    return _Leaf.apply(x, loc, scale)


def product(x, desc):
    x = _check4d(x, "SpatialProductLayer.forward")
    _check_shape(x, (desc.channels, desc.height, desc.width), "SpatialProductLayer.forward")
    return _Product.apply(x, desc)


def mixture(x, weight):
    x = _check4d(x, "SpatialSumLayer.forward")
    _check_shape(x, weight.shape[1:], "SpatialSumLayer.forward")
    return _Sum.apply(x, weight)


def root(x, weight):
    _lib.require_cuda(x, "SpatialRootLayer.forward")
    if x.dim() < 2 or x[0].numel() != weight.shape[1]:
        raise ValueError("SpatialRootLayer.forward: expected %d values per sample, got %s" % (weight.shape[1], tuple(x.shape)))
    return _Root.apply(x, weight)


def product_mixture(x, desc, weight):
    """Inference fusion of a depthwise SpatialProductLayer and the SpatialSumLayer behind it
    (dpk_dgc_prodsum_forward): the product output is never materialised.  No autograd node: callers use it only
    when no gradient is needed."""
    x, weight = _check4d(x, "SpatialProductLayer.forward"), _f32c(weight.detach())
    _check_shape(x, (desc.channels, desc.height, desc.width), "SpatialProductLayer.forward")
    if tuple(weight.shape[1:]) != (desc.out_channels, desc.out_height, desc.out_width):
        raise ValueError("SpatialSumLayer.forward: weight %s does not match the product output" % (tuple(weight.shape),))
    b = x.shape[0]
    cout = weight.shape[0]
    out = torch.empty(b, cout, desc.out_height, desc.out_width, dtype=torch.float32, device=x.device)
    scratch = torch.empty(2 * weight.numel(), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().dpk_dgc_prodsum_forward(ctypes.byref(desc), _ptr(x), _ptr(weight), b, cout, _ptr(out),
                                                _ptr(scratch), _stream(x.device))
    _lib.check(rc, "dpk_dgc_prodsum_forward")
    return out


class _ProductSum(torch.autograd.Function):
    """Training form of the fusion: the forward is the same single kernel (the product output is not kept); the backward
    recomputes the product output from the saved input (one gather pass), then runs the sum and product backward kernels."""

    @staticmethod
    def forward(ctx, x, weight, desc):
        out = product_mixture(x, desc, weight)
        ctx.save_for_backward(_f32c(x), _f32c(weight), out)
        ctx.desc = desc
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight, out = ctx.saved_tensors
        desc = ctx.desc
        g = _f32c(g)
        b, cout = x.shape[0], weight.shape[0]
        hw = desc.out_height * desc.out_width
        dev = x.device
        gw = torch.zeros_like(weight) if ctx.needs_input_grad[1] else None
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gprod = (torch.empty(b, desc.out_channels, desc.out_height, desc.out_width, dtype=torch.float32, device=dev)
                 if gx is not None else None)
        scratch = torch.empty(3 * weight.numel(), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = _stream(dev)
            # DPK_DGC_BWD_GATHER=1: product values recomputed inside the sum-backward kernel from the four taps (no
            # temporary, one pass less) -- measured slower, 38.9 vs 32.8 ms at config 3: four predicated loads per value
            if desc.out_channels <= 8 and cout <= 8 and os.environ.get("DPK_DGC_BWD_GATHER", "0") == "1":
                rc = _lib.lib().dpk_dgc_prodsum_backward(ctypes.byref(desc), _ptr(x), _ptr(weight), _ptr(out), _ptr(g), b, cout,
                                                         _ptr(gprod), _ptr(gw), _ptr(scratch), st)
                _lib.check(rc, "dpk_dgc_prodsum_backward")
            else:
                prod = torch.empty(b, desc.out_channels, desc.out_height, desc.out_width, dtype=torch.float32, device=dev)
                rc = _lib.lib().dpk_dgc_product_forward(ctypes.byref(desc), _ptr(x), b, _ptr(prod), st)
                _lib.check(rc, "dpk_dgc_product_forward")
                rc = _lib.lib().dpk_dgc_sum_backward(_ptr(prod), _ptr(weight), _ptr(out), _ptr(g), b, desc.out_channels, cout, hw,
                                                     _ptr(gprod), _ptr(gw), _ptr(scratch), st)
                _lib.check(rc, "dpk_dgc_sum_backward")
                del prod
            if gx is not None:
                rc = _lib.lib().dpk_dgc_product_backward(ctypes.byref(desc), _ptr(gprod), b, _ptr(gx), st)
                _lib.check(rc, "dpk_dgc_product_backward")
        return gx, gw, None


def product_mixture_train(x, desc, weight):
    """product_mixture with an autograd node (see _ProductSum)."""
    x = _check4d(x, "SpatialProductLayer.forward")
    return _ProductSum.apply(x, weight, desc)


def can_fuse_product_mixture(prod_layer, sum_layer) -> bool:
    d = prod_layer._desc
    ok = (bool(d.depthwise) and d.channels == d.out_channels and d.channels in (2, 4, 8, 16, 32)
          and tuple(sum_layer.weight.shape[1:]) == (d.out_channels, d.out_height, d.out_width))
    if d.channels == 32:      # weights through L1: output chunks of 8, no shared-memory limit
        return ok
    return ok and d.channels * min(8, max(2, sum_layer.weight.shape[0])) * 512 <= 96 * 1024
