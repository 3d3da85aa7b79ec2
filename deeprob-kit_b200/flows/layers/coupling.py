"""RealNVP / NICE coupling layers (interface of deeprob/flows/layers/coupling.py: CouplingLayer1d :15-104,
CouplingLayer2d :107-272, CouplingBlock2d :275-408).  The 1-D conditioner MLP is evaluated by the tcgen05 GEMM of
csrc/ratspn_leaf_mma.cu in inference (dpk_linear_forward) and by library GEMMs when gradients are needed; the 2-D
conditioners are library convolutions; everything after it -- chunk, ScaledTanh, masking, affine transform, per-sample
log-det reduction -- is ONE kernel (csrc/flows.cu: dpk_coupling_forward/backward)."""
from typing import Tuple

import numpy as np
import torch
from torch import nn

from ...torch.utils import ScaledTanh
from .. import _engine
from ..utils import BatchNormLayer2d, Bijector, squeeze_depth2d, unsqueeze_depth2d
from .densenet import DenseNetwork
from .resnet import ResidualNetwork


class _Fp32ConvBackward(torch.autograd.Function):
    """Identity behind a conv conditioner.  Its backward runs before every backward node of the conditioner: it
    turns cuDNN's TF32 convolutions off and queues the restore for the end of the running backward pass."""

    @staticmethod
    def forward(ctx, z):
        return z.view_as(z)

    @staticmethod
    def backward(ctx, g):
        prev = torch.backends.cudnn.allow_tf32
        if prev:
            torch.backends.cudnn.allow_tf32 = False

            def restore():
                torch.backends.cudnn.allow_tf32 = True
            torch.autograd.Variable._execution_engine.queue_callback(restore)
        return g


class CouplingLayer1d(Bijector):
    def __init__(self, in_features: int, depth: int, units: int, affine: bool = True, reverse: bool = False):
        super().__init__(in_features)
        self.affine = affine
        self.reverse = reverse
        mask, inv_mask = self.build_alternating_masks()
        if reverse:
            mask, inv_mask = inv_mask, mask
        self.register_buffer('mask', torch.tensor(mask, dtype=torch.float32))
        self.register_buffer('inv_mask', torch.tensor(inv_mask, dtype=torch.float32))
        dims = [in_features] + [units] * depth
        layers = []
        for a, b in zip(dims[:-1], dims[1:]):
            layers += [nn.Linear(a, b), nn.ReLU(inplace=True)]
        layers.append(nn.Linear(dims[-1], in_features * 2 if affine else in_features))
        self.network = nn.Sequential(*layers)
        if affine:
            self.scale_act = ScaledTanh()

    def build_alternating_masks(self) -> Tuple[np.ndarray, np.ndarray]:
        mask = np.arange(self.in_features) % 2
        return mask, 1.0 - mask

    def _transform(self, x, direction, bn=None):
        # inference: compact conditioner (live input / output columns only) + optional fused eval batch-norm
        fast = _engine.coupling1d_infer(self, x, direction, bn)
        if fast is not None:
            return fast
        if bn is not None:
            return None
        z = _engine.mlp(self.network, x, self.mask)
        w = self.scale_act.weight if self.affine else None
        out, ldj = _engine.coupling(x, z, w, self.inv_mask, self.in_features, 0, self.affine, direction, 1)
        return out, (ldj if self.affine else 0.0)

    def apply_backward(self, x):
        return self._transform(x, 0)

    def apply_backward_with(self, x, bn):
        """apply_backward of this layer followed by apply_backward of the eval-mode batch-norm bijector `bn`, in one
        pass when the inference fast path applies; None otherwise (the caller then runs the two layers)."""
        return self._transform(x, 0, bn)

    def apply_forward(self, u):
        return self._transform(u, 1)


class CouplingLayer2d(Bijector):
    def __init__(self, in_features: Tuple[int, int, int], network: str, n_blocks: int, channels: int,
                 affine: bool = True, channelwise: bool = False, reverse: bool = False):
        super().__init__(in_features)
        self.affine = affine
        self.channelwise = channelwise
        self.reverse = reverse
        if not channelwise:
            mask, inv_mask = self.build_checkerboard_masks()
            if reverse:
                mask, inv_mask = inv_mask, mask
            self.register_buffer('mask', torch.tensor(mask, dtype=torch.float32))
            self.register_buffer('inv_mask', torch.tensor(inv_mask, dtype=torch.float32))
            # the kernel reads a mask entry per transformed element: expand over the channels once
            full = np.broadcast_to(inv_mask, (self.in_channels, self.in_height, self.in_width)).reshape(-1)
            self.register_buffer('_inv_mask_flat', torch.tensor(np.ascontiguousarray(full), dtype=torch.float32),
                                 persistent=False)
        cin = self.in_channels // 2 if channelwise else self.in_channels
        cout = cin * 2 if affine else cin
        if network == 'resnet':
            self.network = ResidualNetwork(cin, channels, cout, n_blocks)
        elif network == 'densenet':
            self.network = DenseNetwork(cin, channels, cout, n_blocks)
        else:
            raise NotImplementedError("Unknown network conditioner {}".format(network))
        if affine:
            self.scale_act = ScaledTanh([cin, 1, 1])

    @property
    def in_channels(self) -> int:
        return self.in_features[0]

    @property
    def in_height(self) -> int:
        return self.in_features[1]

    @property
    def in_width(self) -> int:
        return self.in_features[2]

    def build_checkerboard_masks(self) -> Tuple[np.ndarray, np.ndarray]:
        mask = np.sum(np.indices([1, self.in_height, self.in_width]), axis=0) % 2
        return mask, 1.0 - mask

    def _conditioner(self, x):
        # cuDNN would otherwise run the fp32 conditioner convolutions (forward AND backward) on TF32 tensor cores
        # (10-bit mantissa); the 1e-4 parity bar of the log-likelihood needs true fp32.  The switch is scoped to
        # this conditioner: off around its forward, and off again from the moment autograd enters its backward
        # until the end of that backward pass (see _Fp32ConvBackward); the process-wide setting is restored.
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            z = self.network(x)
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        return _Fp32ConvBackward.apply(z) if z.requires_grad else z

    def _transform(self, x, direction):
        hw = self.in_height * self.in_width
        w = self.scale_act.weight if self.affine else None
        if self.channelwise:
            half = (self.in_channels // 2) * hw
            first, second = torch.chunk(x, chunks=2, dim=1)
            cond = first if self.reverse else second           # untouched half feeds the conditioner
            offset = half if self.reverse else 0               # the other half is transformed in place
            z = self._conditioner(cond)
            out, ldj = _engine.coupling(x, z, w, None, half, offset, self.affine, direction, hw)
        else:
            z = self._conditioner(self.mask * x)
            out, ldj = _engine.coupling(x, z, w, self._inv_mask_flat, self.in_channels * hw, 0, self.affine, direction, hw)
        return out, (ldj if self.affine else 0.0)

    def apply_backward(self, x):
        return self._transform(x, 0)

    def apply_forward(self, u):
        return self._transform(u, 1)


class CouplingBlock2d(Bijector):
    def __init__(self, in_features: Tuple[int, int, int], network: str, n_blocks: int, channels: int,
                 affine: bool = True, last_block: bool = False):
        super().__init__(in_features)
        self.last_block = last_block
        c, h, w = self.in_features

        def checker(reverse):
            return [CouplingLayer2d(self.in_features, network, n_blocks, channels, affine, channelwise=False,
                                    reverse=reverse), BatchNormLayer2d(c)]

        self.in_couplings = nn.ModuleList(checker(False) + checker(True) + checker(False))
        if last_block:
            self.in_couplings.extend(checker(True))
        else:
            squeezed = (c * 4, h // 2, w // 2)

            def chanwise(reverse):
                return [CouplingLayer2d(squeezed, network, n_blocks, channels * 2, affine, channelwise=True,
                                        reverse=reverse), BatchNormLayer2d(c * 4)]

            self.out_couplings = nn.ModuleList(chanwise(False) + chanwise(True) + chanwise(False))

    @property
    def in_channels(self) -> int:
        return self.in_features[0]

    @property
    def in_height(self) -> int:
        return self.in_features[1]

    @property
    def in_width(self) -> int:
        return self.in_features[2]

    def apply_backward(self, x):
        total = 0.0
        for layer in self.in_couplings:
            x, ildj = layer.apply_backward(x)
            total = total + ildj
        if not self.last_block:
            x = squeeze_depth2d(x)
            for layer in self.out_couplings:
                x, ildj = layer.apply_backward(x)
                total = total + ildj
            x = unsqueeze_depth2d(x)
        return x, total

    def apply_forward(self, u):
        total = 0.0
        if not self.last_block:
            u = squeeze_depth2d(u)
            for layer in reversed(self.out_couplings):
                u, ldj = layer.apply_forward(u)
                total = total + ldj
            u = unsqueeze_depth2d(u)
        for layer in reversed(self.in_couplings):
            u, ldj = layer.apply_forward(u)
            total = total + ldj
        return u, total
