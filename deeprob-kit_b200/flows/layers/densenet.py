"""Dense conv conditioner (interface and state_dict layout of deeprob/flows/layers/densenet.py:11-189);
convolutions are library calls, see resnet.py."""
from typing import List

import torch
from torch import nn
from torch.utils.checkpoint import checkpoint

from ...torch.utils import WeightNormConv2d


def _bn_relu_conv(cin, cout, k, pad, bias):
    return nn.Sequential(nn.BatchNorm2d(cin), nn.ReLU(inplace=True),
                         WeightNormConv2d(cin, cout, kernel_size=k, padding=pad, bias=bias))


class DenseLayer(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, use_checkpoint: bool = False):
        super().__init__()
        self.use_checkpoint = use_checkpoint
        mid = 4 * out_channels
        self.bottleneck_network = _bn_relu_conv(in_channels, mid, 1, 0, False)
        self.network = _bn_relu_conv(mid, out_channels, 3, 1, False)

    def bottleneck(self, inputs: List[torch.Tensor]) -> torch.Tensor:
        return self.bottleneck_network(torch.cat(inputs, dim=1))

    def forward(self, inputs: List[torch.Tensor]) -> torch.Tensor:
        if self.use_checkpoint and any(t.requires_grad for t in inputs):
            h = checkpoint(lambda *ts: self.bottleneck(ts), *inputs)
        else:
            h = self.bottleneck(inputs)
        return self.network(h)


class DenseBlock(nn.Module):
    def __init__(self, n_layers: int, in_channels: int, out_channels: int, use_checkpoint: bool = False):
        super().__init__()
        self.layers = nn.ModuleList(
            DenseLayer(in_channels + i * out_channels, out_channels, use_checkpoint=use_checkpoint) for i in range(n_layers)
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        feats = [x]
        for layer in self.layers:
            feats.append(layer(feats))
        return torch.cat(feats, dim=1)


class Transition(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.network = _bn_relu_conv(in_channels, out_channels, 1, 0, bias)

    def forward(self, x):
        return self.network(x)


class DenseNetwork(nn.Module):
    def __init__(self, in_channels: int, mid_channels: int, out_channels: int, n_blocks: int, use_checkpoint: bool = False):
        super().__init__()
        self.blocks = nn.ModuleList()
        self.in_conv = WeightNormConv2d(in_channels, mid_channels, kernel_size=3, padding=1, bias=False)
        for i in range(n_blocks):
            self.blocks.append(DenseBlock(4, mid_channels, mid_channels, use_checkpoint=use_checkpoint))
            last = i == n_blocks - 1
            self.blocks.append(Transition(5 * mid_channels, out_channels if last else mid_channels, bias=last))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = self.in_conv(x)
        for block in self.blocks:
            h = block(h)
        return h
